"""Joining several LoRA files into one wider LoRA, and driving the joined model.

Mirrors `lora_diffusion/lora_manager.py` of the reference:
  * `lora_join`              lora_manager.py:13-71
  * `DummySafeTensorObject`  lora_manager.py:74-86
  * `LoRAManager`            lora_manager.py:89-144

A join of LoRAs with ranks r_1..r_n is the rank-(sum r_i) LoRA whose down factor stacks the A_i
along the rank axis and whose up factor concatenates the B_i along the rank axis, so that
B·A = sum_i B_i·A_i; per-LoRA strengths are then a diagonal between the factors
(`set_lora_diag`, applied inside the fused kernel as its `diag` argument). The fused kernels hold
the rank axis in 16 operand rows, so a joined rank above 16 raises `LoraB200Error` at the first
forward (lora_b200/modules.py); joining itself is file-level work and has no such limit.
"""
from typing import List

import torch

from .patch import (apply_learned_embed_in_clip, monkeypatch_or_replace_safeloras, set_lora_diag)
from .persist import parse_safeloras_embeds

EMBED_FLAG = "<embed>"


def _file_rank(metadata: dict) -> int:
    """The single rank a file declares in its `*:rank` metadata entries (0 when it has none)."""
    ranks = {int(v) for k, v in metadata.items() if k.endswith("rank")}
    assert len(ranks) <= 1, "Rank should be the same per model"
    return ranks.pop() if ranks else 0


def lora_join(lora_safetenors: list):
    """-> (tensors, metadata, ranklist, token_size_list); argument name as in the reference.

    Each input only needs `.keys()`, `.metadata()`, `.get_tensor(key)` (a `safe_open` handle or a
    `DummySafeTensorObject`). Embedding tokens are renamed `<s{file}-{j}>` in sorted-token order;
    every `unet:*`/`text_encoder:*` rank entry of the result carries the summed rank.
    """
    metas = [dict(f.metadata()) for f in lora_safetenors]
    ranklist = [_file_rank(m) for m in metas]
    total_rank = sum(ranklist)

    merged_meta = {}
    for m in metas:
        merged_meta.update(m)
    metadata = {k: v for k, v in merged_meta.items() if v != EMBED_FLAG}

    all_keys = set()
    for f in lora_safetenors:
        all_keys.update(f.keys())

    tensors = {}
    for key in all_keys:
        if not (key.startswith("text_encoder") or key.startswith("unet")):
            continue
        parts = [f.get_tensor(key) for f in lora_safetenors]
        axis = 0 if key.endswith("down") else 1              # rank axis of A [r,K] / B [N,r]
        joined = torch.cat(parts, dim=axis)
        assert joined.shape[axis] == total_rank
        tensors[key] = joined
        metadata[key.rsplit(":", 1)[0] + ":rank"] = str(total_rank)

    token_size_list = []
    for i, f in enumerate(lora_safetenors):
        tokens = sorted(k for k, v in f.metadata().items() if v == EMBED_FLAG)
        for j, tok in enumerate(tokens):
            new = f"<s{i}-{j}>"
            tensors[new] = f.get_tensor(tok)
            metadata[new] = EMBED_FLAG
            print(f"Embedding {tok} replaced to {new}")
        token_size_list.append(len(tokens))
    return tensors, metadata, ranklist, token_size_list


class DummySafeTensorObject:
    """In-memory stand-in for a `safe_open` handle (what `lora_join` returns, re-wrapped)."""

    def __init__(self, tensor: dict, metadata):
        self.tensor = tensor
        self._metadata = metadata

    def keys(self):
        return self.tensor.keys()

    def metadata(self):
        return self._metadata

    def get_tensor(self, key):
        return self.tensor[key]


class LoRAManager:
    """Several LoRA files patched into one pipeline as a single joined LoRA.

    `pipe` is duck-typed: `.unet`, `.text_encoder`, `.tokenizer` (a diffusers
    StableDiffusionPipeline in the reference). `tune(scales)` sets one strength per source file
    through the selector diagonal; `prompt(text)` expands `<k>` into file k's embedding tokens.
    """

    def __init__(self, lora_paths_list: List[str], pipe):
        self.lora_paths_list = lora_paths_list
        self.pipe = pipe
        self._setup()

    def _setup(self):
        from safetensors import safe_open
        self._lora_safetenors = [safe_open(p, framework="pt", device="cpu") for p in self.lora_paths_list]
        tensors, metadata, self.ranklist, self.token_size_list = lora_join(self._lora_safetenors)
        self.total_safelora = DummySafeTensorObject(tensors, metadata)
        monkeypatch_or_replace_safeloras(self.pipe, self.total_safelora)
        tok_dict = parse_safeloras_embeds(self.total_safelora)
        apply_learned_embed_in_clip(tok_dict, self.pipe.text_encoder, self.pipe.tokenizer,
                                    token=None, idempotent=True)

    def tune(self, scales):
        assert len(scales) == len(self.ranklist), "Scale list should be the same length as ranklist"
        diag = [s for s, r in zip(scales, self.ranklist) for _ in range(r)]
        set_lora_diag(self.pipe.unet, torch.tensor(diag))

    def prompt(self, prompt):
        if prompt is not None:
            for i, n_tok in enumerate(self.token_size_list):
                prompt = prompt.replace(f"<{i + 1}>", "".join(f"<s{i}-{j}>" for j in range(n_tok)))
        return prompt
