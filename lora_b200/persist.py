"""On-disk LoRA formats (bit-compatible with the reference's files).

Contract restated from /root/reference/lora_diffusion/lora.py:424-632,1045-1110 and pinned by
the ten fixtures in /root/reference/example_loras (tests/golden/example_loras_manifest.json):

  .pt           torch.save([up0, down0, up1, down1, ...]) as fp16 CPU tensors, raw factors
                (NOT multiplied by scale)                                   lora.py:424-436
  .safetensors  tensors  "{model}:{i}:up" / "{model}:{i}:down"   (up is pre-multiplied by the
                module's scale via realize_as_lora; fp16)                   lora.py:451-483
                metadata "{model}" -> json list of target class names,
                         "{model}:{i}:rank" -> str(rank),
                textual-inversion embeds: tensor key = token, metadata token -> "<embed>"
"""
import json
from itertools import groupby
from typing import Dict, List, Optional, Set, Tuple

import torch
import torch.nn as nn

from .inject import (DEFAULT_TARGET_REPLACE, EMBED_FLAG, TEXT_ENCODER_DEFAULT_TARGET_REPLACE,
                     extract_lora_as_tensor, extract_lora_ups_down)

try:
    from safetensors.torch import safe_open
    from safetensors.torch import save_file as safe_save

    safetensors_available = True
except ImportError:  # pragma: no cover - safetensors ships in this image
    safetensors_available = False

    def safe_open(*a, **k):
        raise EnvironmentError("safetensors is required to read .safetensors LoRA files")

    def safe_save(tensors, filename, metadata=None):
        raise EnvironmentError(
            "Saving safetensors requires the safetensors library. Please install with pip or similar."
        )


def save_lora_weight(model, path="./lora.pt", target_replace_module=DEFAULT_TARGET_REPLACE):
    flat = []
    for up, down in extract_lora_ups_down(model, target_replace_module=target_replace_module):
        flat.append(up.weight.to("cpu").to(torch.float16))
        flat.append(down.weight.to("cpu").to(torch.float16))
    torch.save(flat, path)


def save_lora_as_json(model, path="./lora.json"):
    flat = []
    for up, down in extract_lora_ups_down(model):
        flat.append(up.weight.detach().cpu().numpy().tolist())
        flat.append(down.weight.detach().cpu().numpy().tolist())
    with open(path, "w") as fh:
        json.dump(flat, fh)


def _add_embeds(weights, metadata, embeds):
    for token, tensor in embeds.items():
        metadata[token] = EMBED_FLAG
        weights[token] = tensor


def save_safeloras_with_embeds(
    modelmap: Dict[str, Tuple[nn.Module, Set[str]]] = {},
    embeds: Dict[str, torch.Tensor] = {},
    outpath="./lora.safetensors",
):
    """modelmap: {"model name": (module, target_replace_module)} -> one safetensors file."""
    weights, metadata = {}, {}
    for name, (model, targets) in modelmap.items():
        metadata[name] = json.dumps(list(targets))
        for i, (up, down) in enumerate(extract_lora_as_tensor(model, targets)):
            metadata[f"{name}:{i}:rank"] = str(down.shape[0])
            # safetensors refuses non-contiguous / shared storage; factors may be arena views
            weights[f"{name}:{i}:up"] = up.contiguous()
            weights[f"{name}:{i}:down"] = down.contiguous()
    _add_embeds(weights, metadata, embeds)
    print(f"Saving weights to {outpath}")
    safe_save(weights, outpath, metadata)


def save_safeloras(modelmap: Dict[str, Tuple[nn.Module, Set[str]]] = {}, outpath="./lora.safetensors"):
    return save_safeloras_with_embeds(modelmap=modelmap, outpath=outpath)


def convert_loras_to_safeloras_with_embeds(
    modelmap: Dict[str, Tuple[str, Set[str], int]] = {},
    embeds: Dict[str, torch.Tensor] = {},
    outpath="./lora.safetensors",
):
    """modelmap: {"model name": (path to .pt list, target_replace_module, rank)}."""
    weights, metadata = {}, {}
    for name, (path, targets, r) in modelmap.items():
        metadata[name] = json.dumps(list(targets))
        for flat_idx, w in enumerate(torch.load(path)):
            site, is_down = divmod(flat_idx, 2)
            if is_down:
                weights[f"{name}:{site}:down"] = w
            else:
                metadata[f"{name}:{site}:rank"] = str(r)
                weights[f"{name}:{site}:up"] = w
    _add_embeds(weights, metadata, embeds)
    print(f"Saving weights to {outpath}")
    safe_save(weights, outpath, metadata)


def convert_loras_to_safeloras(modelmap: Dict[str, Tuple[str, Set[str], int]] = {},
                               outpath="./lora.safetensors"):
    convert_loras_to_safeloras_with_embeds(modelmap=modelmap, outpath=outpath)


def parse_safeloras(safeloras) -> Dict[str, Tuple[List[nn.parameter.Parameter], List[int], List[str]]]:
    """{model: ([up0, down0, ...] Parameters, [rank per site], target list)} (lora.py:538-596)."""
    meta = safeloras.metadata()
    model_of = lambda key: key.split(":")[0]
    keys = sorted(safeloras.keys(), key=model_of)
    out = {}
    for name, group in groupby(keys, model_of):
        info = meta.get(name)
        if not info:
            raise ValueError(f"Tensor {name} has no metadata - is this a Lora safetensor?")
        if info == EMBED_FLAG:
            continue
        group = list(group)
        n_sites = len(group) // 2
        ranks = [4] * n_sites
        weights = [None] * len(group)
        for key in group:
            _, idx, direction = key.split(":")
            idx = int(idx)
            ranks[idx] = int(meta[f"{name}:{idx}:rank"])
            weights[2 * idx + (direction == "down")] = nn.parameter.Parameter(safeloras.get_tensor(key))
        out[name] = (weights, ranks, json.loads(info))
    return out


def parse_safeloras_embeds(safeloras) -> Dict[str, torch.Tensor]:
    meta = safeloras.metadata()
    return {k: safeloras.get_tensor(k) for k in safeloras.keys() if meta.get(k) == EMBED_FLAG}


def load_safeloras(path, device="cpu"):
    return parse_safeloras(safe_open(path, framework="pt", device=device))


def load_safeloras_embeds(path, device="cpu"):
    return parse_safeloras_embeds(safe_open(path, framework="pt", device=device))


def load_safeloras_both(path, device="cpu"):
    f = safe_open(path, framework="pt", device=device)
    return parse_safeloras(f), parse_safeloras_embeds(f)


def _text_lora_path(path: str) -> str:
    assert path.endswith(".pt"), "Only .pt files are supported"
    return ".".join(path.split(".")[:-1] + ["text_encoder", "pt"])


def _ti_lora_path(path: str) -> str:
    assert path.endswith(".pt"), "Only .pt files are supported"
    return ".".join(path.split(".")[:-1] + ["ti", "pt"])


def _collect_embeds(text_encoder, placeholder_tokens, placeholder_token_ids):
    table = text_encoder.get_input_embeddings().weight
    got = {}
    for tok, tok_id in zip(placeholder_tokens, placeholder_token_ids):
        vec = table[tok_id]
        print(f"Current Learned Embeddings for {tok}:, id {tok_id} ", vec[:4])
        got[tok] = vec.detach().cpu()
    return got


def save_all(
    unet,
    text_encoder,
    save_path,
    placeholder_token_ids=None,
    placeholder_tokens=None,
    save_lora=True,
    save_ti=True,
    target_replace_module_text=TEXT_ENCODER_DEFAULT_TARGET_REPLACE,
    target_replace_module_unet=DEFAULT_TARGET_REPLACE,
    safe_form=True,
):
    """lora.py:1045-1110: either {.pt, .text_encoder.pt, .ti.pt} or one .safetensors."""
    if safe_form:
        assert save_path.endswith(".safetensors"), f"Save path : {save_path} should end with .safetensors"
        loras, embeds = {}, {}
        if save_lora:
            loras["unet"] = (unet, target_replace_module_unet)
            loras["text_encoder"] = (text_encoder, target_replace_module_text)
        if save_ti:
            embeds = _collect_embeds(text_encoder, placeholder_tokens, placeholder_token_ids)
        save_safeloras_with_embeds(loras, embeds, save_path)
        return
    if save_ti:
        ti_path = _ti_lora_path(save_path)
        torch.save(_collect_embeds(text_encoder, placeholder_tokens, placeholder_token_ids), ti_path)
        print("Ti saved to ", ti_path)
    if save_lora:
        save_lora_weight(unet, save_path, target_replace_module=target_replace_module_unet)
        print("Unet saved to ", save_path)
        save_lora_weight(text_encoder, _text_lora_path(save_path),
                         target_replace_module=target_replace_module_text)
        print("Text Encoder saved to ", _text_lora_path(save_path))
