"""`.pt` LoRA / textual-inversion files -> one `.safetensors` LoRA file.

Mirrors `lora_diffusion/cli_pt_to_safetensors.py:19-77` of the reference.
A `.pt` that unpickles to a dict is a token -> embedding table; anything else is a LoRA pair list
[up, down, up, down, ...] whose model name is the second-to-last dotted part of the file name
(`x.text_encoder.pt` -> text_encoder, `x.pt` -> unet). Per-model overrides come as keyword
arguments `"<name>.rank"`, `"<name>.target_modules"`.
"""
import os

import torch

from .inject import (DEFAULT_TARGET_REPLACE, TEXT_ENCODER_DEFAULT_TARGET_REPLACE,
                     UNET_DEFAULT_TARGET_REPLACE)
from .persist import convert_loras_to_safeloras_with_embeds

_TARGETS = {"unet": UNET_DEFAULT_TARGET_REPLACE, "text_encoder": TEXT_ENCODER_DEFAULT_TARGET_REPLACE}


def model_name_of(path: str) -> str:
    parts = os.path.basename(path).split(".")
    return parts[-2] if len(parts) > 2 else "unet"


def convert(*paths, outpath, overwrite=False, **settings):
    if os.path.exists(outpath) and not overwrite:
        raise ValueError(f"Output path {outpath} already exists, and overwrite is not True")
    modelmap, embeds = {}, {}
    for path in paths:
        data = torch.load(path)
        if isinstance(data, dict):
            print(f"Loading textual inversion embeds {data.keys()} from {path}")
            embeds.update(data)
            continue
        name = model_name_of(path)
        cfg = {"target_modules": _TARGETS.get(name, DEFAULT_TARGET_REPLACE), "rank": 4}
        cfg.update({k[len(name) + 1:]: v for k, v in settings.items() if k.startswith(name + ".")})
        print(f"Loading Lora for {name} from {path} with settings {cfg}")
        modelmap[name] = (path, cfg["target_modules"], cfg["rank"])
    convert_loras_to_safeloras_with_embeds(modelmap, embeds, outpath)
