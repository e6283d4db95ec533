"""Flat namespace equal to `lora_diffusion.lora` of the reference (lora_diffusion/__init__.py:1
does `from .lora import *`): code written against the reference can swap the import only."""
from .inject import (DEFAULT_TARGET_REPLACE, EMBED_FLAG, TEXT_ENCODER_DEFAULT_TARGET_REPLACE,  # noqa: F401
                     TEXT_ENCODER_EXTENDED_TARGET_REPLACE, UNET_DEFAULT_TARGET_REPLACE,
                     UNET_EXTENDED_TARGET_REPLACE, _find_children, _find_modules, _find_modules_v2,
                     extract_lora_as_tensor, extract_lora_ups_down, inject_trainable_lora,
                     inject_trainable_lora_extended)
from .modules import LoraInjectedConv2d, LoraInjectedLinear  # noqa: F401
from .patch import (apply_learned_embed_in_clip, collapse_lora, inspect_lora,  # noqa: F401
                    load_learned_embed_in_clip, monkeypatch_add_lora, monkeypatch_or_replace_lora,
                    monkeypatch_or_replace_lora_extended, monkeypatch_or_replace_safeloras,
                    monkeypatch_remove_lora, patch_pipe, set_lora_diag, tune_lora_scale)
from .persist import (_text_lora_path, _ti_lora_path, convert_loras_to_safeloras,  # noqa: F401
                      convert_loras_to_safeloras_with_embeds, load_safeloras, load_safeloras_both,
                      load_safeloras_embeds, parse_safeloras, parse_safeloras_embeds, safe_open,
                      safe_save, safetensors_available, save_all, save_lora_as_json,
                      save_lora_weight, save_safeloras, save_safeloras_with_embeds)
