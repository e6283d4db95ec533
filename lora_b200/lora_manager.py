"""Import-path compatibility with `lora_diffusion.lora_manager`; the code lives in `join.py`."""
from .join import DummySafeTensorObject, LoRAManager, lora_join  # noqa: F401
