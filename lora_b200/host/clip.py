"""SD1.5 text encoder host model: transformers' CLIPTextModel with the SD1.5 (CLIP ViT-L/14 text)
hyper-parameters, random-initialised (no weights offline). The reference injects LoRA into its
`CLIPAttention` blocks (lora_diffusion/lora.py:163; train_lora_dreambooth.py:608-613): 12 layers x
{k_proj, v_proj, q_proj, out_proj} = 48 sites of (77, 768 -> 768)."""
from transformers import CLIPTextConfig, CLIPTextModel


def sd15_text_config(tiny: bool = False) -> CLIPTextConfig:
    if tiny:
        return CLIPTextConfig(vocab_size=1000, hidden_size=48, intermediate_size=96,
                              num_hidden_layers=2, num_attention_heads=2,
                              max_position_embeddings=77, hidden_act="quick_gelu",
                              projection_dim=48)
    return CLIPTextConfig(vocab_size=49408, hidden_size=768, intermediate_size=3072,
                          num_hidden_layers=12, num_attention_heads=12,
                          max_position_embeddings=77, hidden_act="quick_gelu",
                          projection_dim=768)


def build_text_encoder(tiny: bool = False) -> CLIPTextModel:
    cfg = sd15_text_config(tiny)
    cfg._attn_implementation = "sdpa"
    return CLIPTextModel(cfg)
