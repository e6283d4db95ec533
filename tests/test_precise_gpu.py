"""fp32-faithful mode (split-bf16 operands, lora_b200.set_fp32_mode("split")), GPU.

This is the configuration behind the north star's "FP32 loss within 1e-4 of the reference for
the same seed": fp32 frozen weights, fp32 activations, no autocast (the reference's
mixed_precision="no", BASELINE.json configs[0])."""
import copy

import pytest
import torch
import torch.nn as nn

from oracle import lora_ops as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(autouse=True)
def _split_mode():
    import lora_b200 as L
    L.set_fp32_mode("split")
    yield
    L.set_fp32_mode("bf16")


@pytest.mark.parametrize("M,K,N,r,bias", [(512, 320, 320, 4, False), (77, 768, 1280, 8, True), (130, 72, 88, 3, True),
                                          (1024, 640, 2560, 16, True)])
def test_split_mode_operator_is_fp32_faithful(M, K, N, r, bias):
    """fp32 module, fp32 input: forward/backward vs the float64 oracle on the UNROUNDED fp32
    operands. Base GEMM error ~2^-16; the LoRA branch keeps one bf16 rounding of T' and of the up
    tile => bound: 2^-7 of the branch + 1e-4 of the output (branch is ~10 % of the output here)."""
    import lora_b200 as L
    torch.manual_seed(M + N)
    m = L.LoraInjectedLinear(K, N, bias=bias, r=r, dropout_p=0.0, scale=0.7).to(DEV)
    m.linear.requires_grad_(False)
    m.lora_up.weight.data.normal_(0, 0.02)
    x = torch.randn(M, K, device=DEV, requires_grad=True)
    y = m(x)
    assert y.dtype == torch.float32
    gy = torch.randn(M, N, device=DEV)
    y.backward(gy)
    W, b = m.linear.weight.detach(), (m.linear.bias.detach() if bias else None)
    A, B = m.lora_down.weight.detach(), m.lora_up.weight.detach()
    ref = O.lora_linear_forward(x, W, b, A, B, 0.7)
    base = O.lora_linear_forward(x, W, b, A, torch.zeros_like(B), 0.0)
    branch = float((ref - base).norm())
    assert float((y.double().cpu() - ref).norm()) <= 2 ** -7 * branch + 1e-4 * float(ref.norm())
    # the frozen part alone is fp32-grade
    L.tune_lora_scale(nn.Sequential(m), 0.0)
    y0 = m(x.detach())
    assert rel(y0, base) < 5e-5
    dX, dA, dB = O.lora_linear_backward(gy, x, W, A, B, 0.7)
    assert rel(x.grad, dX) < 2e-3
    assert rel(m.lora_down.weight.grad, dA) < 2e-3      # dTs from bf16-rounded B^T terms is 3-term split too
    assert rel(m.lora_up.weight.grad, dB) < 1e-4


def test_fp32_training_loss_within_1e4_of_reference_step():
    """Whole Dreambooth step in fp32 (no autocast) on the tiny host models, same seeds: our loss vs
    the oracle's reference step (torch eager fp32 + torch.optim.AdamW), 4 steps: |dLoss| <= 1e-4."""
    import lora_b200 as L
    from lora_b200.host.clip import build_text_encoder
    from lora_b200.host.ddpm import DDPMNoiser
    from lora_b200.host.unet_sd15 import UNet2DConditionModel, UNetConfig
    from lora_b200.train import LoraTrainStep, StepConfig
    from oracle.ref_modules import ref_inject
    from oracle.ref_step import RefDreamboothStep
    torch.manual_seed(0)
    unet = UNet2DConditionModel(UNetConfig.tiny()).to(DEV)
    text = build_text_encoder(tiny=True).to(DEV)
    unet.requires_grad_(False); text.requires_grad_(False)
    unet_r, text_r = copy.deepcopy(unet), copy.deepcopy(text)
    L.inject_trainable_lora(unet, r=4)
    L.inject_trainable_lora(text, target_replace_module={"CLIPAttention"}, r=4)
    us = ref_inject(unet_r, {"CrossAttention", "Attention", "GEGLU"}, r=4)
    ts = ref_inject(text_r, {"CLIPAttention"}, r=4)
    ours = [m for m in list(unet.modules()) + list(text.modules()) if type(m).__name__ == "LoraInjectedLinear"]
    g = torch.Generator(device=DEV).manual_seed(3)
    for o, r in zip(ours, us + ts):
        o.lora_up.weight.data.normal_(0, 0.02, generator=g)
        r.up.data.copy_(o.lora_up.weight.data); r.down.data.copy_(o.lora_down.weight.data)
    old = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        cfg = StepConfig(use_cuda_graph=False, autocast_dtype=None)
        tr = LoraTrainStep(unet, text, cfg, latent_shape=(1, 4, 16, 16), seq_len=77, device=DEV)
        ref = RefDreamboothStep(unet_r, text_r, DDPMNoiser(device=DEV), us, ts)
        lat = torch.randn(1, 4, 16, 16, device=DEV) * 0.18215
        ids = torch.randint(0, 1000, (1, 77), device=DEV)
        tr.latents.copy_(lat); tr.input_ids.copy_(ids)
        for step in range(4):
            torch.manual_seed(500 + step)
            l_ref = float(ref.step(lat, ids))
            torch.manual_seed(500 + step)
            l_ours = float(tr.step_device())
            assert abs(l_ours - l_ref) <= 1e-4, (step, l_ours, l_ref)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
