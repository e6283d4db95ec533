"""Step-level parity at a BASELINE.json config size (configs[1]: SD1.5-shaped UNet + text encoder,
rank 4, 512x512, bs 1) -- the whole training step of lora_b200.train.LoraTrainStep, eager AND
CUDA-graph replayed, against the oracle's restated reference step (oracle/ref_step.py: torch eager
+ torch.optim.AdamW + clip_grad_norm_, pinned to the reference's real perform_tuning loop) on the
same device with the SAME noise / timesteps (StepConfig.external_noise).

Tolerances (north_star): bf16 autocast on both sides -> loss within 1e-2 relative, first-step LoRA
gradient within 3e-2 relative; fp32-faithful mode (set_fp32_mode("split"), TF32 off) on fp32
models -> |loss - reference loss| <= 1e-4."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _sd15_pair(seed=0, r=4, tiny=False):
    """(ours unet, ours text, ref unet, ref text, ref unet sites, ref text sites), fp32 weights,
    identical frozen weights and identical non-zero LoRA factors on both sides."""
    import lora_b200 as L
    from lora_b200.host.clip import build_text_encoder
    from lora_b200.host.unet_sd15 import UNet2DConditionModel, UNetConfig
    from oracle.ref_modules import ref_inject
    torch.manual_seed(seed)
    with torch.device(DEV):
        unet = UNet2DConditionModel(UNetConfig.tiny() if tiny else UNetConfig.sd15())
        text = build_text_encoder(tiny=tiny)
    unet.requires_grad_(False)
    text.requires_grad_(False)
    unet = unet.to(memory_format=torch.channels_last)
    unet_r, text_r = copy.deepcopy(unet), copy.deepcopy(text)
    L.inject_trainable_lora(unet, r=r)
    L.inject_trainable_lora(text, target_replace_module={"CLIPAttention"}, r=r)
    us = ref_inject(unet_r, {"CrossAttention", "Attention", "GEGLU"}, r=r)
    ts = ref_inject(text_r, {"CLIPAttention"}, r=r)
    ours = [m for m in list(unet.modules()) + list(text.modules()) if type(m).__name__ == "LoraInjectedLinear"]
    assert len(ours) == len(us) + len(ts)
    g = torch.Generator(device=DEV).manual_seed(3)
    for o, rs in zip(ours, us + ts):
        o.lora_up.weight.data.normal_(0, 0.02, generator=g)
        rs.up.data.copy_(o.lora_up.weight.data)
        rs.down.data.copy_(o.lora_down.weight.data)
    return unet, text, unet_r, text_r, us, ts


def _draw(step, shape):
    g = torch.Generator(device=DEV).manual_seed(1000 + step)
    noise = torch.randn(shape, device=DEV, generator=g)
    t = torch.randint(0, 1000, (shape[0],), device=DEV, generator=g).long()
    return noise, t


@pytest.mark.parametrize("tiny", [False])
def test_sd15_512_bf16_step_eager_and_graph_match_reference_step(tiny):
    import lora_b200 as L
    from lora_b200.host.ddpm import DDPMNoiser
    from lora_b200.train import LoraTrainStep, StepConfig
    from oracle.ref_step import RefDreamboothStep
    unet, text, unet_r, text_r, us, ts = _sd15_pair(tiny=tiny)
    assert len(us) == (144 if not tiny else len(us)) and len(ts) == (48 if not tiny else len(ts))
    Ls = 16 if tiny else 64
    shape = (1, 4, Ls, Ls)
    L.set_grouping(True)                      # the benched configuration
    try:
        cfg = StepConfig(use_cuda_graph=True, graph_warmup=2, autocast_dtype=torch.bfloat16, external_noise=True)
        tr = LoraTrainStep(unet, text, cfg, latent_shape=shape, seq_len=77, device=DEV)
        ref = RefDreamboothStep(unet_r, text_r, DDPMNoiser(device=DEV), us, ts, autocast_dtype=torch.bfloat16)
        torch.manual_seed(7)
        lat = torch.randn(shape, device=DEV) * 0.18215
        ids = torch.randint(0, text.config.vocab_size, (1, 77), device=DEV)
        tr.latents.copy_(lat)
        tr.input_ids.copy_(ids)

        # ---- first-step gradients (no optimizer step on either side)
        noise, t = _draw(0, shape)
        loss_r = ref.forward_loss(lat, ids, noise, t)
        loss_r.backward()
        g_ref = torch.cat([p.grad.flatten() for p in ref.unet_params + ref.text_params]).clone()
        ref.opt.zero_grad()
        tr.noise.copy_(noise); tr.timesteps.copy_(t)
        tr._fwd_bwd()
        g_ours = torch.cat([p.grad.flatten() for p in tr.arena.parameters()]).clone()
        assert abs(float(tr.loss) - float(loss_r)) < 1e-2 * abs(float(loss_r))
        assert rel(g_ours, g_ref) < 3e-2
        tr.arena.zero_grad()

        p0 = torch.cat([p.detach().flatten() for p in tr.arena.parameters()]).clone()
        # ---- one eager step, then capture (must not disturb the trajectory), then 3 graph replays
        traj = []
        for step in range(4):
            noise, t = _draw(step, shape)
            l_ref = float(ref.step(lat, ids, noise=noise, timesteps=t))
            tr.noise.copy_(noise); tr.timesteps.copy_(t)
            if step == 0:
                tr._body()
                l_ours = float(tr.loss)
                tr.prepare()
                assert tr.graph is not None, tr.graph_error
                assert int(tr.arena.step_dev) == 1          # warm-up steps were rolled back
            else:
                l_ours = float(tr.step_device())
            traj.append((l_ours, l_ref))
            assert abs(l_ours - l_ref) < 1e-2 * abs(l_ref), (step, traj)
        assert int(tr.arena.step_dev) == 4
        # the LoRA factors after 4 optimizer steps: both sides moved the same way. Adam's first
        # steps are sign-like (+-lr per element), so elements whose bf16-noisy gradient is ~0 flip;
        # uncorrelated updates would give 1.41 here.
        p_ref = torch.cat([p.detach().flatten().float() for p in ref.unet_params + ref.text_params])
        p_ours = torch.cat([p.detach().flatten() for p in tr.arena.parameters()])
        assert rel(p_ours - p0, p_ref - p0) < 0.5
    finally:
        L.set_grouping(False)


def test_sd15_512_fp32_split_mode_loss_within_1e4():
    import lora_b200 as L
    from lora_b200.host.ddpm import DDPMNoiser
    from lora_b200.train import LoraTrainStep, StepConfig
    from oracle.ref_step import RefDreamboothStep
    old = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    L.set_fp32_mode("split")
    try:
        unet, text, unet_r, text_r, us, ts = _sd15_pair(seed=1)
        shape = (1, 4, 64, 64)
        cfg = StepConfig(use_cuda_graph=False, autocast_dtype=None, external_noise=True)
        tr = LoraTrainStep(unet, text, cfg, latent_shape=shape, seq_len=77, device=DEV)
        ref = RefDreamboothStep(unet_r, text_r, DDPMNoiser(device=DEV), us, ts)
        torch.manual_seed(8)
        lat = torch.randn(shape, device=DEV) * 0.18215
        ids = torch.randint(0, text.config.vocab_size, (1, 77), device=DEV)
        tr.latents.copy_(lat)
        tr.input_ids.copy_(ids)
        for step in range(2):
            noise, t = _draw(10 + step, shape)
            l_ref = float(ref.step(lat, ids, noise=noise, timesteps=t))
            tr.noise.copy_(noise); tr.timesteps.copy_(t)
            l_ours = float(tr.step_device())
            assert abs(l_ours - l_ref) <= 1e-4, (step, l_ours, l_ref)
    finally:
        L.set_fp32_mode("bf16")
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
