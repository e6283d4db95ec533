"""Textual-inversion step (lb_ti_embed_step + TextualInversionRows) vs the oracle's restatement of
cli_lora_pti.py:446-479 on the whole table, GPU."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_ti_rows_step_matches_full_table_reference():
    from lora_b200.host.clip import build_text_encoder
    from lora_b200.ti import TextualInversionRows
    from oracle.ti_ref import ti_table_step
    torch.manual_seed(0)
    te = build_text_encoder(tiny=True).to(DEV)
    te.requires_grad_(False)
    emb = te.get_input_embeddings()
    V, D = emb.weight.shape
    tok = [V - 2, V - 1]
    orig = emb.weight.detach().clone()
    ti = TextualInversionRows(te, tok, lr=5e-3)
    index_updates = torch.zeros(V, dtype=torch.bool)
    index_updates[tok] = True
    o_table, o_m, o_v = orig.clone(), torch.zeros(V, D), torch.zeros(V, D)
    ids = torch.randint(0, V - 2, (2, 77), device=DEV)
    ids[0, 5], ids[1, 9], ids[1, 10] = tok[0], tok[1], tok[0]
    for step in range(1, 4):
        out = te(ids)[0]
        loss = (out.float() ** 2).mean() * 10
        loss.backward()
        # dense gradient of the same loss w.r.t. the full table (what the reference's autograd gives)
        dense = torch.zeros(V, D)
        dense[tok] = ti.rows.grad.detach().cpu()
        o_table, o_m, o_v = ti_table_step(o_table, dense, o_m, o_v, step, 5e-3, index_updates, orig)
        ti.step()
        torch.cuda.synchronize()
        got = emb.weight.detach().double().cpu()
        assert float((got[tok] - o_table[tok]).norm() / o_table[tok].norm()) < 1e-5
        assert torch.equal(emb.weight.detach()[: V - 2], orig[: V - 2])          # untouched rows stay bit-identical
        assert float(ti.rows.grad.abs().max()) == 0.0
        assert torch.allclose(ti.rows.detach(), emb.weight.detach()[tok].float(), atol=1e-6)
    assert int(ti.step_dev) == 3


def test_ti_loop_through_the_step_engine_matches_oracle():
    """The phase-1 LOOP (lora_b200.train.TextualInversionStep: scheduler first, loss_step body with
    the UNet in eval mode, rows-only AdamW + norm decay) against the oracle's full-table update
    (oracle/ti_ref.py, pinned to the reference's real train_inversion) fed the dense gradient of the
    SAME loss computed by plain autograd on an untouched copy of the models."""
    import copy
    from lora_b200.host.clip import build_text_encoder
    from lora_b200.host.ddpm import DDPMNoiser
    from lora_b200.host.unet_sd15 import UNet2DConditionModel, UNetConfig
    from lora_b200.train import StepConfig, TextualInversionStep
    from oracle.ti_ref import ti_table_step
    torch.manual_seed(0)
    unet = UNet2DConditionModel(UNetConfig.tiny()).to(DEV)
    te = build_text_encoder(tiny=True).to(DEV)
    unet.requires_grad_(False); te.requires_grad_(False)
    unet_r, te_r = copy.deepcopy(unet), copy.deepcopy(te)
    emb_r = te_r.get_input_embeddings()
    emb_r.weight.requires_grad_(True)
    V, D = emb_r.weight.shape
    tok = [V - 3, V - 1]
    index_updates = torch.zeros(V, dtype=torch.bool)
    index_updates[tok] = True
    orig = emb_r.weight.detach().clone()
    cfg = StepConfig(use_cuda_graph=False, external_noise=True, lr_scheduler="linear", lr_warmup_steps=0,
                     max_train_steps=10, use_mask=True)
    lr0 = 5e-3
    eng = TextualInversionStep(unet, te, tok, cfg, lr=lr0, latent_shape=(2, 4, 16, 16), device=DEV)
    noiser = DDPMNoiser(device=DEV)
    lat = torch.randn(2, 4, 16, 16, device=DEV) * 0.18215
    ids = torch.randint(0, V - 3, (2, 77), device=DEV)
    ids[0, 3], ids[1, 7] = tok[0], tok[1]
    mask = (torch.rand(2, 1, 16, 16, device=DEV) > 0.5).float()
    eng.latents.copy_(lat); eng.input_ids.copy_(ids); eng.mask.copy_(mask)
    o_table, o_m, o_v = orig.clone(), torch.zeros(V, D), torch.zeros(V, D)
    unet_r.eval(); te_r.train()
    for step in range(1, 4):
        g = torch.Generator(device=DEV).manual_seed(step)
        noise = torch.randn(2, 4, 16, 16, device=DEV, generator=g)
        t = torch.randint(0, 1000, (2,), device=DEV, generator=g)
        # reference side: plain autograd w.r.t. the whole table
        with torch.no_grad():
            emb_r.weight.copy_(o_table.to(DEV, torch.float32))
        emb_r.weight.grad = None
        noisy = noiser.add_noise(lat, noise, t)
        pred = unet_r(noisy, t, te_r(ids)[0]).sample
        m = (mask + 0.01).pow(1.0)
        m = m / m.max()
        loss_r = torch.nn.functional.mse_loss((pred * m).float(), (noise * m).float(), reduction="none").mean([1, 2, 3]).mean()
        loss_r.backward()
        lr_k = lr0 * max(0.0, (10 - step) / 10)          # linear schedule, stepped first
        o_table, o_m, o_v = ti_table_step(o_table, emb_r.weight.grad, o_m, o_v, step, lr_k, index_updates, orig)
        # our loop
        eng.noise.copy_(noise); eng.timesteps.copy_(t)
        l_ours = float(eng.step_device())
        assert abs(l_ours - float(loss_r)) <= 1e-4 * abs(float(loss_r)) + 1e-6, (step, l_ours, float(loss_r))
        got = te.get_input_embeddings().weight.detach().double().cpu()
        # fp32 autograd on both sides; Adam's first steps amplify gradient noise (sign-like update)
        assert float((got[tok] - o_table[tok]).norm() / o_table[tok].norm()) < 1e-3
        assert torch.equal(te.get_input_embeddings().weight.detach()[: V - 3], orig[: V - 3].to(DEV))
    eng.release()
