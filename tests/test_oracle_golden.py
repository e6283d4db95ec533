"""Pin the oracle (oracle/*.py) against golden vectors produced by RUNNING the real reference
(/root/reference/lora_diffusion/lora.py, cli_svd.py; scripts/make_golden.py). CPU only.

Tolerance: the reference ran in fp32 (torch eager CPU), the oracle in float64: 2e-5 relative."""
import os

import pytest
import torch
import torch.nn as nn

from oracle import lora_ops as O
from oracle import svd_ref
from oracle.ref_modules import RefLoraSite

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def test_linear_forward_backward_vs_reference_golden():
    cases = torch.load(f"{GOLD}/ops_linear.pt")
    assert len(cases) >= 4
    for c in cases:
        x = c["x"]
        M = x.shape[0] * x.shape[1]
        x2, gy2 = x.reshape(M, -1), c["gy"].reshape(M, -1)
        y = O.lora_linear_forward(x2, c["W"], c["b"], c["A"], c["B"], c["scale"], diag=c["diag"])
        assert rel(y, c["y"].reshape(M, -1)) < 2e-5
        dX, dA, dB = O.lora_linear_backward(gy2, x2, c["W"], c["A"], c["B"], c["scale"], diag=c["diag"])
        assert rel(dX, c["dX"].reshape(M, -1)) < 2e-5
        assert rel(dA, c["dA"]) < 2e-5
        assert rel(dB, c["dB"]) < 2e-5


def test_conv_forward_backward_vs_reference_golden():
    cases = torch.load(f"{GOLD}/ops_conv.pt")
    for c in cases:
        y = O.lora_conv2d_forward(c["x"], c["W"], c["b"], c["A"], c["B"], c["scale"], padding=c["padding"])
        assert rel(y, c["y"]) < 2e-5
        dX, dA, dB = O.lora_conv2d_backward(c["gy"], c["x"], c["W"], c["A"], c["B"], c["scale"],
                                            padding=c["padding"])
        assert rel(dX, c["dX"]) < 2e-5
        assert rel(dA, c["dA"]) < 2e-5
        assert rel(dB, c["dB"]) < 2e-5


def test_dropout_is_identity_in_eval_mode_golden():
    c = torch.load(f"{GOLD}/ops_dropout_eval.pt")
    y = O.lora_linear_forward(c["x"], c["W"], None, c["A"], c["B"], 1.0)
    assert rel(y, c["y_eval"]) < 2e-5


def test_dropout_mask_semantics():
    """keep-mask m: branch * m / (1-p) (nn.Dropout, lora.py:45,56); gradients see the same mask."""
    torch.manual_seed(0)
    x, W, A, B = torch.randn(6, 8), torch.randn(5, 8), torch.randn(2, 8), torch.randn(5, 2)
    mask = (torch.rand(6, 5) > 0.3).double()
    y = O.lora_linear_forward(x, W, None, A, B, 0.5, keep_mask=mask, dropout_p=0.3)
    base = x.double() @ W.double().T
    assert torch.allclose(y - base, (x.double() @ A.double().T @ B.double().T) * mask / 0.7 * 0.5)
    xa = x.clone().requires_grad_(True)
    Aa, Ba = A.clone().double().requires_grad_(True), B.clone().double().requires_grad_(True)
    ya = xa.double() @ W.double().T + ((xa.double() @ Aa.T) @ Ba.T) * mask / 0.7 * 0.5
    gy = torch.randn(6, 5).double()
    ya.backward(gy)
    dX, dA, dB = O.lora_linear_backward(gy, x, W, A, B, 0.5, keep_mask=mask, dropout_p=0.3)
    assert rel(dX, xa.grad) < 1e-6 and rel(dA, Aa.grad) < 1e-12 and rel(dB, Ba.grad) < 1e-12


def test_ref_modules_match_reference_golden():
    """oracle/ref_modules.RefLoraSite (torch eager restatement) == reference outputs and grads."""
    for c in torch.load(f"{GOLD}/ops_linear.pt"):
        N, K = c["W"].shape
        base = nn.Linear(K, N, bias=c["b"] is not None)
        base.weight.data.copy_(c["W"])
        if c["b"] is not None:
            base.bias.data.copy_(c["b"])
        base.requires_grad_(False)
        s = RefLoraSite(base, r=c["A"].shape[0], dropout_p=0.0, scale=c["scale"])
        s.down.data.copy_(c["A"]); s.up.data.copy_(c["B"]); s.diag = c["diag"]
        x = c["x"].clone().requires_grad_(True)
        y = s(x)
        y.backward(c["gy"])
        assert rel(y, c["y"]) < 1e-6 and rel(x.grad, c["dX"]) < 1e-5
        assert rel(s.down.grad, c["dA"]) < 1e-5 and rel(s.up.grad, c["dB"]) < 1e-5
    for c in torch.load(f"{GOLD}/ops_conv.pt"):
        Cout, Cin, k, _ = c["W"].shape
        base = nn.Conv2d(Cin, Cout, k, 1, c["padding"], bias=c["b"] is not None)
        base.weight.data.copy_(c["W"])
        if c["b"] is not None:
            base.bias.data.copy_(c["b"])
        base.requires_grad_(False)
        s = RefLoraSite(base, r=c["A"].shape[0], dropout_p=0.0, scale=c["scale"])
        s.down.data.copy_(c["A"]); s.up.data.copy_(c["B"])
        x = c["x"].clone().requires_grad_(True)
        y = s(x)
        y.backward(c["gy"])
        assert rel(y, c["y"]) < 1e-6 and rel(x.grad, c["dX"]) < 1e-5
        assert rel(s.down.grad, c["dA"]) < 1e-5 and rel(s.up.grad, c["dB"]) < 1e-5


def test_clip_adamw_oracle_vs_torch_golden():
    """torch.optim.AdamW + clip_grad_norm_ (the calls the reference makes) trajectory."""
    d = torch.load(f"{GOLD}/adamw_clip.pt")
    p = [x.clone() for x in d["p0"]]
    m = [torch.zeros_like(x) for x in p]
    v = [torch.zeros_like(x) for x in p]
    for step, t in enumerate(d["traj"], start=1):
        p, m, v, total = O.clip_adamw_step(p, t["grads"], m, v, step, d["lrs"])
        assert abs(total - t["total_norm"]) < 1e-5 * t["total_norm"]
        for a, b in zip(p, t["params"]):
            assert rel(a, b) < 2e-6


def test_clip_adamw_inv_world_equals_mean_of_rank_grads():
    torch.manual_seed(1)
    p = [torch.randn(5, 3)]
    g0, g1 = torch.randn(5, 3), torch.randn(5, 3)
    a = O.clip_adamw_step(p, [g0 + g1], [torch.zeros(5, 3)], [torch.zeros(5, 3)], 1, [1e-3], inv_world=0.5)
    b = O.clip_adamw_step(p, [(g0 + g1) / 2], [torch.zeros(5, 3)], [torch.zeros(5, 3)], 1, [1e-3])
    assert rel(a[0][0], b[0][0]) < 1e-14 and abs(a[3] - b[3]) < 1e-12


def test_svd_oracle_vs_reference_golden():
    """cli_svd.overwrite_base outputs. Same LAPACK here as when the golden was made, so factors
    match elementwise up to fp32 noise; the sign-invariant quantities are checked as well."""
    for c in torch.load(f"{GOLD}/svd_distill.pt"):
        up, down, S, hi = svd_ref.svd_distill_pair(c["Wb"], c["Wt"], c["rank"], c["q"])
        assert up.shape == c["up"].shape and down.shape == c["down"].shape
        prod = up.flatten(1) @ down.flatten(1)
        prod_ref = c["up"].flatten(1) @ c["down"].flatten(1)
        assert rel(prod, prod_ref) < 1e-4
        # clamp rule: nothing exceeds hi, and (same LAPACK build) the factors agree elementwise
        assert float(max(up.abs().max(), down.abs().max())) <= hi * (1 + 1e-6)
        assert float(max(c["up"].abs().max(), c["down"].abs().max())) <= hi * (1 + 1e-4)
        assert rel(up, c["up"]) < 1e-3 and rel(down, c["down"]) < 1e-3


def test_step_loss_matches_reference_loss_step_golden():
    """oracle/ref_step.py::RefDreamboothStep.forward_loss against losses the reference's own
    `loss_step` (cli_lora_pti.py:260-370) produced on the same tiny models and inputs (golden
    written by scripts/make_golden.py::gen_loss_step): plain, t_mutliplier, masked loss with two
    temperatures, the 9-channel inpainting input, inpainting + mask. CPU fp32."""
    import os
    from lora_b200.host.clip import build_text_encoder
    from lora_b200.host.ddpm import DDPMNoiser
    from lora_b200.host.unet_sd15 import UNet2DConditionModel, UNetConfig
    from oracle.ref_modules import ref_inject
    from oracle.ref_step import RefDreamboothStep
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "pti_loss_step.pt"))
    lat, ids = G["latents"], G["input_ids"]
    assert len(G["cases"]) == 6
    for case in G["cases"]:
        torch.manual_seed(G["model_seed"])
        cfg = UNetConfig.tiny()
        cfg.in_channels = case["in_channels"]
        unet = UNet2DConditionModel(cfg)
        text = build_text_encoder(tiny=True)
        sites = ref_inject(unet, {"CrossAttention", "Attention", "GEGLU"}, r=4)
        gg = torch.Generator().manual_seed(G["up_seed"])
        for s in sites:
            s.down.data.normal_(0, 0.25, generator=gg)
            s.up.data.normal_(0, 0.05, generator=gg)
        unet.train(False), text.train(False)
        kw = case["kwargs"]
        ref = RefDreamboothStep(unet, text, DDPMNoiser(device="cpu"), sites, None,
                                t_multiplier=kw.get("t_mutliplier", 1.0))
        unet.train(False), text.train(False)
        torch.manual_seed(G["step_seed"])
        noise = torch.randn_like(lat)                       # the reference's draw order: noise, then t
        t = torch.randint(0, int(1000 * ref.t_multiplier), (lat.shape[0],)).long()
        loss = ref.forward_loss(
            lat, ids, noise, t,
            loss_mask=G["loss_mask"] if case["with_mask"] else None,
            mask_temperature=kw.get("mask_temperature", 1.0),
            inpaint=(G["inpaint_mask"], G["masked_latents"]) if kw.get("train_inpainting") else None)
        assert abs(float(loss) - case["loss"]) <= 2e-6 * abs(case["loss"]), (case["name"], float(loss), case["loss"])


def test_ti_oracle_matches_reference_train_inversion_golden():
    """oracle/ti_ref.py::ti_table_step replayed over the gradients the reference's own
    `train_inversion` loop (cli_lora_pti.py:373-542) saw for 3 steps (golden from
    scripts/make_golden.py::gen_ti: per-step placeholder-row gradients, the scheduled lr, rows and
    table sum afterwards): trained rows equal after every step, every other row restored."""
    import os
    from lora_b200.host.clip import build_text_encoder
    from lora_b200.host.unet_sd15 import UNet2DConditionModel, UNetConfig
    from oracle.ti_ref import ti_table_step
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "ti_train_inversion.pt"))
    torch.manual_seed(G["model_seed"])
    UNet2DConditionModel(UNetConfig.tiny())              # consumes the RNG exactly as the generator did
    te = build_text_encoder(tiny=True)
    table0 = te.get_input_embeddings().weight.detach().clone()
    tok = G["token_ids"]
    assert torch.equal(table0[tok], G["table0_rows"]) and abs(float(table0.double().sum()) - G["table0_sum"]) < 1e-9
    V, D = table0.shape
    upd = torch.zeros(V, dtype=torch.bool)
    upd[tok] = True
    table, m, v = table0.clone(), torch.zeros(V, D), torch.zeros(V, D)
    for k, st in enumerate(G["steps"], start=1):
        dense = torch.zeros(V, D)
        dense[tok] = st["grad_rows"]
        table, m, v = ti_table_step(table, dense, m, v, k, st["lr"], upd, table0,
                                    weight_decay=G["weight_decay"], clip_ti_decay=True)
        want = st["rows_after"].double()
        assert float((table[tok] - want).norm() / want.norm()) < 2e-6, k
        assert torch.equal(table[~upd], table0[~upd].double())
        assert abs(float(table.sum()) - st["table_sum"]) < 1e-4
    assert G["steps"][0]["lr"] != G["steps"][2]["lr"]       # the schedule really moved the lr / lambda


def test_oracle_step_matches_reference_perform_tuning_golden():
    """oracle/ref_step.py::RefDreamboothStep.step (noise, t < 0.8*T, masked loss, backward,
    clip_grad_norm_(1.0) over all parameters, two-group AdamW) against 3 iterations of the
    reference's real `perform_tuning` loop (cli_lora_pti.py:545-680; golden from
    scripts/make_golden.py::gen_tuning): per-step loss and every LoRA factor afterwards."""
    import os
    from lora_b200.host.clip import build_text_encoder
    from lora_b200.host.ddpm import DDPMNoiser
    from lora_b200.host.unet_sd15 import UNet2DConditionModel, UNetConfig
    from oracle.ref_modules import ref_inject
    from oracle.ref_step import RefDreamboothStep
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "pti_perform_tuning.pt"))
    torch.manual_seed(G["model_seed"])
    unet = UNet2DConditionModel(UNetConfig.tiny())
    text = build_text_encoder(tiny=True)
    us = ref_inject(unet, {"CrossAttention", "Attention", "GEGLU"}, r=4)
    ts = ref_inject(text, {"CLIPAttention"}, r=4)
    assert len(us) == G["n_unet_sites"] and len(us) + len(ts) == len(G["factor_moments"])
    gg = torch.Generator().manual_seed(G["factor_seed"])
    for s in us + ts:
        s.down.data.normal_(0, 0.25, generator=gg)
        s.up.data.normal_(0, 0.05, generator=gg)
    ref = RefDreamboothStep(unet, text, DDPMNoiser(device="cpu"), us, ts, lr=1e-4, lr_text=1e-5,
                            weight_decay=1e-3, t_multiplier=0.8)
    torch.manual_seed(G["step_seed"])
    for st, batch in zip(G["steps"], G["batches"]):
        for grp, lr in zip(ref.opt.param_groups, st["lrs"]):
            grp["lr"] = lr                                   # lr_scheduler.step() precedes the update
        loss = ref.step(batch["pixel_values"], batch["input_ids"], loss_mask=batch["mask"].float(),
                        mask_temperature=st["kwargs"]["mask_temperature"])
        assert abs(float(loss) - st["loss"]) <= 5e-6 * abs(st["loss"]), (float(loss), st["loss"])
    sites = us + ts
    for i, (up, down) in G["factors"].items():
        assert torch.allclose(sites[i].up.data, up, rtol=0, atol=2e-7), i
        assert torch.allclose(sites[i].down.data, down, rtol=0, atol=2e-7), i
    for s, (su, sd, qu, qd) in zip(sites, G["factor_moments"]):
        assert abs(float(s.up.data.double().sum()) - su) < 1e-5 and abs(float(s.down.data.double().sum()) - sd) < 1e-5
        assert abs(float((s.up.data.double() ** 2).sum()) - qu) < 1e-5 * max(1.0, qu)
        assert abs(float((s.down.data.double() ** 2).sum()) - qd) < 1e-5 * max(1.0, qd)
