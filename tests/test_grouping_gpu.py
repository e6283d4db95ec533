"""Grouped launches (lora_b200.set_grouping(True)): sites that receive the same input run as one
kernel launch; every site must still return exactly its own forward/backward (GPU)."""
import copy

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(autouse=True)
def _restore():
    import lora_b200 as L
    yield
    L.set_grouping(False)


def _block(seed=0):
    import lora_b200 as L
    from lora_b200.host.unet_sd15 import BasicTransformerBlock
    torch.manual_seed(seed)
    blk = BasicTransformerBlock(320, 8, 40, 768).to(DEV).to(torch.bfloat16)
    blk.requires_grad_(False)
    L.inject_trainable_lora(blk, r=4)
    g = torch.Generator(device=DEV).manual_seed(1)
    for m in blk.modules():
        if type(m).__name__ == "LoraInjectedLinear":
            m.lora_up.weight.data.normal_(0, 0.05, generator=g)
    return blk


def _run(blk, x, ctx, gy, passes):
    from lora_b200 import ops
    outs = []
    for _ in range(passes):
        for p in blk.parameters():
            p.grad = None
        xi = x.clone().requires_grad_(True)
        ops.LAUNCH_COUNT = 0
        y = blk(xi, ctx)
        y.backward(gy)
        torch.cuda.synchronize()
        grads = {n: p.grad.clone() for n, p in blk.named_parameters() if p.grad is not None}
        outs.append((y.detach().clone(), xi.grad.clone(), grads, ops.LAUNCH_COUNT))
    return outs


def test_grouped_block_equals_ungrouped_block():
    import lora_b200 as L
    x = torch.randn(1, 1024, 320, device=DEV, dtype=torch.bfloat16)
    ctx = torch.randn(1, 77, 768, device=DEV, dtype=torch.bfloat16)
    gy = torch.randn(1, 1024, 320, device=DEV, dtype=torch.bfloat16)
    L.set_grouping(False)
    refs = _run(_block(), x, ctx, gy, 3)
    ref = refs[0]
    L.set_grouping(True)
    got = _run(_block(), x, ctx, gy, 3)          # pass 1 learns the families, passes 2-3 use them
    assert got[0][3] == refs[0][3]               # learning pass: same launches as ungrouped
    # 9 sites: q,k,v (one launch), out, geglu, q2, (k2,v2 one launch), out2 -> 6 forward launches
    # instead of 9; the same for the dX launches and for the dA/dB reductions: 9 fewer per pass
    assert got[2][3] == refs[2][3] - 9, (got[2][3], refs[2][3])
    for y, dx, grads, _ in got:
        assert rel(y, ref[0]) < 2e-2 and rel(dx, ref[1]) < 2e-2     # bf16 outputs, different tile shapes
        assert set(grads) == set(ref[2])
        for n in grads:
            assert rel(grads[n], ref[2][n]) < 2e-2, n


def test_sibling_with_a_different_input_computes_alone():
    """Same parent, but the second site is fed a different tensor after the family was learned."""
    import lora_b200 as L
    from oracle import lora_ops as O

    class Attention(nn.Module):
        def __init__(self):
            super().__init__()
            self.to_q = nn.Linear(320, 320, bias=False)
            self.to_k = nn.Linear(320, 320, bias=False)

    torch.manual_seed(0)
    att = Attention().to(DEV).to(torch.bfloat16)
    L.inject_trainable_lora(att, r=4)
    for m in (att.to_q, att.to_k):
        m.lora_up.weight.data.normal_(0, 0.05)
    L.set_grouping(True)
    x = torch.randn(256, 320, device=DEV, dtype=torch.bfloat16)
    for _ in range(2):                                   # learn: q and k share x
        att.to_q(x); att.to_k(x)
    other = torch.randn(256, 320, device=DEV, dtype=torch.bfloat16)
    q = att.to_q(x)
    k = att.to_k(other)                                  # speculation for k is discarded
    wantq = O.lora_linear_forward(x, att.to_q.linear.weight, None, att.to_q.lora_down.weight, att.to_q.lora_up.weight, 1.0)
    wantk = O.lora_linear_forward(other, att.to_k.linear.weight, None, att.to_k.lora_down.weight, att.to_k.lora_up.weight, 1.0)
    assert rel(q, wantq) < 2 ** -7 and rel(k, wantk) < 2 ** -7
    x.add_(1.0)                                          # in-place edit bumps the version: no stale reuse
    q2 = att.to_q(x.clone()); x.sub_(1.0)
    k2 = att.to_k(x)
    assert rel(k2, O.lora_linear_forward(x, att.to_k.linear.weight, None, att.to_k.lora_down.weight, att.to_k.lora_up.weight, 1.0)) < 2 ** -7


def test_sibling_edited_between_the_two_calls_is_not_served_stale():
    """The family is launched at the FIRST sibling's call; a sibling whose scale / factors / frozen
    weight are edited before its own call must not receive the parked (now stale) output."""
    import lora_b200 as L
    from oracle import lora_ops as O

    class Attention(nn.Module):
        def __init__(self):
            super().__init__()
            self.to_q = nn.Linear(320, 320, bias=False)
            self.to_k = nn.Linear(320, 320, bias=False)

    torch.manual_seed(0)
    att = Attention().to(DEV).to(torch.bfloat16)
    L.inject_trainable_lora(att, r=4)
    for m in (att.to_q, att.to_k):
        m.lora_up.weight.data.normal_(0, 0.05)
    L.set_grouping(True)
    x = torch.randn(256, 320, device=DEV, dtype=torch.bfloat16)
    for _ in range(2):
        att.to_q(x); att.to_k(x)

    def want(m, scale):
        return O.lora_linear_forward(x, m.linear.weight, None, m.lora_down.weight, m.lora_up.weight, scale)

    att.to_q(x)                                   # launches the family, parks k's output
    att.to_k.scale = 0.25                         # tune_lora_scale on one sibling
    assert rel(att.to_k(x), want(att.to_k, 0.25)) < 2 ** -7
    att.to_q(x)
    with torch.no_grad():
        att.to_k.lora_up.weight.mul_(3.0)         # in-place factor edit (version bump)
    assert rel(att.to_k(x), want(att.to_k, 0.25)) < 2 ** -7
    att.to_q(x)
    att.to_k.linear.weight = nn.Parameter(torch.randn(320, 320, device=DEV, dtype=torch.bfloat16) * 0.05,
                                          requires_grad=False)     # external re-assignment (lora.py:290)
    assert rel(att.to_k(x), want(att.to_k, 0.25)) < 2 ** -7


def test_training_step_with_grouping_matches_without():
    import lora_b200 as L
    from lora_b200.host.clip import build_text_encoder
    from lora_b200.host.unet_sd15 import UNet2DConditionModel, UNetConfig
    from lora_b200.train import LoraTrainStep, StepConfig
    losses = []
    for grouped in (False, True):
        L.set_grouping(grouped)
        torch.manual_seed(5)
        unet = UNet2DConditionModel(UNetConfig.tiny()).to(DEV).to(torch.bfloat16)
        text = build_text_encoder(tiny=True).to(DEV).to(torch.bfloat16)
        unet.requires_grad_(False); text.requires_grad_(False)
        L.inject_trainable_lora(unet, r=4)
        L.inject_trainable_lora(text, target_replace_module={"CLIPAttention"}, r=4)
        g = torch.Generator(device=DEV).manual_seed(2)
        for m in list(unet.modules()) + list(text.modules()):
            if type(m).__name__ == "LoraInjectedLinear":
                m.lora_up.weight.data.normal_(0, 0.05, generator=g)
        tr = LoraTrainStep(unet, text, StepConfig(use_cuda_graph=grouped, graph_warmup=2),
                           latent_shape=(1, 4, 16, 16), device=DEV)
        torch.manual_seed(9)
        tr.latents.copy_(torch.randn(1, 4, 16, 16, device=DEV) * 0.18215)
        tr.input_ids.copy_(torch.randint(0, 1000, (1, 77), device=DEV))
        out = []
        torch.manual_seed(77)
        if grouped:
            tr.prepare()
            assert tr.graph is not None, tr.graph_error
        for _ in range(4 if not grouped else 2):
            out.append(float(tr.step_device()))
        losses.append(out)
        params = tr.arena.p.clone()
    assert all(l == l and 0 < l < 10 for l in losses[0] + losses[1])
