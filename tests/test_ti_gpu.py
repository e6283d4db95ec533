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
