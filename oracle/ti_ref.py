"""ORACLE (test infrastructure only): the textual-inversion update of the reference,
lora_diffusion/cli_lora_pti.py:446-479, restated on the WHOLE embedding table exactly as the
reference performs it (AdamW over every row, norm decay of the trained rows, restore of all other
rows from the original table), in float64."""
import torch


def ti_table_step(table, grad, m, v, step, lr, index_updates, orig_table, betas=(0.9, 0.999), eps=1e-8,
                  weight_decay=0.0, clip_ti_decay=True):
    """table, grad, m, v, orig_table: [V, D]; index_updates: bool [V]. Returns (table, m, v)."""
    t64 = lambda x: x.detach().to("cpu", torch.float64).clone()
    table, grad, m, v, orig = t64(table), t64(grad), t64(m), t64(v), t64(orig_table)
    b1, b2 = betas
    # torch.optim.AdamW over the full table (cli_lora_pti.py:448, optimizer built at :905-909)
    table = table * (1.0 - lr * weight_decay)
    m = b1 * m + (1 - b1) * grad
    v = b2 * v + (1 - b2) * grad * grad
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    table = table - (lr / bc1) * m / (v.sqrt() / bc2 ** 0.5 + eps)
    if clip_ti_decay:                                           # :451-468
        rows = table[index_updates]
        pre = rows.norm(dim=-1, keepdim=True)
        lam = min(1.0, 100 * lr)
        table[index_updates] = torch.nn.functional.normalize(rows, dim=-1) * (pre + lam * (0.4 - pre))
    table[~index_updates] = orig[~index_updates]                # :477-479
    return table, m, v
