"""ORACLE (test infrastructure only): restatement of the reference's SVD distillation arithmetic,
lora_diffusion/cli_svd.py:24-92 (`overwrite_base`), per weight pair.

    residual = (W_tuned - W_base).float()                 (conv: flattened to [Cout, Cin*kh*kw])
    U, S, Vh = svd(residual) ; U_r = U[:, :r] diag(S[:r]) ; Vh_r = Vh[:r]
    hi = quantile(cat(U_r.flatten(), Vh_r.flatten()), clamp_quantile) ; clamp both to [-hi, hi]
    up = U_r (conv: [Cout, r, 1, 1]) ; down = Vh_r (conv: [r, Cin, kh, kw])

Sign convention: singular vectors are defined up to a joint sign per component; the symmetric
clamp commutes with that sign, but `hi` itself depends on it (the quantile is taken over signed
values). Parity against a different SVD algorithm is therefore stated on sign-invariant
quantities: the singular values, the unclamped rank-r product, and the clamp rule applied to
whichever factors an implementation produced (tests/test_svd_*.py).
"""
import torch


def svd_distill_pair(W_base: torch.Tensor, W_tuned: torch.Tensor, rank: int,
                     clamp_quantile: float = 0.99, dtype=torch.float32):
    residual = (W_tuned - W_base).to(dtype)
    shape = residual.shape
    mat = residual.flatten(start_dim=1)
    U, S, Vh = torch.linalg.svd(mat)           # full_matrices default, as in the reference
    U = U[:, :rank] @ torch.diag(S[:rank])
    Vh = Vh[:rank, :]
    hi = torch.quantile(torch.cat([U.flatten(), Vh.flatten()]), clamp_quantile)
    U = U.clamp(-hi, hi)
    Vh = Vh.clamp(-hi, hi)
    if len(shape) == 4:
        U = U.reshape(U.shape[0], U.shape[1], 1, 1)
        Vh = Vh.reshape(rank, shape[1], shape[2], shape[3])
    return U, Vh, S[:rank], float(hi)


def clamp_rule(up: torch.Tensor, down: torch.Tensor, clamp_quantile: float = 0.99):
    """The reference's clamp applied to arbitrary (up = U diag(S), down = Vh) factors."""
    hi = torch.quantile(torch.cat([up.flatten(), down.flatten()]).float(), clamp_quantile)
    return up.clamp(-hi, hi), down.clamp(-hi, hi), float(hi)
