"""Shared comparison helpers for the parity tests (oracle vs CUDA path)."""
import numpy as np

RTOL = 1e-10  # north_star: imputed paths and per-block ll within 1e-10 relative (FP64)


def close(a, b, rtol=RTOL):
    """|a-b| <= rtol * max(1, |b|) element-wise."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    with np.errstate(invalid="ignore"):
        ok = np.abs(a - b) <= rtol * np.maximum(1.0, np.abs(b))
    return bool(np.all(ok | both_inf))


def maxrel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    fin = np.isfinite(a) & np.isfinite(b)
    if not fin.any():
        return 0.0
    return float(np.max(np.abs(a - b)[fin] / np.maximum(1.0, np.abs(b)[fin])))


def interior_mask(N):
    """points 1..N_j of every interval container (point 0 is the carried start point)."""
    m = []
    for n in N:
        m += [False] + [True] * int(n)
    return np.array(m)


def accept_mismatch_is_tie(acc_g, acc_o, E, llp, ll_before, rtol=RTOL):
    """Accept sequences must agree except where |E + (ll deg - ll)| is within tolerance
    (SURVEY Appendix D item 3)."""
    bad = acc_g != acc_o
    if not bad.any():
        return True
    gap = np.abs(E + (llp - ll_before))
    return bool(np.all(gap[bad] <= rtol * np.maximum(1.0, np.abs(ll_before[bad]))))
