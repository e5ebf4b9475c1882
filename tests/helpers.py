"""Shared comparison helpers for the parity tests (oracle vs CUDA path)."""
import os

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


_REPORT = os.environ.get("DMCMC_PARITY_REPORT", "")


def check(what, got, ref, rtol=RTOL):
    """Parity assertion that also RECORDS the achieved error: |got - ref| <= rtol * max(1, |ref|)
    element-wise; one line `what, max-rel, tolerance` goes to stdout (pytest -s / -rA) and, when
    DMCMC_PARITY_REPORT names a file, is appended there (profiles/r2_parity_report.txt is such a run)."""
    err = maxrel(got, ref)
    line = f"parity {what}: max-rel {err:.3e} (tol {rtol:.0e})"
    print(line)
    if _REPORT:
        with open(_REPORT, "a") as f:
            f.write(line + "\n")
    assert close(got, ref, rtol), line
    return err


def check_scaled(what, got, ref, rtol=RTOL):
    """Norm-wise variant for tables whose entries span orders of magnitude (H, Psi ~ 1/pin_eps next to O(1) entries):
    max |got - ref| <= rtol * max(1, max |ref|)."""
    got, ref = np.asarray(got, float), np.asarray(ref, float)
    sc = max(1.0, float(np.abs(ref).max())) if ref.size else 1.0
    err = float(np.abs(got - ref).max()) / sc if ref.size else 0.0
    line = f"parity {what}: max-abs/scale {err:.3e} (scale {sc:.2e}, tol {rtol:.0e})"
    print(line)
    if _REPORT:
        with open(_REPORT, "a") as f:
            f.write(line + "\n")
    assert err <= rtol, line
    return err


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


class OracleEngine:
    """Adapter giving the CPU oracle the Engine interface that parallel.BlockShardedSampler
    drives (tests only): lets the multi-process host logic run under gloo without a GPU."""

    def __init__(self, orc, spec, interval_offset=0, seed=0, chain_offset=0):
        self.o = orc.Oracle(spec, interval_offset=interval_offset, chain_offset=chain_offset)
        self.seed = seed
        self.C, self.d, self.N = self.o.C, self.o.d, self.o.N
        self.layout, self.mask = None, None

    def close(self):
        pass

    def init_paths(self):
        self.o.init_paths(seed=self.seed)

    def set_layout(self, i1, i2, pin):
        self.layout = (np.asarray(i1, np.int32), np.asarray(i2, np.int32), np.asarray(pin, np.int32))

    def set_block_mask(self, act):
        self.mask = None if act is None else np.asarray(act, np.uint8)

    def impute(self, it):
        o = self.o
        o.relayout(*self.layout)   # the CUDA path fuses this into the imputation kernel
        mask = None if self.mask is None else np.broadcast_to(self.mask, (o.C, o.nblk)).copy()
        return o.impute(seed=self.seed, it=it, unit_mask=mask)

    def _acc(self, c, j):
        o = self.o
        return o.X[o.selX[c, j]][c, o.pt_off[j]:o.pt_off[j + 1]]

    def download_segment(self, j0, j1):
        o = self.o
        return np.stack([np.concatenate([self._acc(c, j)[1:] for j in range(j0, j1)]) for c in range(o.C)])

    def upload_segment(self, j0, j1, X):
        o = self.o
        for c in range(o.C):
            off = 0
            for j in range(j0, j1):
                cont = self._acc(c, j)
                if j > 0:
                    cont[0] = self._acc(c, j - 1)[-1]
                cont[1:] = X[c, off:off + o.N[j]]
                off += o.N[j]
            if j1 < o.J:
                self._acc(c, j1)[0] = X[c, -1]

    def set_start(self, x0):
        o = self.o
        for c in range(o.C):
            o.X[0][c, 0] = x0[c, 0]
            o.X[1][c, 0] = x0[c, 0]
