"""Golden vectors (oracle-generated, see tests/golden/make_golden.py): the oracle must keep
reproducing them bit for bit (CPU), and the CUDA path must match them to 1e-10 (GPU)."""
import glob
import os

import numpy as np
import pytest

from helpers import check, interior_mask

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))


def _spec(g):
    s = {k[5:]: g[k] for k in g.files if k.startswith("spec_")}
    for k in ("model", "d", "m"):
        s[k] = int(s[k])
    for k in ("pin_eps", "t0"):
        s[k] = float(s[k])
    return s


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_oracle_reproduces_golden(pkg, orc, path):
    g = np.load(path)
    o = orc.Oracle(_spec(g))
    o.init_paths(Z=g["Z0"])
    assert np.array_equal(o.ll, g["ll_init"])
    h = int(g["half_len"])
    for it in range(3):
        if h:
            o.relayout(*pkg.problems.chequerboard_layout(o.rec_nobs, h, it % 2))
            assert np.array_equal(o.ll, g[f"ll_relayout_{it}"])
        llp, ll, acc = o.impute(g[f"Z_{it}"], g[f"E_{it}"])
        assert np.array_equal(llp, g[f"llp_{it}"]) and np.array_equal(acc, g[f"acc_{it}"])
        assert np.array_equal(o.accepted_X(), g[f"Xacc_{it}"])
    assert np.array_equal(o.glue_path(0), g["glued_chain0"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_cuda_matches_golden(pkg, path):
    """No oracle in the loop: committed vectors only, device backward filter included."""
    from diffusionmcmc_jl_b200.engine import Engine
    g = np.load(path)
    e = Engine(_spec(g))
    e.init_paths(Z=g["Z0"])
    name = os.path.basename(path)[:-4]
    check(f"golden {name} init ll", e.get_ll(), g["ll_init"])
    h = int(g["half_len"])
    im = interior_mask(e.N)
    for it in range(3):
        if h:
            e.relayout(*pkg.problems.chequerboard_layout(e.rec_nobs, h, it % 2))
            check(f"golden {name} it {it} relayout ll", e.get_ll(), g[f"ll_relayout_{it}"])
        llp, ll, acc = e.impute(it, g[f"Z_{it}"], g[f"E_{it}"])
        check(f"golden {name} it {it} ll deg", llp, g[f"llp_{it}"])
        assert np.array_equal(acc, g[f"acc_{it}"])
        X, W = e.download_state(0)
        check(f"golden {name} it {it} X", X[:, im], g[f"Xacc_{it}"][:, im])
        check(f"golden {name} it {it} W", W[:, im], g[f"Wacc_{it}"][:, im])
    check(f"golden {name} glued path", e.download_path(0), g["glued_chain0"])
