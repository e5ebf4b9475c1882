"""The drop-in boundary without Python in between: tests/abi_smoke.c is compiled with gcc against include/dmcmc.h
ONLY and linked to libdmcmc_b200.so.  CPU: it compiles warning-free as C99, every symbol it uses resolves, and the
library refuses to create a context without a GPU.  GPU: it drives create -> set_* -> backward_filter -> init_paths
-> impute -> relayout -> impute -> param_loglik / commit -> download_path / save_path_async, and every number it
prints equals, bit for bit, what the same sequence gives through the ctypes mirror (_lib.py) -- so a struct-layout or
argument-order mismatch between the header and the mirror cannot go unnoticed."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NOBS, NSTEP, NCH, SEED = 6, 10, 4, 20261020


@pytest.fixture(scope="module")
def binary(pkg, tmp_path_factory):
    from diffusionmcmc_jl_b200 import build
    lib = build.build()
    libdir = os.path.dirname(lib)
    exe = str(tmp_path_factory.mktemp("abi") / "abi_smoke")
    cmd = ["gcc", "-std=c99", "-O1", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "abi_smoke.c"), "-o", exe, "-L", libdir, "-l:libdmcmc_b200.so",
           f"-Wl,-rpath,{libdir}", "-lm"]
    subprocess.check_call(cmd)
    return exe


def test_header_is_plain_c_and_links(binary):
    import torch
    out = subprocess.run([binary, "--symbols"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "sizeof(dmcmc_config) 56" in out.stdout      # the ctypes mirror's layout (below) must say the same
    from diffusionmcmc_jl_b200 import _lib
    import ctypes
    assert ctypes.sizeof(_lib.Config) == 56
    if not torch.cuda.is_available():
        assert "no CPU fallback" in out.stdout


def _spec(pkg):
    P = pkg.problems
    obs = np.array([-0.92, -0.71, -0.35, 0.48, 1.12, 1.31])
    t = np.array([0.1 * j + 0.1 * k / NSTEP for j in range(NOBS) for k in range(NSTEP + 1)])
    spec = P.make_spec(P.MODEL_FHN, [0.1, -0.8, 1.5, 0.0, 0.3], [-1.0, -1.0], 0.1 * np.arange(1, NOBS + 1), obs[:, None],
                       [[1.0, 0.0]], [[0.01]], 0.01, C=NCH, rho=0.9)
    assert list(spec["N"]) == [NSTEP] * NOBS
    spec["t"] = t
    return spec


@pytest.mark.gpu
def test_c_consumer_equals_python_mirror(pkg, binary):
    from diffusionmcmc_jl_b200.engine import Engine
    out = subprocess.run([binary], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "abi_smoke ok" in out.stdout
    got = {}
    for line in out.stdout.splitlines():
        k, *vals = line.split()
        got[k] = vals
    f = lambda k: np.array([float(x) for x in got[k]])
    P = pkg.problems
    e = Engine(_spec(pkg), seed=SEED)
    e.init_paths()
    assert np.array_equal(e.get_ll().ravel(), f("init_ll"))
    llp, ll, acc = e.impute(0)
    assert np.array_equal(llp.ravel(), f("impute0_llp")) and np.array_equal(ll.ravel(), f("impute0_ll"))
    assert [int(x) for x in got["impute0_acc"]] == list(acc.ravel().astype(int))
    e.relayout(*P.chequerboard_layout(e.rec_nobs, 1, 1))
    llp, ll, acc = e.impute(1)
    assert np.array_equal(llp.ravel(), f("impute1_llp")) and np.array_equal(ll.ravel(), f("impute1_ll"))
    assert [int(x) for x in got["impute1_acc"]] == list(acc.ravel().astype(int))
    llp, ok = e.param_loglik([0.1, -0.8, 1.55, 0.0, 0.3], critical=True)
    assert np.array_equal(llp.ravel(), f("param_llp"))
    assert [int(x) for x in got["param_ok"]] == list(ok.astype(int))
    e.param_commit(True)
    assert np.array_equal(e.get_ll().ravel(), f("param_ll_after_commit"))
    assert np.array_equal(e.download_path(2).ravel(), f("path_chain2"))
    saved = e.save_path_wait(e.save_path_async([2, 0]))
    assert np.array_equal(saved[0].ravel(), f("path_chain2")) and np.array_equal(saved[1].ravel(), f("path_chain0"))
    a, p = e.accept_counts()
    assert got["accept_counts"] == [f"{x}/{y}" for x, y in zip(a, p)]
