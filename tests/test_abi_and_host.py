"""CPU-side checks: the C-ABI library loads and exports every symbol include/dmcmc.h declares
(no compute without a GPU), it refuses to run without CUDA (no CPU fallback), and the host
mirror's bookkeeping (rho handling, adaptation, path ring) follows the reference."""
import ctypes
import math
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib(pkg):
    from diffusionmcmc_jl_b200 import _lib, build
    build.build()
    return _lib


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "dmcmc.h")).read()
    declared = set(re.findall(r"\b(dmcmc_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    l = lib.load()
    for name in declared:
        assert hasattr(l, name), f"{name} declared in include/dmcmc.h but not exported"
    assert declared == set(lib.SYMBOLS), declared ^ set(lib.SYMBOLS)


def test_ctypes_mirror_matches_the_header_prototypes(lib):
    """The hand-written ctypes table (_lib.SYMBOLS) against the prototypes of include/dmcmc.h: same number of parameters,
    and pointer / 64-bit / 32-bit / double kinds in the same positions (a mismatch would pass the symbol-name test and
    corrupt arguments at run time)."""
    hdr = open(os.path.join(ROOT, "include", "dmcmc.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    protos = re.findall(r"\b([A-Za-z_][A-Za-z0-9_ \*]*?)\b(dmcmc_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr)
    assert len(protos) >= 40

    def kind(decl):
        d = decl.strip()
        if "*" in d or "[" in d:   # array parameters decay to pointers
            return "ptr"
        if re.search(r"\b(int64_t|uint64_t|size_t|long long)\b", d):
            return "i64"
        if re.search(r"\bdouble\b", d):
            return "f64"
        return "i32"

    C = C_ = ctypes
    ck = {C.c_void_p: "ptr", C.c_char_p: "ptr", C.c_int64: "i64", C.c_uint64: "i64", C.c_size_t: "i64", C.c_double: "f64",
          C.c_int: "i32", C.c_int32: "i32", C.c_uint32: "i32", C.c_uint8: "i32"}
    seen = 0
    for ret, name, args in protos:
        if name not in lib.SYMBOLS:
            continue
        params = [a for a in (x.strip() for x in args.split(",")) if a and a != "void"]
        restype, argtypes = lib.SYMBOLS[name]
        assert len(params) == len(argtypes), f"{name}: header has {len(params)} parameters, ctypes mirror {len(argtypes)}"
        for i, (a, t) in enumerate(zip(params, argtypes)):
            tk = "ptr" if (isinstance(t, type) and issubclass(t, C_._Pointer)) or t in (C.c_void_p, C.c_char_p) else ck.get(t)
            assert tk == kind(a), f"{name}: parameter {i} is `{a}` in the header, {t} in the ctypes mirror"
        seen += 1
    assert seen == len(lib.SYMBOLS)


def test_layout_helper_matches_host_mirror(lib, pkg):
    l = lib.load()
    for nobs in ([100], [5, 8, 3], [1]):
        for h in (0, 1, 2, 7):
            for pat in (0, 1):
                r = np.array(nobs, np.int32)
                J = int(r.sum())
                i1, i2, pin = (np.zeros(J, np.int32) for _ in range(3))
                nb = l.dmcmc_chequerboard_layout(len(r), r.ctypes.data_as(lib._ip), h, pat,
                                                 i1.ctypes.data_as(lib._ip), i2.ctypes.data_as(lib._ip),
                                                 pin.ctypes.data_as(lib._ip))
                a = pkg.problems.chequerboard_layout(nobs, h, pat)
                assert nb == len(a[0])
                assert np.array_equal(i1[:nb], a[0]) and np.array_equal(i2[:nb], a[1]) and np.array_equal(pin[:nb], a[2])


def test_no_cpu_fallback(lib):
    """Without a CUDA device the product refuses to create a context."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    l = lib.load()
    cfg = lib.Config(lib.ABI_VERSION, 0, 2, 4, 0, 0, 0, 1, 1e-4, 1)
    h = ctypes.c_void_p()
    assert l.dmcmc_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    assert b"no CPU fallback" in l.dmcmc_last_error(None)


def test_product_does_not_touch_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may use oracle/."""
    pk = os.path.join(ROOT, "diffusionmcmc.jl_b200")
    for dp, _, fs in os.walk(pk):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                for pat in (r"^\s*(from|import)\s+oracle", r"#include\s+[<\"][^>\"]*oracle",
                            r"liboracle", r"orc_[a-z_]+\s*\("):
                    assert not re.search(pat, src, re.M), (f, pat)


def test_path_imputation_rho_expansion(pkg):
    # src/updates.jl:33-74
    u = pkg.PathImputation(0.96)
    assert u.rho_s == [[0.96]]
    u.init_update([2, 2, 1])
    assert u.rho_s == [[0.96, 0.96], [0.96, 0.96], [0.96]]
    rho = u.rho_per_interval([0, 2, 4], [1, 3, 4], 5)
    assert np.array_equal(rho, np.full(5, 0.96))
    u2 = pkg.PathImputation([[0.5, 0.6], [0.7]])
    u2.init_update([2, 3])
    assert u2.rho_s == [[0.5, 0.6], [0.7, 0.7, 0.7]]


def test_adaptation_follows_the_reference(pkg):
    # src/adaptation.jl:32-40 defaults
    a = pkg.AdaptationPathImputation()
    assert (a.target_accpt_rate, a.adapt_every_k_steps, a.scale, a.min, a.max, a.offset) == \
        (0.234, 100, 1.0, 1e-12, 1.0 - 1e-12, 1e2)
    assert a.acceptance_rate() == 0.0
    for i in range(100):
        a.register(i % 2 == 0)
    assert a.time_to_update() and a.acceptance_rate() == 0.5
    rho = a.readjust(0.9, 100)   # acceptance above target => rho decreases by delta in logit space
    d = 1.0 / math.sqrt(max(1.0, 100 / 100 - 100.0))
    expect = 1 / (1 + math.exp(-(math.log(0.9) - math.log(0.1) - d)))
    assert abs(rho - expect) < 1e-15 and rho < 0.9
    assert a.proposed == 0 and a.accepted == 0
    for i in range(100):
        a.register(False)
    assert a.readjust(0.5, 200) > 0.5
    m = pkg.AdaptationMultiPathImputation(pkg.AdaptationPathImputation(adapt_every_k_steps=10))
    m.resize(3)
    assert len(m.v) == 3 and m.v[2].adapt_every_k_steps == 10 and m.v[2] is not m.v[0]


def test_path_saving_buffer_ring(pkg):
    # src/path_saving_buffer.jl:7-64
    cc = pkg.CyclicCounter(3)
    assert [cc(nxt=True) for _ in range(7)] == [1, 2, 3, 1, 2, 3, 1]
    assert np.array_equal(pkg.glue_containers([[1, 2, 3], [3, 4, 5], [5, 6]]), [1, 2, 4 - 1, 4, 5, 6][:0] + [1, 2, 3, 4, 5, 6])
    buf = pkg.PathSavingBuffer([np.arange(5.0)], 2, 2)
    buf.save([np.ones((5, 2))])
    buf.save([2 * np.ones((5, 2))])
    buf.save([3 * np.ones((5, 2))])
    assert buf.XX[0][0]["x"][0, 0] == 3 and buf.XX[1][0]["x"][0, 0] == 2


def test_blocking_plan_carries_the_pattern_across_iterations(pkg):
    """extra_schedule_params with update_layout defined: each BlockingChequerboardSwitch moves to the other pattern.
    Even number of switches per pass (the reference's assertion, src/schedule.jl:89): an update keeps its pattern;
    odd number: the switch count carries over, so the pattern alternates from one MCMC iteration to the next --
    in the general path exactly as in the device loop (otherwise [Switch, Imputation, ParamUpdate] would never
    re-impute the block end points)."""
    from diffusionmcmc_jl_b200.backend import _BlockingPlan
    S, I, R = pkg.BlockingChequerboardSwitch, pkg.PathImputation, pkg.RandomWalkUpdate
    plan = _BlockingPlan([10], [S(2), I(0.9), S(2), I(0.9)])
    assert [plan.pattern(0, it) for it in (1, 2, 3)] == [0, 0, 0] and [plan.pattern(1, it) for it in (1, 2, 3)] == [1, 1, 1]
    assert plan.patterns_seen(0) == [0] and plan.patterns_seen(1) == [1]
    plan = _BlockingPlan([10], [S(1), I(0.9), R(0.1, [2])])
    assert [plan.pattern(0, it) for it in (1, 2, 3, 4)] == [0, 1, 0, 1]
    assert [plan.pattern(1, it) for it in (1, 2, 3, 4)] == [0, 1, 0, 1]    # the parameter step runs under the same layout
    assert plan.patterns_seen(0) == [0, 1]
    lay0, lay1 = plan.layout(0, 0), plan.layout(0, 1)
    assert list(lay0[0]) == [0, 2, 4, 6, 8] and list(lay1[0]) == [0, 1, 3, 5, 7, 9]
    plan = _BlockingPlan([10], [I(0.9), R(0.1, [2])])
    assert plan.pattern(0, 5) is None and plan.patterns_seen(0) == [None]
    assert list(plan.layout(0, None)[1]) == [9]
    plan = _BlockingPlan([10], [S(1), I(0.9), S(1), I(0.9), S(1), I(0.9)])   # three switches per pass
    assert [[plan.pattern(u, it) for u in range(3)] for it in (1, 2)] == [[0, 1, 0], [1, 0, 1]]


def test_imputation_state_keeps_rho_per_block_and_adapts(pkg):
    from diffusionmcmc_jl_b200.backend import _ImputationState
    lay = pkg.problems.chequerboard_layout([10], 2, 1)       # blocks [0,1] [2..5] [6..9]
    u = pkg.PathImputation(0.5, adpt=pkg.AdaptationPathImputation(adapt_every_k_steps=4, scale=1.0, offset=0.0))
    st = _ImputationState(u, lay)
    assert [len(r) for r in st.rho_s] == [2, 4, 4] and st.rho_per_interval(10).tolist() == [0.5] * 10
    assert st.adpt is not u.adpt and len(st.adpt.v) == 3
    assert not st.register_and_adapt(np.array([2, 0, 2]), np.full(3, 2), 1)
    assert st.register_and_adapt(np.array([2, 0, 0]), np.full(3, 2), 2)        # 4 proposals per block registered
    r = st.rho_per_interval(10)
    # acceptance 100 % and 50 % are above the 0.234 target: rho goes down (src/adaptation.jl:82-90); 0 %: up
    assert r[0] < 0.5 and r[2] > 0.5 and r[6] < 0.5 and r[1] == r[0] and r[9] == r[6]


def test_parameter_updates_refuse_a_chain_batch(pkg):
    """The chains of one context share theta: a joint accept/reject over C chains would target likelihood^C."""
    data = pkg.problems.fhn_spec(n_obs=5, C=4, dt=0.01)
    mcmc = pkg.MCMC([pkg.PathImputation(0.9), pkg.RandomWalkUpdate(0.1, [2])])
    with pytest.raises(ValueError, match="n_chains == 1"):
        pkg.run_(mcmc, 3, data, data["theta"])


def test_table_fragment_order_of_the_tensor_core_kernel(tmp_path):
    """The d = 40 table record stores H and Psi' in mma.m8n8k4 fragment order (big_frag): a bijection, and exactly the
    element a lane of the Lorenz-96 kernel expects at the address it loads (host-only program, compiled with nvcc)."""
    import shutil
    import subprocess
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not on PATH")
    exe = tmp_path / "frag_order_check"
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "frag_order_check.cu")
    subprocess.check_call(["nvcc", "-std=c++17", "-o", str(exe), src])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "fragment order ok" in out.stdout, out.stdout + out.stderr
