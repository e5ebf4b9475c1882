"""B200-native path-imputation backend with the DiffusionMCMC.jl plugin surface.

Only the hot path: pCN Wiener refresh, guided Euler-Maruyama solve,
Girsanov log-likelihood, Metropolis accept/swap, for every (chain, block).  The arithmetic runs
in hand-written CUDA for sm_100a behind the C ABI of include/dmcmc.h; this package is the
host-side mirror of the reference's operator interface (names as in
/root/reference src/DiffusionMCMC.jl:31-33).  There is no CPU fallback.
"""
from . import problems  # noqa: F401
from .adaptation import (AdaptationMultiPathImputation, AdaptationPathImputation,  # noqa: F401
                         NoAdaptation)
from .blocking import BlockingChequerboardSwitch  # noqa: F401
from .path_saving_buffer import CyclicCounter, PathSavingBuffer, glue_containers  # noqa: F401
from .schedule import MCMCSchedule, Step, extra_transitions, init_block_layout  # noqa: F401


def __getattr__(name):
    # backend/engine bind the CUDA library lazily so that host-only logic imports without it
    if name in ("DiffusionMCMCBackend", "PathImputation", "RandomWalkUpdate", "DiffusionConjugGsnUpdate", "MCMC", "run_",
                "ImputationRunner", "DiffusionGlobalWorkspace", "DiffusionLocalWorkspace"):
        from . import backend
        return getattr(backend, name)
    if name == "Engine":
        from .engine import Engine
        return Engine
    raise AttributeError(name)
