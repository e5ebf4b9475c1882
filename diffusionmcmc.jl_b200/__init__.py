"""B200-native path-imputation backend with the DiffusionMCMC.jl plugin surface.

Only the hot path: pCN Wiener refresh, guided Euler-Maruyama solve,
Girsanov log-likelihood, Metropolis accept/swap, for every (chain, block).  The arithmetic runs
in hand-written CUDA for sm_100a behind the C ABI of include/dmcmc.h; this package is the
host-side mirror of the reference's operator interface (names as in
/root/reference src/DiffusionMCMC.jl:31-33).  There is no CPU fallback.
"""
from . import problems  # noqa: F401
