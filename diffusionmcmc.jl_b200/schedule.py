"""Schedule bookkeeping of the plugin (/root/reference src/schedule.jl) plus a minimal stand-in
for ExtensibleMCMC's `MCMCSchedule` iterator [UPSTREAM-RECALL: eMCMC is not on disk], pinned by
the reference's only golden vector (test/runtests.jl:8-54 -> tests/test_schedule.py).

A schedule state is the named tuple (prev_mcmciter, prev_pidx, mcmciter, pidx, same_layout)
built by `extra_transitions` (src/schedule.jl:1-13): `same_layout[r]` tells whether recording r
keeps its block layout between the previous update and this one
(`lt[prev_pidx] .== lt[pidx]`, src/schedule.jl:11).  Indices are 1-based like the reference.
"""
from collections import namedtuple

Step = namedtuple("Step", "prev_mcmciter prev_pidx mcmciter pidx same_layout")


def extra_transitions(extra_info, new_state):
    """src/schedule.jl:1-13."""
    lt = extra_info["layout_types"]
    a, b = lt[new_state["prev_pidx"] - 1], lt[new_state["pidx"] - 1]
    return Step(new_state["prev_mcmciter"], new_state["prev_pidx"], new_state["mcmciter"],
                new_state["pidx"], tuple(x == y for x, y in zip(a, b)))


class MCMCSchedule:
    """for step in MCMCSchedule(num_mcmc_steps, num_updates, exclude_updates; start, extra_info).

    `exclude_updates` is a list of (pidx, iterations) pairs: update `pidx` is skipped on those
    MCMC iterations (semantics pinned by test/runtests.jl:9-13, :39-50)."""

    def __init__(self, num_mcmc_steps, num_updates, exclude_updates=(), start=None, extra_info=None):
        self.num_mcmc_steps, self.num_updates = int(num_mcmc_steps), int(num_updates)
        self.excl = {int(p): set(int(i) for i in its) for p, its in exclude_updates}
        self.extra_info = extra_info or {"layout_types": tuple((1,) for _ in range(num_updates))}
        nrec = len(self.extra_info["layout_types"][0])
        self.start = start or Step(None, None, 1, 1, tuple(True for _ in range(nrec)))

    def _runs(self, it, p):
        return it not in self.excl.get(p, ())

    def __iter__(self):
        prev = None
        for it in range(1, self.num_mcmc_steps + 1):
            for p in range(1, self.num_updates + 1):
                if not self._runs(it, p):
                    continue
                if prev is None:
                    st = self.start
                    if (st.mcmciter, st.pidx) != (it, p):
                        st = Step(None, None, it, p, st.same_layout)
                else:
                    st = extra_transitions(self.extra_info, dict(
                        prev_mcmciter=prev.mcmciter, prev_pidx=prev.pidx, mcmciter=it, pidx=p))
                prev = st
                yield st


def init_block_layout(rec_nobs, blocking=None):
    """src/schedule.jl:97-100: one block (rec_idx, 1:n_obs) per recording when there is no blocking."""
    return tuple((r + 1, range(1, int(n) + 1)) for r, n in enumerate(rec_nobs))
