"""Ring of the N most recently saved (glued) paths (/root/reference src/path_saving_buffer.jl)."""
import numpy as np


class CyclicCounter:
    """src/path_saving_buffer.jl:7-27 (1-based like the reference)."""

    def __init__(self, n, first=1, start_from=None):
        self.first, self.last, self.i = first, n, (first if start_from is None else start_from)

    def next(self):
        self.i = (self.i % self.last) + 1  # mod1(i+1, last)
        if self.i < self.first:
            self.i = self.first

    def __call__(self, nxt=False):
        i = self.i
        if nxt:
            self.next()
        return i


def glue_containers(xs):
    """src/path_saving_buffer.jl:46-50: all but the last point of every segment, then the end."""
    return np.concatenate([np.asarray(x)[:-1] for x in xs] + [np.asarray(xs[-1])[-1:]])


class PathSavingBuffer:
    """`XX[k][rec]` = dict(t=glued times, x=glued states) as in
    docs/src/tutorials/simple_inference.md:110-117."""

    def __init__(self, times_per_recording, d, N, n_chains=1):
        # one saved chain: x is [points][d] like the reference's glued trajectory; several: [chain][points][d]
        shape = (lambda t: (len(t), d)) if n_chains == 1 else (lambda t: (n_chains, len(t), d))
        self.XX = [[dict(t=np.array(t, float), x=np.zeros(shape(t))) for t in times_per_recording]
                   for _ in range(N)]
        self.iter = CyclicCounter(N)

    def save(self, glued_paths):
        """save!: copy glued accepted paths (one per recording) into the next ring slot."""
        slot = self.XX[self.iter(nxt=True) - 1]
        for tr, x in zip(slot, glued_paths):
            tr["x"][...] = x
