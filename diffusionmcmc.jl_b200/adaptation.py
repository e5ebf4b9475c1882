"""Adaptation of the pCN memory parameter rho (/root/reference src/adaptation.jl), host side.
Consumes accept flags, produces rho per block; cheap, so it stays on the host."""
import copy
import math


def sigmoid(x, a=1.0):   # src/adaptation.jl:92
    return 1.0 / (1.0 + math.exp(-a * x))


def logit(x, a=1.0):     # src/adaptation.jl:93
    return (math.log(x) - math.log(1 - x)) / a


class NoAdaptation:
    pass


class AdaptationPathImputation:
    """src/adaptation.jl:22-41 (same keyword names and defaults)."""

    def __init__(self, target_accpt_rate=0.234, adapt_every_k_steps=100, scale=1.0, min=1e-12,
                 max=1.0 - 1e-12, offset=1e2):
        self.proposed, self.accepted = 0, 0
        self.target_accpt_rate, self.adapt_every_k_steps = target_accpt_rate, adapt_every_k_steps
        self.scale, self.min, self.max, self.offset = scale, min, max, offset

    def acceptance_rate(self):            # :48-50
        return 0.0 if self.proposed == 0 else self.accepted / self.proposed

    def reset(self):                      # :69-72
        self.proposed = self.accepted = 0

    def acceptance_rate_and_reset(self):  # acceptance_rate!, :58-62
        a = self.acceptance_rate()
        self.reset()
        return a

    def register(self, accepted, count=1):  # :100-103 (count > 1: many chains share the block)
        self.accepted += int(accepted)
        self.proposed += int(count)

    def time_to_update(self):             # :111-113
        return self.proposed >= self.adapt_every_k_steps

    def compute_delta(self, mcmc_iter):
        """eMCMC.compute_delta [UPSTREAM-RECALL]: scale / sqrt(max(1, iter/k - offset))."""
        return self.scale / math.sqrt(max(1.0, mcmc_iter / self.adapt_every_k_steps - self.offset))

    def readjust(self, rho, mcmc_iter):
        """src/adaptation.jl:82-90 with eMCMC.compute_eps(rho, adpt, a_r, delta, -1.0, logit, sigmoid)
        [UPSTREAM-RECALL]: step in logit space, DOWN when the acceptance rate is above target."""
        delta = self.compute_delta(mcmc_iter)
        a_r = self.acceptance_rate_and_reset()
        step = -1.0 * (1.0 if a_r > self.target_accpt_rate else -1.0) * delta
        r = min(max(rho, self.min), self.max)
        new = sigmoid(logit(r) + step)
        return min(max(new, self.min), self.max)


class AdaptationMultiPathImputation:
    """One AdaptationPathImputation per block/recording (src/adaptation.jl:122-167)."""

    def __init__(self, adpt):
        self.v = [adpt]

    def resize(self, n):                  # resize_adpt!, :162-167
        while len(self.v) < n:
            self.v.append(copy.deepcopy(self.v[0]))

    def register(self, accepted_counts, proposed_counts):
        for a, acc, prop in zip(self.v, accepted_counts, proposed_counts):
            a.register(acc, prop)

    def time_to_update(self):             # :139-141
        return self.v[0].time_to_update()
