// Host-only check of the table record's mma.m8n8k4 fragment order (big_frag, csrc/dmcmc_device.cuh): the solve kernel's
// lane l reads, for k-pair kp and component tile nt, the 16 bytes at ((kp * 5 + nt) * 32 + l) * 2 and expects there
// T[8 kp + 2 (l % 4) + e][8 nt + l / 4], e = 0, 1 -- the B operands of its two DMMAs (dmcmc_solve_l96.cu).
#include <cstdio>
#include <vector>
#include "../diffusionmcmc.jl_b200/csrc/dmcmc_device.cuh"
int main() {
    std::vector<int> seen(1600, 0);
    for (int kk = 0; kk < 40; ++kk)
        for (int c = 0; c < 40; ++c) {
            const int f = dmcmc::big_frag(kk, c);
            if (f < 0 || f >= 1600 || seen[f]++) { printf("not a bijection at (%d, %d) -> %d\n", kk, c, f); return 1; }
        }
    for (int kp = 0; kp < 5; ++kp)
        for (int nt = 0; nt < 5; ++nt)
            for (int l = 0; l < 32; ++l)
                for (int e = 0; e < 2; ++e)
                    if (dmcmc::big_frag(8 * kp + 2 * (l % 4) + e, 8 * nt + l / 4) != ((kp * 5 + nt) * 32 + l) * 2 + e) {
                        printf("fragment order broken at kp %d nt %d lane %d e %d\n", kp, nt, l, e);
                        return 1;
                    }
    printf("fragment order ok\n");
    return 0;
}
