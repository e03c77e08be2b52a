// Device-side pieces of one differential-evolution generation shared by de.cu (stand-alone cl_de_trial) and
// icrf_energy.cu (cl_de_trial_curves: the trial step fused with the construction of the candidate curves).
#pragma once

#include "common.cuh"

namespace cl {
namespace de {

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// draw `slot` of candidate `i` in generation `gen`: uniform double in [0, 1) with 53 random bits
__host__ __device__ __forceinline__ double draw(uint64_t seed, uint64_t gen, uint32_t i, uint32_t slot) {
    const uint64_t key = splitmix64(seed ^ (gen * 0xD1342543DE82EF95ull));
    const uint64_t z = splitmix64(key + (((uint64_t)i << 8) | slot));
    return (double)(z >> 11) * 0x1.0p-53;
}

constexpr uint32_t kSlotR0 = 0, kSlotR1 = 1, kSlotFill = 2, kSlotCross = 3;     // then P crossover, P redraw slots
constexpr uint32_t kScaleCandidate = 0xFFFFFFFFu;                               // the per-generation dither draw

struct TrialConfig {
    double dither_lo, dither_hi, crossover;
    uint64_t seed;
};

// Component j of the trial vector of member i (unit cube) and its scaled parameter value.
__device__ __forceinline__ void trial_component(const double* __restrict__ pop, int S, int P, const TrialConfig& cfg,
                                                uint64_t gen, int i, int j, const double* __restrict__ lo,
                                                const double* __restrict__ hi, double& trial, double& param) {
    const uint64_t seed = cfg.seed;
    const double scale = cfg.dither_lo + (cfg.dither_hi - cfg.dither_lo) * draw(seed, gen, kScaleCandidate, 0);
    // two distinct members, both different from i
    int r0 = (int)(draw(seed, gen, i, kSlotR0) * (double)(S - 1));
    if (r0 >= i) ++r0;
    int r1 = (int)(draw(seed, gen, i, kSlotR1) * (double)(S - 2));
    const int a = i < r0 ? i : r0, b = i < r0 ? r0 : i;
    if (r1 >= a) ++r1;
    if (r1 >= b) ++r1;
    const int fill = (int)(draw(seed, gen, i, kSlotFill) * (double)P);
    const double xi = pop[i * P + j];
    double v = xi;
    if (j == fill || draw(seed, gen, i, kSlotCross + j) < cfg.crossover) {
        // same association as SciPy: x_i + scale * (((x_best - x_i) + x_r0) - x_r1), unfused
        const double d = __dsub_rn(__dadd_rn(__dsub_rn(pop[j], xi), pop[r0 * P + j]), pop[r1 * P + j]);
        v = __dadd_rn(xi, __dmul_rn(scale, d));
    }
    if (v > 1.0 || v < 0.0) v = draw(seed, gen, i, kSlotCross + P + j);          // _ensure_constraint
    trial = v;
    // _scale_parameters: 0.5 (lo + hi) + (x - 0.5) |hi - lo|
    param = __dadd_rn(__dmul_rn(0.5, __dadd_rn(lo[j], hi[j])), __dmul_rn(__dsub_rn(v, 0.5), fabs(__dsub_rn(hi[j], lo[j]))));
}

}  // namespace de
}  // namespace cl
