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

// One block (any size): selection, promotion of the best member to row 0, convergence test, generation counter.
// Shared by de_select_kernel (de.cu) and the last block of K4's tail kernel (icrf_energy.cu).
// status: [0] converged (0/1), [1] generations done, [2] members replaced this generation, [3] index the best
// member came from;  best[0] = lowest energy, best[1] = std(E), best[2] = mean(E)
__device__ __forceinline__ void select_block(double* __restrict__ pop, double* __restrict__ energies,
                                             const double* __restrict__ trial, const double* trial_energies, int S,
                                             int P, double tol, double atol, int64_t* __restrict__ generation,
                                             int32_t* __restrict__ status, double* __restrict__ best) {
    __shared__ int n_replaced;
    __shared__ int best_idx;
    if (threadIdx.x == 0) n_replaced = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
        const double te = trial_energies[i];
        if (te <= energies[i]) {                           // scipy _accept_trial: '<=' (a plateau keeps moving); NaN never replaces
            energies[i] = te;
            for (int j = 0; j < P; ++j) pop[i * P + j] = trial[i * P + j];
            atomicAdd(&n_replaced, 1);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // np.argmin: first index of the minimum (NaN energies cannot occur: K4 maps NaN to +inf)
        int l = 0;
        double m = energies[0];
        bool any_inf = false;
        double sum = 0.0;
        for (int i = 0; i < S; ++i) {
            const double e = energies[i];
            if (e < m) { m = e; l = i; }
            any_inf = any_inf || isinf(e);
            sum += e;
        }
        const double mean = sum / (double)S;
        double ss = 0.0;
        for (int i = 0; i < S; ++i) {
            const double d = energies[i] - mean;
            ss += d * d;
        }
        const double sd = sqrt(ss / (double)S);            // np.std: population standard deviation
        best_idx = l;
        status[0] = (!any_inf && sd <= atol + tol * fabs(mean)) ? 1 : 0;
        generation[0] += 1;
        status[1] = (int32_t)generation[0];
        status[2] = n_replaced;
        status[3] = l;
        best[0] = m;
        best[1] = sd;
        best[2] = mean;
    }
    __syncthreads();
    const int l = best_idx;                                // _promote_lowest_energy: swap rows 0 and l
    if (l != 0) {
        for (int j = threadIdx.x; j < P; j += blockDim.x) {
            const double a = pop[j], b = pop[l * P + j];
            pop[j] = b;
            pop[l * P + j] = a;
        }
        if (threadIdx.x == 0) {
            const double a = energies[0];
            energies[0] = energies[l];
            energies[l] = a;
        }
    }
}


}  // namespace de
}  // namespace cl
