// Differential evolution on the device: the population, mutation / crossover
// (strategy 'currenttobest1bin'), the bound repair and the selection of one generation run on the GPU, so
// a generation of the ICRF calibration is  cl_de_trial -> [K4: curves, energy partial, (all-reduce),
// finalize] -> cl_de_select  with no host round trip; the host reads a 32-byte status block every few
// generations.  Spec: scipy.optimize.differential_evolution as the reference calls it
// (ICRF_calibration_exposure.py:357-361: strategy 'currenttobest1bin', mutation (0, 1.95) -> dither,
// recombination 0.4, tol 0.01) in its vectorized 'deferred' updating form:
//   scale  ~ U(dither)                       once per generation
//   b'     = x_i + scale * (x_best - x_i + x_r0 - x_r1),  r0 != r1, both != i   (population in [0,1]^P)
//   trial  = where(U < CR or j == fill_point, b', x_i);  components outside [0,1] are redrawn ~ U(0,1)
//   x_i    = trial_i  if  E(trial_i) < E(x_i);  the best member is swapped into row 0
//   converged  <=>  no inf energy and std(E) <= atol + tol * |mean(E)|
// SciPy draws from a NumPy Generator; here every draw is a counter-based splitmix64 value keyed by
// (seed, generation, candidate, slot), so the stream does not depend on the launch geometry and
// oracle/de.py reproduces it bit for bit.  The search is therefore statistically, not bitwise, the one
// SciPy would run (tests compare both drivers' optima).
#include "de_common.cuh"

namespace cl {
namespace {

// one thread per (candidate, parameter)
__global__ void de_trial_kernel(const double* __restrict__ pop, int S, int P, const de::TrialConfig cfg,
                                const int64_t* __restrict__ generation, const double* __restrict__ lo,
                                const double* __restrict__ hi, double* __restrict__ trial, double* __restrict__ params) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= S * P) return;
    const int i = t / P, j = t - i * P;
    de::trial_component(pop, S, P, cfg, (uint64_t)generation[0], i, j, lo, hi, trial[t], params[t]);
}

__global__ void de_select_kernel(double* __restrict__ pop, double* __restrict__ energies,
                                 const double* __restrict__ trial, const double* __restrict__ trial_energies, int S,
                                 int P, double tol, double atol, int64_t* __restrict__ generation,
                                 int32_t* __restrict__ status, double* __restrict__ best) {
    de::select_block(pop, energies, trial, trial_energies, S, P, tol, atol, generation, status, best);
}

}  // namespace
}  // namespace cl

extern "C" {

int cl_de_trial(const double* pop, int n_members, int n_params, double dither_lo, double dither_hi, double crossover,
                uint64_t seed, const int64_t* generation, const double* lower, const double* upper, double* trial,
                double* params, void* stream) {
    using namespace cl;
    CL_REQUIRE(pop && generation && lower && upper && trial && params);
    CL_REQUIRE(n_members >= 4 && n_params >= 1 && n_params <= 64);
    CL_REQUIRE(dither_lo <= dither_hi && crossover >= 0.0 && crossover <= 1.0);
    const int n = n_members * n_params;
    const de::TrialConfig cfg{dither_lo, dither_hi, crossover, seed};
    de_trial_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(pop, n_members, n_params, cfg, generation, lower,
                                                                      upper, trial, params);
    return launched();
}

int cl_de_select(double* pop, double* energies, const double* trial, const double* trial_energies, int n_members,
                 int n_params, double tol, double atol, int64_t* generation, int32_t* status, double* best,
                 void* stream) {
    using namespace cl;
    CL_REQUIRE(pop && energies && trial && trial_energies && generation && status && best);
    CL_REQUIRE(n_members >= 4 && n_params >= 1 && n_params <= 64);
    de_select_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(pop, energies, trial, trial_energies, n_members, n_params, tol,
                                                         atol, generation, status, best);
    return launched();
}

}  // extern "C"
