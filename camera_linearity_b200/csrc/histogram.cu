// Per-channel histogram of a Measurand: the array part of
// compute_channel_histogram (measurand.py:430-469) = np.histogram over the finite values of one channel,
// optionally weighted by 1/sigma (samples with sigma == 0 are dropped).  Bin assignment follows NumPy's
// uniform-bin fast path operation for operation (subtract, divide, multiply, truncate, then the two
// edge corrections against the np.linspace edges the caller passes), so counts are exact; weighted sums
// are float64 atomics (order differs from np.bincount: ~1e-15 relative).
#include "common.cuh"

namespace cl {
namespace {

constexpr int kThreads = 256;
constexpr int kSmemBins = 4096;       // 32 KB of double bins per block; more bins go straight to global memory

__global__ void __launch_bounds__(kThreads)
histogram_kernel(const double* __restrict__ val, const double* __restrict__ std, int64_t n_px, int C, int c, int bins,
                 double first_edge, double last_edge, const double* __restrict__ edges, double* __restrict__ hist) {
    extern __shared__ double sh[];
    const bool use_smem = bins <= kSmemBins;
    if (use_smem) {
        for (int b = threadIdx.x; b < bins; b += kThreads) sh[b] = 0.0;
        __syncthreads();
    }
    const double denom = __dsub_rn(last_edge, first_edge);
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n_px; i += stride) {
        const double v = val[i * C + c];
        if (!isfinite(v)) continue;                                   // finite_mask, measurand.py:452
        double w = 1.0;
        if (std) {
            const double s = std[i * C + c];
            if (s == 0.0) continue;                                   // non_zeros, :456
            w = __ddiv_rn(1.0, s);                                    // weights = 1 / stds, :459
        }
        if (!(v >= first_edge && v <= last_edge)) continue;           // keep, np.histogram
        int idx = (int)(__dmul_rn(__ddiv_rn(__dsub_rn(v, first_edge), denom), (double)bins));
        if (idx == bins) idx = bins - 1;
        if (v < edges[idx]) --idx;                                    // decrement
        else if (idx != bins - 1 && v >= edges[idx + 1]) ++idx;       // increment (exclusive of the last bin)
        if (use_smem) atomicAdd(&sh[idx], w);
        else atomicAdd(&hist[idx], w);
    }
    if (use_smem) {
        __syncthreads();
        for (int b = threadIdx.x; b < bins; b += kThreads)
            if (sh[b] != 0.0) atomicAdd(&hist[b], sh[b]);
    }
}

}  // namespace
}  // namespace cl

extern "C" int cl_channel_histogram(const double* val, const double* std, int64_t n_pixels, int channels, int channel,
                                    int bins, double first_edge, double last_edge, const double* edges, double* hist,
                                    void* stream) {
    using namespace cl;
    CL_REQUIRE(n_pixels >= 0 && channels >= 1 && channel >= 0 && channel < channels && bins >= 1);
    CL_REQUIRE(hist && edges && (val || n_pixels == 0));
    CL_REQUIRE(last_edge > first_edge);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(hist, 0, (size_t)bins * sizeof(double), s);
    if (e != cudaSuccess) return cuda_status(e);
    if (n_pixels == 0) return CL_OK;
    int64_t blocks = (n_pixels + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)sm_count() * 4;
    if (blocks > cap) blocks = cap;
    const size_t smem = bins <= kSmemBins ? (size_t)bins * sizeof(double) : 0;
    histogram_kernel<<<(unsigned)blocks, kThreads, smem, s>>>(val, std, n_pixels, channels, channel, bins, first_edge,
                                                             last_edge, edges, hist);
    return launched();
}
