// Egress: the 8-bit export of ImageSet.save_8bit (image_set.py:321-358) on the
// device, so that a float64 result leaves the GPU as 1 byte per sample instead of 8:
//     max_float = np.amax(val);  if max_float > 1: val /= max_float
//     val = np.around(val * MAX_DN).astype(uint8)
// Two launches: block-wise max (order-independent, so exact; a NaN anywhere makes the max NaN, as
// np.amax does, and then no normalisation happens) and the quantisation itself.  Division, product and
// round-half-even are the IEEE operations NumPy performs, so the bytes are identical.
#include "common.cuh"

namespace cl {
namespace {

constexpr int kThreads = 256;
constexpr int kBlocks = 592;      // fixed: 148 SMs x 4 resident blocks
constexpr int kVec = 4;           // samples per thread per step (uchar4 store)

__global__ void __launch_bounds__(kThreads)
max_partial_kernel(const double* __restrict__ val, int64_t n, double* __restrict__ partial /* [kBlocks][2] */) {
    __shared__ double smax[kThreads / 32];
    __shared__ int snan[kThreads / 32];
    const double ninf = __longlong_as_double(0xfff0000000000000LL);
    double m = ninf;
    int has_nan = 0;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        const double a = __ldg(val + i), b = __ldg(val + i + stride), c = __ldg(val + i + 2 * stride),
                     d = __ldg(val + i + 3 * stride);
        has_nan |= (a != a) | (b != b) | (c != c) | (d != d);
        m = fmax(fmax(m, a), fmax(fmax(b, c), d));        // fmax drops NaNs; the flag keeps them
    }
    for (; i < n; i += stride) {
        const double a = __ldg(val + i);
        has_nan |= a != a;
        m = fmax(m, a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        has_nan |= __shfl_xor_sync(0xffffffffu, has_nan, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { smax[warp] = m; snan[warp] = has_nan; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kThreads / 32; ++w) { m = fmax(m, smax[w]); has_nan |= snan[w]; }
        partial[2 * blockIdx.x] = m;
        partial[2 * blockIdx.x + 1] = has_nan ? 1.0 : 0.0;
    }
}

// (uint8) of an integral double the way NumPy's C cast does on x86-64: through a 32-bit integer, low byte
__device__ __forceinline__ uint32_t to_u8(double r) { return (uint32_t)__double2int_rz(r) & 0xFFu; }

__global__ void __launch_bounds__(kThreads)
quantize_kernel(const double* __restrict__ val, int64_t n, double max_dn, const double* __restrict__ partial,
                uint8_t* __restrict__ out, double* __restrict__ out_max) {
    __shared__ double smax;
    __shared__ int snan;
    if (threadIdx.x < 32) {                                  // every block folds the 592 partial maxima
        double m = __longlong_as_double(0xfff0000000000000LL);
        int has_nan = 0;
        for (int b = threadIdx.x; b < kBlocks; b += 32) {
            m = fmax(m, partial[2 * b]);
            has_nan |= partial[2 * b + 1] != 0.0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
            has_nan |= __shfl_xor_sync(0xffffffffu, has_nan, o);
        }
        if (threadIdx.x == 0) { smax = m; snan = has_nan; }
    }
    __syncthreads();
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const double mx = snan ? nan : smax;
    if (out_max && blockIdx.x == 0 && threadIdx.x == 0) *out_max = mx;
    const bool scale = mx > 1.0;                             // false for NaN, like the reference's `if`
    auto q = [&](double v) {
        if (scale) v = __ddiv_rn(v, mx);
        return to_u8(rint(__dmul_rn(v, max_dn)));            // np.around == round-half-even
    };
    const int64_t n4 = n / kVec;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(val) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 3) == 0);
    if (vec_ok) {
        for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += stride) {
            const double2 a = __ldcs(reinterpret_cast<const double2*>(val) + 2 * i);
            const double2 b = __ldcs(reinterpret_cast<const double2*>(val) + 2 * i + 1);
            const uint32_t w = q(a.x) | (q(a.y) << 8) | (q(b.x) << 16) | (q(b.y) << 24);
            reinterpret_cast<uint32_t*>(out)[i] = w;
        }
        for (int64_t i = n4 * kVec + (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride)
            out[i] = (uint8_t)q(val[i]);
    } else {
        for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) out[i] = (uint8_t)q(val[i]);
    }
}

}  // namespace
}  // namespace cl

extern "C" {

size_t cl_quantize_8bit_workspace_bytes(void) { return (size_t)cl::kBlocks * 2 * sizeof(double); }

int cl_quantize_8bit(const double* val, int64_t n, double max_dn, uint8_t* out, double* out_max,
                     void* workspace, size_t workspace_bytes, void* stream) {
    using namespace cl;
    CL_REQUIRE(n >= 0);
    if (n == 0) return CL_OK;                    // np.amax of an empty array raises; the host mirror does too
    CL_REQUIRE(val && out);
    if (!workspace || workspace_bytes < cl_quantize_8bit_workspace_bytes()) return CL_ERR_WORKSPACE;
    if (!aligned(workspace, 8)) return CL_ERR_ALIGNMENT;
    cudaStream_t s = (cudaStream_t)stream;
    double* partial = reinterpret_cast<double*>(workspace);
    max_partial_kernel<<<kBlocks, kThreads, 0, s>>>(val, n, partial);
    int st = launched();
    if (st != CL_OK) return st;
    quantize_kernel<<<kBlocks, kThreads, 0, s>>>(val, n, max_dn, partial, out, out_max);
    return launched();
}

}  // extern "C"
