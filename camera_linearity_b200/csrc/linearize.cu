// K1: ICRF linearisation (LUT gather + derivative for uncertainty propagation).
// Replaces measurand.py:471-541 of the reference (see include/camera_linearity.h).
//
// HBM-bound streaming kernel: each thread owns 8 strided PAIRS of samples so that every warp-level
// load and store is contiguous (512 B of float64 per instruction), the per-channel LUTs live in
// shared memory ([bits][C] doubles, 12 KB for 8-bit RGB) and outputs are written as 16-byte vectors.
// Algorithmic bytes per sample: b_dn + 8 (out) [+ 8 std in + 8 std out].
#include "common.cuh"

namespace cl {

std::atomic<uint64_t> g_launch_count{0};

namespace {

constexpr int kThreads = 256;

// SRC: 0 = uint8 DN, 1 = uint16 DN, 2 = float64 value in [0, 1]
template <int SRC>
struct Src;
template <>
struct Src<0> {
    using T = uint8_t;
    using Pair = uchar2;
};
template <>
struct Src<1> {
    using T = uint16_t;
    using Pair = ushort2;
};
template <>
struct Src<2> {
    using T = double;
    using Pair = double2;
};

constexpr int kSlots = 8;   // pairs of samples per thread; slot u of thread t = pair (u*blockDim + t)

template <int SRC>
__device__ __forceinline__ uint32_t bin_of(typename Src<SRC>::T v, double max_dn, uint32_t wrap_mask) {
    if (SRC == 2) return wrap_bin(__dmul_rn((double)v, max_dn), wrap_mask);
    return (uint32_t)v;
}

// Every warp-level access is contiguous: in slot u the 32 lanes read 32 consecutive sample PAIRS
// (64 B of uint8, 512 B of float64) and write 512 B -- fully coalesced loads and stores.
template <int SRC, bool LUT_SMEM>
__global__ void __launch_bounds__(kThreads, SRC == 2 ? 2 : 3)
linearize_kernel(const typename Src<SRC>::T* __restrict__ src, double max_dn, uint32_t wrap_mask,
                 const double* __restrict__ std_in, const double* __restrict__ lut,
                 const double* __restrict__ dlut, double* __restrict__ out_val,
                 double* __restrict__ out_std, uint16_t* __restrict__ bin_out, int64_t n, int C,
                 int bits) {
    using T = typename Src<SRC>::T;
    using Pair = typename Src<SRC>::Pair;
    extern __shared__ double smem[];
    const double* tv = lut;
    const double* td = dlut;
    if (LUT_SMEM) {
        const int rows = bits * C;
        for (int i = threadIdx.x; i < rows; i += blockDim.x) {
            smem[i] = lut[i];
            if (dlut) smem[rows + i] = dlut[i];
        }
        __syncthreads();
        tv = smem;
        td = smem + rows;
    }
    const bool use_std = (std_in != nullptr) && (dlut != nullptr) && (out_std != nullptr);
    const int64_t n_pairs = n / 2;
    const int64_t per_block = (int64_t)kSlots * blockDim.x;
    const int64_t n_blocks_work = (n_pairs + per_block - 1) / per_block;
    // channel of the first sample of a pair, kept incrementally (a 64-bit modulo per pair cost more
    // registers and instructions than the rest of the loop): c(q + d) = (c(q) + 2d) mod C
    const int step_slot = (int)((2 * (int64_t)blockDim.x) % C);
    const int step_block = (int)((2 * per_block * gridDim.x) % C);
    int c_first = (int)((2 * ((int64_t)blockIdx.x * per_block + threadIdx.x)) % C);
    for (int64_t b = blockIdx.x; b < n_blocks_work; b += gridDim.x) {
        const int64_t first = b * per_block + threadIdx.x;
        Pair in[kSlots];
        double2 sd[kSlots];
#pragma unroll
        for (int u = 0; u < kSlots; ++u) {
            const int64_t q = first + (int64_t)u * blockDim.x;
            if (q < n_pairs) {
                in[u] = reinterpret_cast<const Pair*>(src)[q];
                if (use_std) sd[u] = reinterpret_cast<const double2*>(std_in)[q];
            }
        }
        int c0 = c_first;
#pragma unroll
        for (int u = 0; u < kSlots; ++u) {
            const int64_t q = first + (int64_t)u * blockDim.x;
            const int c_here = c0;
            c0 += step_slot;
            if (c0 >= C) c0 -= C;
            if (q < n_pairs) {
                const uint32_t b0 = bin_of<SRC>(in[u].x, max_dn, wrap_mask);
                const uint32_t b1 = bin_of<SRC>(in[u].y, max_dn, wrap_mask);
                const int c0 = c_here;
                const int c1 = (c0 + 1 >= C) ? c0 + 1 - C : c0 + 1;
                const int i0 = (int)b0 * C + c0, i1 = (int)b1 * C + c1;
                reinterpret_cast<double2*>(out_val)[q] = make_double2(tv[i0], tv[i1]);
                if (use_std)
                    reinterpret_cast<double2*>(out_std)[q] =
                        make_double2(__dmul_rn(td[i0], sd[u].x), __dmul_rn(td[i1], sd[u].y));
                if (bin_out) reinterpret_cast<ushort2*>(bin_out)[q] = make_ushort2((uint16_t)b0, (uint16_t)b1);
            }
        }
        c_first += step_block;
        if (c_first >= C) c_first -= C;
    }
    // odd tail sample
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const int64_t i = n - 1;
        const uint32_t b = bin_of<SRC>(src[i], max_dn, wrap_mask);
        const int idx = (int)b * C + (int)(i % C);
        out_val[i] = tv[idx];
        if (use_std) out_std[i] = __dmul_rn(td[idx], std_in[i]);
        if (bin_out) bin_out[i] = (uint16_t)b;
    }
}

template <int SRC>
int launch(const void* src, double max_dn, const double* std_in, const double* lut,
           const double* dlut, double* out_val, double* out_std, uint16_t* bin_out, int64_t n,
           int C, int bits, uint32_t wrap_mask, cudaStream_t stream) {
    using T = typename Src<SRC>::T;
    if (n == 0) return CL_OK;
    // pair accesses need the natural alignment of a pair
    if (!aligned(src, sizeof(T) * 2) || !aligned(out_val, 16) || (std_in && !aligned(std_in, 16)) ||
        (out_std && !aligned(out_std, 16)) || (bin_out && !aligned(bin_out, 4)))
        return CL_ERR_ALIGNMENT;
    const size_t lut_bytes = (size_t)bits * C * sizeof(double) * 2;
    const bool lut_smem = lut_bytes <= 96 * 1024;
    const int64_t per_block = (int64_t)kSlots * kThreads;
    int64_t blocks = (n / 2 + per_block - 1) / per_block;
    const int64_t cap = (int64_t)sm_count() * (SRC == 2 ? 2 : 3);   // one wave of resident CTAs
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (lut_smem) {
        auto k = linearize_kernel<SRC, true>;
        if (lut_bytes > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)lut_bytes);
            if (e != cudaSuccess) return cuda_status(e);
        }
        k<<<(unsigned)blocks, kThreads, lut_bytes, stream>>>(
            (const T*)src, max_dn, wrap_mask, std_in, lut, dlut, out_val, out_std, bin_out, n, C,
            bits);
    } else {
        linearize_kernel<SRC, false><<<(unsigned)blocks, kThreads, 0, stream>>>(
            (const T*)src, max_dn, wrap_mask, std_in, lut, dlut, out_val, out_std, bin_out, n, C,
            bits);
    }
    return launched();
}

}  // namespace
}  // namespace cl

extern "C" {

int cl_abi_version(void) { return CL_ABI_VERSION; }

uint64_t cl_launch_count(void) { return cl::g_launch_count.load(); }

const char* cl_status_string(int status) {
    switch (status) {
        case CL_OK: return "ok";
        case CL_ERR_INVALID_ARGUMENT: return "invalid argument";
        case CL_ERR_UNSUPPORTED: return "unsupported configuration";
        case CL_ERR_WORKSPACE: return "workspace missing or too small";
        case CL_ERR_ALIGNMENT: return "pointer not sufficiently aligned";
        default: break;
    }
    if (status <= CL_ERR_CUDA) return cudaGetErrorString((cudaError_t)(CL_ERR_CUDA - status));
    return "unknown status";
}

int cl_linearize_dn(const void* dn, int dn_bytes, const double* std_in, const double* lut,
                    const double* dlut, double* out_val, double* out_std, int64_t n_samples,
                    int channels, int bits, void* stream) {
    CL_REQUIRE(n_samples >= 0 && channels >= 1 && channels <= CL_MAX_CHANNELS && bits >= 1);
    if (n_samples == 0) return CL_OK;
    CL_REQUIRE(dn && lut && out_val);
    CL_REQUIRE(dn_bytes == 1 || dn_bytes == 2);
    if ((dn_bytes == 1 && bits < 256) || (dn_bytes == 2 && bits < 65536)) return CL_ERR_INVALID_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream;
    if (dn_bytes == 1)
        return cl::launch<0>(dn, 255.0, std_in, lut, dlut, out_val, out_std, nullptr, n_samples,
                             channels, bits, 0xFFu, s);
    return cl::launch<1>(dn, 65535.0, std_in, lut, dlut, out_val, out_std, nullptr, n_samples,
                         channels, bits, 0xFFFFu, s);
}

int cl_linearize_f64(const double* val, double max_dn, const double* std_in, const double* lut,
                     const double* dlut, double* out_val, double* out_std, uint16_t* bin_out,
                     int64_t n_samples, int channels, int bits, void* stream) {
    CL_REQUIRE(n_samples >= 0 && channels >= 1 && channels <= CL_MAX_CHANNELS);
    if (n_samples == 0) return CL_OK;
    CL_REQUIRE(val && lut && out_val);
    // the wrapping cast (uint8 for 8-bit data as in measurand.py:503, uint16 beyond) yields bins
    // in [0, 256) or [0, 65536): the LUT must cover them
    const uint32_t wrap_mask = max_dn <= 255.0 ? 0xFFu : 0xFFFFu;
    CL_REQUIRE(bits >= (int)wrap_mask + 1);
    return cl::launch<2>(val, max_dn, std_in, lut, dlut, out_val, out_std, bin_out, n_samples,
                         channels, bits, wrap_mask, (cudaStream_t)stream);
}

}  // extern "C"
