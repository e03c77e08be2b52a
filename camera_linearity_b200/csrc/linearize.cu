// K1: ICRF linearisation (LUT gather + derivative for uncertainty propagation).
// Replaces measurand.py:471-541 of the reference (see include/camera_linearity.h).
//
// HBM-bound streaming kernel: each thread owns VEC consecutive samples, loaded with one
// 16/8-byte vector load (uint8 x16, uint16 x8, f64 x2), the per-channel LUTs live in shared
// memory ([bits][C] doubles, 12 KB for 8-bit RGB) and outputs are written as 16-byte vectors.
// Algorithmic bytes per sample: b_dn + 8 (out) [+ 8 std in + 8 std out].
#include "common.cuh"

namespace cl {

std::atomic<uint64_t> g_launch_count{0};

namespace {

constexpr int kThreads = 256;

template <typename T, int VEC>
struct alignas(sizeof(T) * VEC) Pack {
    T v[VEC];
};

// SRC: 0 = uint8 DN, 1 = uint16 DN, 2 = float64 value in [0, 1]
template <int SRC>
struct Src;
template <>
struct Src<0> {
    using T = uint8_t;
    static constexpr int VEC = 16;
};
template <>
struct Src<1> {
    using T = uint16_t;
    static constexpr int VEC = 8;
};
template <>
struct Src<2> {
    using T = double;
    static constexpr int VEC = 2;
};

template <int SRC, bool LUT_SMEM>
__global__ void __launch_bounds__(kThreads)
linearize_kernel(const typename Src<SRC>::T* __restrict__ src, double max_dn, uint32_t wrap_mask,
                 const double* __restrict__ std_in, const double* __restrict__ lut,
                 const double* __restrict__ dlut, double* __restrict__ out_val,
                 double* __restrict__ out_std, uint16_t* __restrict__ bin_out, int64_t n, int C,
                 int bits) {
    using T = typename Src<SRC>::T;
    constexpr int VEC = Src<SRC>::VEC;
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
    const int64_t n_vec = n / VEC;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n_vec; v += stride) {
        const int64_t base = v * VEC;
        const Pack<T, VEC> in = *reinterpret_cast<const Pack<T, VEC>*>(src + base);
        int c = (int)(base % C);
        uint32_t bin[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            if (SRC == 2)
                bin[j] = wrap_bin(__dmul_rn((double)in.v[j], max_dn), wrap_mask);
            else
                bin[j] = (uint32_t)in.v[j];
        }
        double ov[VEC], os[VEC];
        Pack<double, 2> sd[VEC / 2];
        if (use_std) {
#pragma unroll
            for (int j = 0; j < VEC / 2; ++j)
                sd[j] = *reinterpret_cast<const Pack<double, 2>*>(std_in + base + 2 * j);
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const int idx = (int)bin[j] * C + c;
            ov[j] = tv[idx];
            if (use_std) os[j] = __dmul_rn(td[idx], sd[j / 2].v[j & 1]);
            c = (c + 1 == C) ? 0 : c + 1;
        }
#pragma unroll
        for (int j = 0; j < VEC / 2; ++j) {
            Pack<double, 2> o;
            o.v[0] = ov[2 * j];
            o.v[1] = ov[2 * j + 1];
            *reinterpret_cast<Pack<double, 2>*>(out_val + base + 2 * j) = o;
            if (use_std) {
                o.v[0] = os[2 * j];
                o.v[1] = os[2 * j + 1];
                *reinterpret_cast<Pack<double, 2>*>(out_std + base + 2 * j) = o;
            }
        }
        if (bin_out) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) bin_out[base + j] = (uint16_t)bin[j];
        }
    }
    // ragged tail (n % VEC samples), one thread each
    const int64_t tail0 = n_vec * VEC;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < n - tail0) {
        const int64_t i = tail0 + gid;
        uint32_t b;
        if (SRC == 2)
            b = wrap_bin(__dmul_rn((double)src[i], max_dn), wrap_mask);
        else
            b = (uint32_t)src[i];
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
    constexpr int VEC = Src<SRC>::VEC;
    if (n == 0) return CL_OK;
    // vector accesses need the natural alignment of the vectors
    if (!aligned(src, sizeof(T) * VEC) || !aligned(out_val, 16) ||
        (std_in && !aligned(std_in, 16)) || (out_std && !aligned(out_std, 16)))
        return CL_ERR_ALIGNMENT;
    const size_t lut_bytes = (size_t)bits * C * sizeof(double) * 2;
    const bool lut_smem = lut_bytes <= 96 * 1024;
    const int64_t n_vec = (n + VEC - 1) / VEC;
    int64_t blocks = (n_vec + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)sm_count() * 8;  // 8 resident CTAs of 256 threads per SM
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
