// Camera noise profiles: the joint histogram of (uint8 mean-frame DN, frame DN) per channel
// over every frame of a static-scene video.  Replaces the per-frame, per-channel np.add.at scatter of
// compute_noise_profiles (modules/video_processing.py:77-106), whose result the camera's STD table is derived from
// (process_STD_data, :136-158).  Integer work: bit-exact.
//
// The full histogram is 256 x 256 x C counters (768 KB for RGB) -- too big for shared memory -- but a frame value
// sits within a few DN of its pixel's mean, so each CTA keeps a WINDOW histogram [c][mean][delta], delta = dn - mean
// + 8 in [0, 16), in 48 KB of shared memory (one shared-memory atomic per sample-frame, no global traffic); the rare
// values outside the window go straight to the global histogram, and the window is flushed once per CTA with 64-bit
// global atomics.  Each thread owns 4 consecutive samples (one 4-byte load per frame, 8 frames in flight) for all
// frames of the chunk, so the mean bytes are read once.  Measured: cfg4 (600 x 1080x1920x3) in 2.87 ms = 1.3 TB/s; the
// bound is the shared-memory atomic unit (~7 cycles per warp instruction).  A variant with per-thread private windows
// (plain LDS / add / STS, folded into the CTA histogram after the last frame) was slower, 3.55 ms: the read-modify-
// write chains of a thread cannot overlap, while the atomics are fire-and-forget.
#include "common.cuh"

namespace cl {
namespace {

constexpr int kThreads = 256;
constexpr int kWin = 16;          // window bins per (channel, mean)
constexpr int kHalf = 8;
constexpr int kFramesInFlight = 8;

template <int CT>     // compile-time channel count (1, 3) or 0 = run time
__global__ void __launch_bounds__(kThreads)
noise_profiles_kernel(const uint8_t* __restrict__ frames, int F, int64_t n, int C_rt,
                      const uint8_t* __restrict__ mean_u8, unsigned long long* __restrict__ hist) {
    extern __shared__ unsigned int whist[];                 // [C][256][kWin]
    const int C = CT ? CT : C_rt;
    for (int i = threadIdx.x; i < C * 256 * kWin; i += kThreads) whist[i] = 0u;
    __syncthreads();
    const int64_t n_vec = n >> 2;                            // 4-sample groups (the caller handles n % 4 and alignment)
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    const int64_t n4 = n >> 2;
    for (int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x; v < n_vec; v += stride) {
        const int64_t base = v << 2;
        const uint32_t m4 = __ldg(reinterpret_cast<const uint32_t*>(mean_u8) + v);
        uint32_t mean[4], wbase[4], ch[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            mean[j] = (m4 >> (8 * j)) & 0xFFu;
            ch[j] = (uint32_t)((base + j) % C);
            wbase[j] = (ch[j] * 256u + mean[j]) * kWin + kHalf - mean[j];      // + dn gives the window slot
        }
        const uint32_t* fp = reinterpret_cast<const uint32_t*>(frames) + v;
        int f0 = 0;
        auto one = [&](uint32_t q) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t dn = (q >> (8 * j)) & 0xFFu;
                const uint32_t delta = dn + kHalf - mean[j];                    // unsigned: out of window -> >= kWin
                if (delta < (uint32_t)kWin) atomicAdd(&whist[wbase[j] + dn], 1u);
                else atomicAdd(&hist[((uint64_t)mean[j] * 256u + dn) * C + ch[j]], 1ull);
            }
        };
        for (; f0 + kFramesInFlight <= F; f0 += kFramesInFlight) {
            uint32_t q[kFramesInFlight];
#pragma unroll
            for (int u = 0; u < kFramesInFlight; ++u) q[u] = __ldcs(fp + u * n4);
            fp += kFramesInFlight * n4;
#pragma unroll
            for (int u = 0; u < kFramesInFlight; ++u) one(q[u]);
        }
        for (; f0 < F; ++f0) {
            one(__ldcs(fp));
            fp += n4;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * 256 * kWin; i += kThreads) {
        const unsigned int cnt = whist[i];
        if (cnt == 0u) continue;
        const int c = i / (256 * kWin), m = (i / kWin) & 255, d = i % kWin;
        const int dn = m + d - kHalf;                        // in [0, 255] whenever cnt > 0
        atomicAdd(&hist[((uint64_t)m * 256u + (uint32_t)dn) * C + c], (unsigned long long)cnt);
    }
}

// ragged / unaligned remainder: one thread per sample, global atomics
__global__ void noise_profiles_tail_kernel(const uint8_t* __restrict__ frames, int F, int64_t n, int C, int64_t first,
                                           const uint8_t* __restrict__ mean_u8, unsigned long long* __restrict__ hist) {
    const int64_t i = first + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t m = mean_u8[i];
    const int c = (int)(i % C);
    for (int f = 0; f < F; ++f) {
        const uint32_t dn = frames[(int64_t)f * n + i];
        atomicAdd(&hist[((uint64_t)m * 256u + dn) * C + c], 1ull);
    }
}

}  // namespace
}  // namespace cl

extern "C" {

int cl_noise_profiles(const uint8_t* frames, int n_frames, int64_t n_samples, int channels, const uint8_t* mean_u8,
                      int64_t* hist, void* stream) {
    using namespace cl;
    CL_REQUIRE(n_frames >= 0 && n_samples >= 0 && channels >= 1 && channels <= CL_MAX_CHANNELS);
    if (n_frames == 0 || n_samples == 0) return CL_OK;
    CL_REQUIRE(frames && mean_u8 && hist && n_samples % channels == 0);
    cudaStream_t s = (cudaStream_t)stream;
    unsigned long long* h = reinterpret_cast<unsigned long long*>(hist);
    // the vector path needs 4-byte aligned rows: frame f starts at frames + f * n
    const bool vec_ok = aligned(frames, 4) && aligned(mean_u8, 4) && (n_samples % 4 == 0);
    int64_t done = 0;
    if (vec_ok) {
        const int64_t n_vec = n_samples / 4;
        const size_t smem = (size_t)channels * 256 * kWin * sizeof(unsigned int);
        int64_t blocks = (n_vec + kThreads - 1) / kThreads;
        const int64_t cap = (int64_t)sm_count() * 3;
        if (blocks > cap) blocks = cap;
        auto go = [&](auto kernel) -> int {
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return cuda_status(e);
            kernel<<<(unsigned)blocks, kThreads, smem, s>>>(frames, n_frames, n_samples, channels, mean_u8, h);
            return launched();
        };
        const int st = channels == 3 ? go(noise_profiles_kernel<3>)
                                     : channels == 1 ? go(noise_profiles_kernel<1>) : go(noise_profiles_kernel<0>);
        if (st != CL_OK) return st;
        done = n_samples;
    }
    if (done < n_samples) {
        const int64_t rest = n_samples - done;
        noise_profiles_tail_kernel<<<(unsigned)((rest + 255) / 256), 256, 0, s>>>(frames, n_frames, n_samples, channels,
                                                                               done, mean_u8, h);
        return launched();
    }
    return CL_OK;
}

}  // extern "C"
