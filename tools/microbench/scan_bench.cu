// Micro-benchmark for the dark-frame scan (development tool, not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/scan_bench tools/microbench/scan_bench.cu
// Variants: streaming floor (count only), direct bucket filing, shared-list filing; several grids.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#ifdef WITH_LIB
#include "hdr_merge.cuh"
#endif

constexpr int kThreads = 256;
constexpr int kBucketWords = 132;
constexpr int kBucketCap = 32;
constexpr int kTilePx = 512;

struct Params {
#ifdef BIG_PARAMS
    const void* pad0[64];          // mimic the library's 1.5 KB parameter block
    const uint8_t* dark[32];
    double pad1[32];
    uint32_t thr[32];
    uint8_t dark_k[32];
#else
    const uint8_t* dark[8];
    uint32_t thr[8];
#endif
    int n_dark;
    int64_t n;
    uint32_t* buckets;
    uint32_t* overflow;
};

__device__ __forceinline__ uint32_t bytes_ge(uint32_t x, uint32_t add, bool low) {
    const uint32_t lo = (x & 0x7f7f7f7fu) + add;
    return (low ? (x | lo) : (x & lo)) & 0x80808080u;
}

__device__ __forceinline__ void file_hit(const Params& p, int k, uint32_t sample) {
    const uint32_t px = sample / 3u, c = sample - px * 3u;
    const uint32_t tile = px / kTilePx;
    uint32_t* bucket = p.buckets + (size_t)tile * kBucketWords;
    const uint32_t slot = atomicAdd(bucket, 1u);
    if (slot < (uint32_t)kBucketCap) bucket[4 + 4 * slot] = (px - tile * kTilePx) | (c << 9) | ((uint32_t)k << 11);
    else atomicAdd(p.overflow, 1u);
}

// MODE 0: count only; 1: direct filing; 2: shared list
template <int MODE, int VECS>
__global__ void __launch_bounds__(kThreads) scan_units(const __grid_constant__ Params p) {
    __shared__ uint2 hits[2048];
    __shared__ uint32_t n_hits;
    if (threadIdx.x == 0) n_hits = 0;
    __syncthreads();
    const uint32_t n_vec = (uint32_t)(p.n / 16);
    constexpr uint32_t kUnit = kThreads * VECS;
    const uint32_t upf = (n_vec + kUnit - 1) / kUnit;
    const uint32_t total = upf * p.n_dark;
    uint32_t local = 0;
    for (uint32_t unit = blockIdx.x; unit < total; unit += gridDim.x) {
        const uint32_t j = unit / upf;
        const uint32_t base = (unit - j * upf) * kUnit + threadIdx.x;
#ifdef BIG_PARAMS
        const int kk = p.dark_k[j];
#else
        const int kk = j;
#endif
        const uint4* src = reinterpret_cast<const uint4*>(p.dark[kk]);
        const uint32_t thr = p.thr[kk];
        const bool low = thr <= 128u;
        const uint32_t add = (low ? 128u - thr : 256u - thr) * 0x01010101u;
        uint4 q[VECS];
#pragma unroll
        for (int u = 0; u < VECS; ++u) {
            const uint32_t v = base + u * kThreads;
            q[u] = v < n_vec ? __ldg(src + v) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < VECS; ++u) {
            const uint32_t v = base + u * kThreads;
            const uint32_t h[4] = {bytes_ge(q[u].x, add, low), bytes_ge(q[u].y, add, low),
                                   bytes_ge(q[u].z, add, low), bytes_ge(q[u].w, add, low)};
            if ((h[0] | h[1] | h[2] | h[3]) == 0u) continue;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                uint32_t m = h[w];
                while (m) {
                    const int b = (__ffs(m) - 1) >> 3;
                    const uint32_t sample = v * 16 + w * 4 + b;
                    if (MODE == 0) ++local;
                    else if (MODE == 1) file_hit(p, j, sample);
                    else {
                        const uint32_t slot = atomicAdd(&n_hits, 1u);
                        if (slot < 2048u) hits[slot] = make_uint2(sample, j);
                        else file_hit(p, j, sample);
                    }
                    m &= m - 1;
                }
            }
        }
    }
    if (MODE == 0) { if (local) atomicAdd(p.overflow + 1, local); }
    if (MODE == 2) {
        __syncthreads();
        const uint32_t parked = min(n_hits, 2048u);
        for (uint32_t e = threadIdx.x; e < parked; e += kThreads) file_hit(p, (int)hits[e].y, hits[e].x);
    }
}

// the committed (per-frame, grid-stride) structure for reference
__global__ void __launch_bounds__(kThreads) scan_frames(const __grid_constant__ Params p) {
    const int64_t n_vec = p.n / 16;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int k0 = 0; k0 < p.n_dark; ++k0) {
#ifdef BIG_PARAMS
        const int k = p.dark_k[k0];
#else
        const int k = k0;
#endif
        const uint32_t thr = p.thr[k] * 0x01010101u;
        const uint4* src = reinterpret_cast<const uint4*>(p.dark[k]);
        for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < n_vec; base += stride * 4) {
            uint4 q[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t v = base + threadIdx.x + u * stride;
                q[u] = v < n_vec ? __ldg(src + v) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t v = base + threadIdx.x + u * stride;
                if (v < n_vec) {
                    const uint32_t w[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint32_t m = __vcmpgeu4(w[j], thr);
                        while (m) {
                            const int b = (__ffs(m) - 1) >> 3;
                            file_hit(p, k, (uint32_t)(v * 16 + j * 4 + b));
                            m &= ~(0xFFu << (8 * b));
                        }
                    }
                }
            }
        }
    }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

int main() {
    const int64_t n = 2160LL * 3840 * 3;
    const int n_dark = 7;
    Params p{};
    p.n = n; p.n_dark = n_dark;
    std::vector<uint8_t> host(n);
    uint64_t s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    for (int k = 0; k < n_dark; ++k) {
        for (int64_t i = 0; i < n; ++i) {
            const uint64_t r = rnd();
            host[i] = (r % 1000 == 0) ? 40 + (r >> 20) % 160 : (r >> 10) % 6;
        }
        uint8_t* d; CK(cudaMalloc(&d, n));
        CK(cudaMemcpy(d, host.data(), n, cudaMemcpyHostToDevice));
#ifdef BIG_PARAMS
        p.dark[9 + k] = d; p.thr[9 + k] = 13; p.dark_k[k] = 9 + k;
#else
        p.dark[k] = d; p.thr[k] = 13;
#endif
    }
    const int64_t tiles = n / 3 / kTilePx;
    CK(cudaMalloc(&p.buckets, tiles * kBucketWords * 4));
    CK(cudaMalloc(&p.overflow, 16));
    // L2 flush buffer
    void* flush; CK(cudaMalloc(&flush, 512 << 20));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    auto run = [&](const char* name, auto launch) {
        float best = 1e9f, sum = 0;
        uint32_t cnt[4] = {0, 0, 0, 0};
        for (int it = 0; it < 6; ++it) {
            cudaMemset(p.buckets, 0, tiles * kBucketWords * 4);
            cudaMemset(p.overflow, 0, 16);
            cudaMemset(flush, it, 512 << 20);
            cudaEventRecord(e0);
            launch();
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (it) { best = ms < best ? ms : best; sum += ms; }
        }
        cudaMemcpy(cnt, p.overflow, 16, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaGetLastError();
        printf("%-34s best %7.1f us  avg %7.1f us  %6.2f TB/s  (overflow %u, count %u) %s\n", name, best * 1e3, sum / 5 * 1e3,
               n_dark * n / (best * 1e-3) / 1e12, cnt[0], cnt[1], e == cudaSuccess ? "" : cudaGetErrorString(e));
    };
#ifdef WITH_LIB
    {
        cl::MergeParams mp;
        memset(&mp, 0, sizeof(mp));
        mp.n = 16; mp.H = 2160; mp.W = 3840; mp.C = 3; mp.bits = 256; mp.K = 3; mp.max_dn = 255.0;
        for (int k = 0; k < 16; ++k) mp.hot_dn[k] = 0xFFFFFFFFu;
        for (int k = 0; k < n_dark; ++k) {
            mp.dark[9 + k] = p.dark[k]; mp.hot_dn[9 + k] = 13; mp.dark_k[k] = 9 + k;
        }
        mp.n_dark = n_dark; mp.any_dark = 1;
        mp.n_full_tiles = (int)tiles; mp.hot_cap = 65536;
        uint32_t* ws; CK(cudaMalloc(&ws, (tiles * 4 + 4 + 65536 + tiles * 128) * 4));
        mp.bucket_counts = ws; mp.hot_list = ws + tiles * 4; mp.bucket_entries = mp.hot_list + 4 + 65536;
        run("library launch_dark_scan", [&] { cl::launch_dark_scan(mp, 0); });
    }
#endif
    for (int per_sm : {4, 6, 8}) {
        const int grid = sms * per_sm;
        char nm[64];
        snprintf(nm, 64, "frames(committed) grid=%dxSM", per_sm); run(nm, [&] { scan_frames<<<grid, kThreads>>>(p); });
        snprintf(nm, 64, "units count-only V4 grid=%dxSM", per_sm); run(nm, [&] { scan_units<0, 4><<<grid, kThreads>>>(p); });
        snprintf(nm, 64, "units count-only V8 grid=%dxSM", per_sm); run(nm, [&] { scan_units<0, 8><<<grid, kThreads>>>(p); });
        snprintf(nm, 64, "units direct V4 grid=%dxSM", per_sm); run(nm, [&] { scan_units<1, 4><<<grid, kThreads>>>(p); });
        snprintf(nm, 64, "units direct V8 grid=%dxSM", per_sm); run(nm, [&] { scan_units<1, 8><<<grid, kThreads>>>(p); });
        snprintf(nm, 64, "units shared V4 grid=%dxSM", per_sm); run(nm, [&] { scan_units<2, 4><<<grid, kThreads>>>(p); });
        snprintf(nm, 64, "units shared V8 grid=%dxSM", per_sm); run(nm, [&] { scan_units<2, 8><<<grid, kThreads>>>(p); });
    }
    return 0;
}
