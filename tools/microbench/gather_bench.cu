// Micro-benchmark: random 16-byte table gathers on sm_100a (development tool, not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/gather_bench tools/microbench/gather_bench.cu
// How many random gathers per SM clock does each path sustain for a 1 MB table (65536 x double2)?
//   l2     : __ldg from global memory (L1 miss -> L2 hit), the path of merge_wide_kernel in round 1
//   dsmem  : the table spread over the shared memory of an 8-CTA cluster, ld.shared::cluster
//   mixed  : every other gather through each path (both pipes busy)
//   local  : a 128 KB slice in the CTA's own shared memory (upper bound of a shared-memory gather)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

constexpr int kRows = 65536;
constexpr int kCluster = 8;
constexpr int kRowsPerCta = kRows / kCluster;        // 8192 rows x 16 B = 128 KB
constexpr int kThreads = 512;

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

__device__ __forceinline__ double2 ld_cluster(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}

// ---- asynchronous variants: the gathered rows land in shared memory, no register / scoreboard per load ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// MODE 0: cp.async.cg 16 B (LDGSTS) per gather, committed in groups;  MODE 1: cp.async.bulk 16 B (TMA) per gather,
// completion counted by an mbarrier per warp-round
template <int MODE, bool HOT>
__global__ void __launch_bounds__(kThreads, 1)
async_gather_kernel(const double2* __restrict__ table, int iters, double* __restrict__ sink) {
    constexpr int U = 8;
    extern __shared__ __align__(16) unsigned char dyn[];
    double2 (*land)[U][kThreads] = reinterpret_cast<double2 (*)[U][kThreads]>(dyn);
    __shared__ __align__(8) uint64_t bars[2];
    if (threadIdx.x == 0 && MODE == 1) {
        for (int b = 0; b < 2; ++b)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[b])), "r"(kThreads));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t x = hash32(blockIdx.x * kThreads + threadIdx.x + 1);
    double acc = 0.0;
    auto issue = [&](int buf) {
        if (MODE == 1)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[buf])), "r"(U * 16) : "memory");
#pragma unroll
        for (int u = 0; u < U; ++u) {
            x = hash32(x + u);
            uint32_t row = x & (kRows - 1);
            if (HOT && (x & 0x10000u)) row = kRows - 1;          // half of the gathers hit ONE row (saturated pixels)
            const uint32_t dst = smem_u32(&land[buf][u][threadIdx.x]);
            if (MODE == 0)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(table + row) : "memory");
            else if (MODE == 2)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(table + row) : "memory");
            else
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];" ::"r"(dst),
                             "l"(table + row), "r"(smem_u32(&bars[buf]))
                             : "memory");
        }
        if (MODE != 1) asm volatile("cp.async.commit_group;" ::: "memory");
    };
    issue(0);
    for (int it = 0; it < iters; ++it) {
        const int buf = it & 1;
        if (it + 1 < iters) issue(buf ^ 1);
        if (MODE != 1) {
            if (it + 1 < iters) asm volatile("cp.async.wait_group 1;" ::: "memory");
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
        } else {
            const uint32_t parity = (it >> 1) & 1;
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(smem_u32(&bars[buf])), "r"(parity) : "memory");
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += land[buf][u][threadIdx.x].x + land[buf][u][threadIdx.x].y;
    }
    if (acc == 12345.678) sink[0] = acc;
}

template <int MODE>   // 0 l2, 1 dsmem, 2 mixed, 3 local
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads, 1)
gather_kernel(const double2* __restrict__ table, int iters, double* __restrict__ sink) {
    extern __shared__ __align__(16) unsigned char smem[];
    double2* slice = reinterpret_cast<double2*>(smem);
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t crank = cluster.block_rank();
    if (MODE != 0) {
        for (int i = threadIdx.x; i < kRowsPerCta; i += kThreads) slice[i] = table[crank * kRowsPerCta + i];
    }
    cluster.sync();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(slice);
    uint32_t x = hash32(blockIdx.x * kThreads + threadIdx.x + 1);
    double acc = 0.0;
    constexpr int U = 8;
    for (int it = 0; it < iters; ++it) {
        double2 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            x = hash32(x + u);
            const uint32_t row = x & (kRows - 1);
            const bool via_smem = MODE == 1 || MODE == 3 || (MODE == 2 && (u & 1));
            if (via_smem) {
                const uint32_t owner = MODE == 3 ? crank : row / kRowsPerCta;
                const uint32_t local = base + (row % kRowsPerCta) * 16;
                v[u] = ld_cluster(mapa(local, owner));
            } else {
                v[u] = __ldg(table + row);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y;
    }
    if (acc == 12345.678) sink[0] = acc;
    cluster.sync();
}

int main() {
    double2* table;
    double* sink;
    cudaMalloc(&table, kRows * sizeof(double2));
    cudaMalloc(&sink, 8);
    cudaMemset(table, 0, kRows * sizeof(double2));
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clock_khz = 0;
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
    const int iters = 2000;
    const char* names[4] = {"l2", "dsmem", "mixed", "local"};
    for (int mode = 0; mode < 4; ++mode) {
        for (int grid : {128, 144}) {
            const size_t smem = kRowsPerCta * sizeof(double2);
            cudaEvent_t a, b;
            cudaEventCreate(&a); cudaEventCreate(&b);
            auto launch = [&]() {
                switch (mode) {
                    case 0: gather_kernel<0><<<grid, kThreads, smem>>>(table, iters, sink); break;
                    case 1: gather_kernel<1><<<grid, kThreads, smem>>>(table, iters, sink); break;
                    case 2: gather_kernel<2><<<grid, kThreads, smem>>>(table, iters, sink); break;
                    default: gather_kernel<3><<<grid, kThreads, smem>>>(table, iters, sink); break;
                }
            };
            switch (mode) {
                case 0: cudaFuncSetAttribute(gather_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); break;
                case 1: cudaFuncSetAttribute(gather_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); break;
                case 2: cudaFuncSetAttribute(gather_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); break;
                default: cudaFuncSetAttribute(gather_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); break;
            }
            launch();
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s grid %d: %s\n", names[mode], grid, cudaGetErrorString(e)); cudaGetLastError(); continue; }
            cudaEventRecord(a);
            launch();
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms = 0;
            cudaEventElapsedTime(&ms, a, b);
            const double gathers = (double)grid * kThreads * iters * 8;
            printf("%-6s grid %3d: %.3f ms, %.1f G gathers/s, %.2f gathers per CTA-clock (at %.0f MHz)\n", names[mode], grid, ms,
                   gathers / ms / 1e6, gathers / grid / (ms * 1e-3) / (clock_khz * 1e3), clock_khz / 1e3);
        }
    }
    const char* anames[5] = {"ldgsts.cg", "tma16", "ldgsts.ca", "ldgsts.cg hot", "ldgsts.ca hot"};
    for (int mode = 0; mode < 5; ++mode) {
        const int grid = 148;
        const int its = mode == 1 ? iters / 4 : (mode >= 3 ? iters / 8 : iters);
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        auto launch = [&]() {
            const int dsm = 2 * 8 * kThreads * 16;
#define GO(M, H) cudaFuncSetAttribute(async_gather_kernel<M, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, dsm); \
                 async_gather_kernel<M, H><<<grid, kThreads, dsm>>>(table, its, sink)
            switch (mode) {
                case 0: GO(0, false); break;
                case 1: GO(1, false); break;
                case 2: GO(2, false); break;
                case 3: GO(0, true); break;
                default: GO(2, true); break;
            }
        };
        launch();
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", anames[mode], cudaGetErrorString(e)); cudaGetLastError(); continue; }
        cudaEventRecord(a);
        launch();
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        const double gathers = (double)grid * kThreads * its * 8;
        printf("%-14s grid %3d: %.3f ms, %.1f G gathers/s, %.2f gathers per CTA-clock\n", anames[mode], grid, ms,
               gathers / ms / 1e6, gathers / grid / (ms * 1e-3) / (clock_khz * 1e3));
    }
    return 0;
}
