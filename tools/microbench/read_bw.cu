// Micro-benchmark: what read bandwidth does a B200 deliver to a read-dominated streaming kernel?  (development tool)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/read_bw tools/microbench/read_bw.cu && /tmp/read_bw
// MEASURED_PEAKS.json's HBM figure is a copy (1 read : 1 write).  The merge kernels read ~10 bytes per byte written
// and level off at ~84 % of it; this program measures the ceiling for that traffic mix with nothing else in the way:
//   ldg   : grid-stride 16-byte loads, 8 in flight per thread, xor-reduced (pure read)
//   ring  : one persistent CTA per SM, a producer thread streaming 13.8 KB stages with cp.async.bulk into a ring of S
//           stages, consumer warps that only touch each stage and release it (the structure of merge_stream_kernel
//           with the arithmetic removed); optionally 16 bytes written per 176 bytes read, like cfg2
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void __launch_bounds__(256) ldg_kernel(const uint4* __restrict__ src, int64_t n_vec, uint32_t* out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    uint32_t acc = 0;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 7 * stride < n_vec; i += 8 * stride) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcs(src + i + u * stride);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    for (; i < n_vec; i += stride) { const uint4 v = __ldcs(src + i); acc ^= v.x ^ v.y ^ v.z ^ v.w; }
    if (acc == 0x12345678u) out[0] = acc;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int kConsumerWarps = 16;
constexpr int kMaxStages = 16;

// chunks of `stage_bytes`, chunk j of CTA b = global chunk b + j * grid;  every kWriteEvery-th chunk the consumers
// write `write_bytes` of output (0 = pure read)
__global__ void __launch_bounds__((kConsumerWarps + 1) * 32, 1)
ring_kernel(const unsigned char* __restrict__ src, int64_t n_chunks, int stage_bytes, int stages, int write_every,
            double* __restrict__ dst, uint32_t* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full[kMaxStages], empty[kMaxStages];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kConsumerWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == kConsumerWarps) {
        if (lane == 0) {
            int s = 0; uint32_t phase = 0;
            for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
                mbar_wait(&empty[s], phase ^ 1);
                mbar_expect(&full[s], stage_bytes);
                bulk_g2s(smem + (size_t)s * stage_bytes, src + c * stage_bytes, stage_bytes, &full[s]);
                if (++s == stages) { s = 0; phase ^= 1; }
            }
        }
    } else {
        int s = 0; uint32_t phase = 0, acc = 0;
        int64_t j = 0;
        for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x, ++j) {
            mbar_wait(&full[s], phase);
            const uint32_t v = reinterpret_cast<const uint32_t*>(smem + (size_t)s * stage_bytes)[threadIdx.x];
            acc ^= v;
            __syncwarp();
            if (lane == 0 && (acc | 1u)) mbar_arrive(&empty[s]);
            if (++s == stages) { s = 0; phase ^= 1; }
            if (write_every && (j % write_every) == write_every - 1) {
                // 512 threads x 6 doubles = 24 KB per tile, like the merge's val + std outputs
                double* o = dst + ((c / gridDim.x / write_every) * gridDim.x + blockIdx.x) * 3072 + threadIdx.x * 6;
#pragma unroll
                for (int q = 0; q < 6; ++q) __stcs(o + q, (double)acc);
            }
        }
        if (acc == 0x12345678u) out[0] = acc;
    }
}

int main() {
    const int64_t bytes = (int64_t)4 << 30;
    unsigned char* src; uint32_t* out; double* dst;
    CK(cudaMalloc(&src, bytes)); CK(cudaMalloc(&out, 4)); CK(cudaMalloc(&dst, (size_t)512 << 20));
    CK(cudaMemset(src, 1, bytes));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    auto timeit = [&](auto launch, const char* name, double nbytes) {
        for (int i = 0; i < 2; ++i) launch();
        CK(cudaEventRecord(a));
        const int reps = 5;
        for (int i = 0; i < reps; ++i) launch();
        CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        CK(cudaGetLastError());
        printf("%-52s %8.3f ms  %7.1f GB/s\n", name, ms / reps, nbytes / (ms / reps) / 1e6);
    };
    for (int per_sm : {4, 8}) {
        char name[96]; snprintf(name, sizeof name, "ldg  16 B x 8 in flight, %d CTAs of 256 per SM", per_sm);
        timeit([&] { ldg_kernel<<<sms * per_sm, 256>>>(reinterpret_cast<const uint4*>(src), bytes / 16, out); }, name, (double)bytes);
    }
    CK(cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    for (int stage_bytes : {13824, 27648}) {
        for (int stages : {4, 7, 10, 15}) {
            if ((size_t)stages * stage_bytes > 220 * 1024) continue;
            for (int write_every : {0, 17}) {
                const int64_t n_chunks = bytes / stage_bytes;
                const double wbytes = write_every ? (double)(n_chunks / write_every) * 3072 * 8 : 0.0;
                char name[96];
                snprintf(name, sizeof name, "ring %5d B x %2d stages%s", stage_bytes, stages,
                         write_every ? ", 24 KB written per 17 stages" : ", pure read");
                timeit([&] {
                    ring_kernel<<<sms, (kConsumerWarps + 1) * 32, (size_t)stages * stage_bytes>>>(
                        src, n_chunks, stage_bytes, stages, write_every, dst, out);
                }, name, (double)n_chunks * stage_bytes + wbytes);
            }
        }
    }
    return 0;
}
