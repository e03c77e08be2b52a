// Micro-benchmark: do the register-return gather path (LDG) and the shared-memory-landing path (LDGSTS, cp.async)
// add up?  gather_bench.cu measured 0.51 (LDG) and 1.00 (LDGSTS) random 16-byte gathers per SM clock from a 1 MB table;
// here every thread sends NL of its 8 gathers per round through LDG and 8 - NL through cp.async.ca.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/gather_mix tools/microbench/gather_mix.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kRows = 65536;
constexpr int U = 8;

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int NL, int THREADS, int CTAS>
__global__ void __launch_bounds__(THREADS, CTAS)
mix_kernel(const double2* __restrict__ table, int iters, double* __restrict__ sink) {
    extern __shared__ __align__(16) unsigned char dyn[];
    constexpr int NA = U - NL;
    double2 (*land)[NA > 0 ? NA : 1][THREADS] = reinterpret_cast<double2 (*)[NA > 0 ? NA : 1][THREADS]>(dyn);
    uint32_t x = hash32(blockIdx.x * THREADS + threadIdx.x + 1);
    double acc = 0.0;
    double2 v[NL > 0 ? NL : 1];
    auto issue = [&](int buf) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            x = hash32(x + u);
            const uint32_t row = x & (kRows - 1);
            if (u < NA) {
                const uint32_t dst = smem_u32(&land[buf][u][threadIdx.x]);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(table + row) : "memory");
            } else {
                v[u - NA] = __ldg(table + row);
            }
        }
        if (NA > 0) asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int it = 0; it < iters; ++it) {
        const int buf = it & 1;
        issue(buf);
        if (NA > 0) asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
        for (int u = 0; u < NA; ++u) acc += land[buf][u][threadIdx.x].x + land[buf][u][threadIdx.x].y;
#pragma unroll
        for (int u = 0; u < NL; ++u) acc += v[u].x + v[u].y;
    }
    if (acc == 12345.678) sink[0] = acc;
}

// 8-byte and 32-byte register gathers (is the LDG rate per request, whatever the width?)
template <typename T, int THREADS, int CTAS>
__global__ void __launch_bounds__(THREADS, CTAS)
ldg_kernel(const T* __restrict__ table, int rows, int iters, double* __restrict__ sink) {
    uint32_t x = hash32(blockIdx.x * THREADS + threadIdx.x + 1);
    double acc = 0.0;
    for (int it = 0; it < iters; ++it) {
        T v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            x = hash32(x + u);
            v[u] = __ldg(table + (x & (rows - 1)));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += reinterpret_cast<const double*>(&v[u])[0];
    }
    if (acc == 12345.678) sink[0] = acc;
}

template <typename K>
static void time_it(const char* name, K launch, double gathers, int ctas_total, int clock_khz, int sms) {
    launch();
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); cudaGetLastError(); return; }
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    printf("%-44s %.3f ms  %.1f G gathers/s  %.2f gathers per SM clock\n", name, ms, gathers / ms / 1e6,
           gathers / sms / (ms * 1e-3) / (clock_khz * 1e3));
}

template <int NL, int THREADS, int CTAS>
static void run_mix(const double2* table, double* sink, int clock_khz, int sms) {
    const int iters = 1000;
    const int grid = sms * CTAS;
    const int smem = 2 * (U - NL > 0 ? U - NL : 1) * THREADS * 16;
    cudaFuncSetAttribute(mix_kernel<NL, THREADS, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    char name[96];
    snprintf(name, sizeof name, "mix: %d LDG + %d LDGSTS, %d thr x %d CTA/SM", NL, U - NL, THREADS, CTAS);
    time_it(name, [&]() { mix_kernel<NL, THREADS, CTAS><<<grid, THREADS, smem>>>(table, iters, sink); },
            (double)grid * THREADS * iters * U, grid, clock_khz, sms);
}

// all-LDG gathers with `smem` bytes of (unused) dynamic shared memory: the carve-out shrinks L1
static void run_l1_sweep(const double2* table, double* sink, int clock_khz, int sms) {
    const int iters = 1000;
    for (int kb : {0, 8, 16, 32, 64, 100, 132, 164, 196, 224}) {
        const int smem = kb * 1024;
        cudaFuncSetAttribute(mix_kernel<8, 512, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        char name[96];
        snprintf(name, sizeof name, "LDG.128, 512 thr x 1 CTA/SM, %3d KB shared memory", kb);
        time_it(name, [&]() { mix_kernel<8, 512, 1><<<sms, 512, smem>>>(table, iters, sink); },
                (double)sms * 512 * iters * U, sms, clock_khz, sms);
    }
}

int main() {
    double2* table;
    double* sink;
    cudaMalloc(&table, 4 * kRows * sizeof(double2));
    cudaMalloc(&sink, 8);
    cudaMemset(table, 0, 4 * kRows * sizeof(double2));
    int sms = 0, clock_khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);

    run_l1_sweep(table, sink, clock_khz, sms);
    run_mix<8, 512, 1>(table, sink, clock_khz, sms);
    run_mix<8, 512, 2>(table, sink, clock_khz, sms);
    run_mix<8, 256, 8>(table, sink, clock_khz, sms);
    run_mix<0, 512, 1>(table, sink, clock_khz, sms);
    run_mix<0, 512, 2>(table, sink, clock_khz, sms);
    run_mix<0, 512, 4>(table, sink, clock_khz, sms);
    run_mix<6, 512, 2>(table, sink, clock_khz, sms);
    run_mix<5, 512, 2>(table, sink, clock_khz, sms);
    run_mix<4, 512, 2>(table, sink, clock_khz, sms);
    run_mix<3, 512, 2>(table, sink, clock_khz, sms);
    run_mix<2, 512, 2>(table, sink, clock_khz, sms);
    run_mix<1, 512, 2>(table, sink, clock_khz, sms);
    run_mix<3, 512, 4>(table, sink, clock_khz, sms);
    run_mix<2, 512, 4>(table, sink, clock_khz, sms);

    const int iters = 1000;
    const int grid = sms * 2;
    time_it("LDG.64  (8 B rows, 512 KB table)", [&]() {
        ldg_kernel<double, 512, 2><<<grid, 512>>>(reinterpret_cast<const double*>(table), kRows, iters, sink); },
        (double)grid * 512 * iters * U, grid, clock_khz, sms);
    time_it("LDG.128 (16 B rows, 1 MB table)", [&]() {
        ldg_kernel<double2, 512, 2><<<grid, 512>>>(table, kRows, iters, sink); },
        (double)grid * 512 * iters * U, grid, clock_khz, sms);
    time_it("LDG.64  (8 B rows, 128 KB table = L1 resident)", [&]() {
        ldg_kernel<double, 512, 2><<<grid, 512>>>(reinterpret_cast<const double*>(table), 16384, iters, sink); },
        (double)grid * 512 * iters * U, grid, clock_khz, sms);
    time_it("LDG.128 (16 B rows, 128 KB table = L1 resident)", [&]() {
        ldg_kernel<double2, 512, 2><<<grid, 512>>>(table, 8192, iters, sink); },
        (double)grid * 512 * iters * U, grid, clock_khz, sms);
    time_it("LDG.128 (16 B rows, 256 KB table)", [&]() {
        ldg_kernel<double2, 512, 2><<<grid, 512>>>(table, 16384, iters, sink); },
        (double)grid * 512 * iters * U, grid, clock_khz, sms);
    time_it("LDG.128 (16 B rows, 512 KB table)", [&]() {
        ldg_kernel<double2, 512, 2><<<grid, 512>>>(table, 32768, iters, sink); },
        (double)grid * 512 * iters * U, grid, clock_khz, sms);
    return 0;
}
