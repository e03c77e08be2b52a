// K4: ICRF calibration objective for a whole differential-evolution population in one launch.
// Replaces ICRF_calibration_exposure.py:20-44, 66-201 and general_functions.py:149-176.
//
// Mapping: one LANE per candidate curve (32 candidates per warp), one WARP-ITERATION per pixel.
// The pixel's N digital numbers are warp-uniform, so the candidate tables are laid out
// [dn][candidate] and every gather `table[dn_k][lane]` is a contiguous, conflict-free 512-byte
// row of shared memory (a {value, reciprocal} double2 per candidate).  Each lane keeps the
// per-exposure-pair numerator / denominator of ITS candidate in registers, so no cross-lane
// reduction is needed at all; warps are combined through shared memory and CTAs through a fixed-
// order second kernel (deterministic, and the same pair sums are what NCCL all-reduces between
// GPUs).  Range masks are folded into the table as NaN entries (the mask depends only on the DN),
// and |I_i - r I_j| / (r I_j) is evaluated as |I_i * (1/I_j) * (1/r) - 1| -- no division in the
// inner loop.  Bound: FP64 pipe + shared-memory gathers; the 2-18 MB of pixel data stay in L2.
#include "common.cuh"

namespace cl {
namespace {

constexpr int kGroup = 32;            // candidates per warp
constexpr int kMaxN = CL_MAX_PAIR_EXPOSURES;

__host__ __device__ inline int n_pairs(int n) { return n * (n - 1) / 2; }

struct PairRatios {
    double v[kMaxN * (kMaxN - 1) / 2];
};

// ---- candidate curves, gates and tables -----------------------------------------------------------
// One CTA per candidate, one thread per curve point.
__global__ void curves_kernel(const cl_icrf_problem prob, const double* __restrict__ mean_icrf,
                              const double* __restrict__ pca, const double* __restrict__ params,
                              double* __restrict__ curves, int32_t* __restrict__ valid,
                              double2* __restrict__ tables) {
    __shared__ double sc[256];
    const int s = blockIdx.x, d = threadIdx.x, D = prob.datapoints;
    const double* p = params + (int64_t)s * prob.n_params;
    const int n_pc = prob.use_mean_icrf ? prob.n_params : prob.n_params - 1;
    const double* coef = prob.use_mean_icrf ? p : p + 1;
    double v = 0.0;
    if (d < D) {
        double acc = 0.0;                       // matmul(PCA_array, params), :38-40
        for (int k = 0; k < n_pc; ++k) acc = __dadd_rn(acc, __dmul_rn(pca[d * n_pc + k], coef[k]));
        double base;
        if (prob.use_mean_icrf) {
            base = mean_icrf[d];
        } else {                                // linspace(0, 1, BITS) ** p[0], :37
            const double x = (d == D - 1) ? 1.0 : __dmul_rn((double)d, __ddiv_rn(1.0, (double)(D - 1)));
            base = pow(x, p[0]);
        }
        v = __dadd_rn(base, acc);
        sc[d] = v;
    }
    __syncthreads();
    const double shift = __dsub_rn(1.0, sc[D - 1]);     // ICRF += 1 - ICRF[-1], :167
    __syncthreads();
    if (d < D) {
        v = __dadd_rn(v, shift);
        if (d == 0) v = 0.0;                             // ICRF[0] = 0, :168
        sc[d] = v;
    }
    __syncthreads();
    int bad = 0;
    if (d < D) {
        if (v > 1.0 || v < 0.0) bad = 1;                 // :174  (NaN compares false, like NumPy max)
        if (d > 0 && !(v > sc[d - 1])) bad = 1;          // :178  strictly increasing
        if (v != v) bad = 1;                             // NaN: np.max -> NaN, comparisons False, then
                                                         // `all(nan > ..)` is False -> inf as well
    }
    bad = __syncthreads_or(bad);
    if (d < D) {
        curves[(int64_t)s * D + d] = v;
        const double lo = sc[prob.lower], hi = sc[prob.upper];      // :182-183
        const double m = (v < lo || v > hi) ? __longlong_as_double(0x7ff8000000000000LL) : v;  // :97-98
        // table layout [group][dn][lane]
        tables[((int64_t)(s / kGroup) * D + d) * kGroup + (s % kGroup)] = make_double2(m, 1.0 / m);
    }
    if (d == 0) valid[s] = bad ? 0 : 1;
}

// ---- partial energies -----------------------------------------------------------------------------
// One pixel per warp-iteration, two pixels in flight (the DN / sigma loads of the next pixel are
// issued before the arithmetic of the current one).  No branches in the pair loop: invalid pairs
// contribute exact zeros.
template <int N, bool USE_STD>
struct PixelData {
    int bin[N];
    double sg[USE_STD ? N : 1];
};

template <int N, bool USE_STD>
__device__ __forceinline__ void load_pixel(PixelData<N, USE_STD>& d, const uint8_t* __restrict__ dn,
                                           const double* __restrict__ sd, int64_t px, int D) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
        d.bin[k] = min((int)__ldg(dn + px * N + k), D - 1);        // callers guarantee dn < D
        if (USE_STD) d.sg[USE_STD ? k : 0] = __ldg(sd + px * N + k);
    }
}

template <int N, bool USE_STD, int P>
__device__ __forceinline__ void accumulate_pixel(const PixelData<N, USE_STD>& d, const double2* __restrict__ tab,
                                                 int lane, const PairRatios& inv_ratio, double (&num)[P],
                                                 double (&den)[P]) {
    double I[N], R[N];
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double2 e = tab[d.bin[k] * kGroup + lane];
        I[k] = e.x;
        R[k] = e.y;
    }
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    int q = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
#pragma unroll
        for (int j = i + 1; j < N; ++j, ++q) {
            const double m = R[j] * inv_ratio.v[q];   // 1 / (I_j * r)
            const double u = I[i] * m;                // I_i / scaled
            const double a = fabs(u - 1.0);           // |I_i - scaled| / scaled, :115-121
            if (!USE_STD) {
                const bool ok = a == a;               // nanmean, :139 (inf is kept, as NumPy does)
                num[q] += ok ? a : 0.0;
                den[q] += ok ? 1.0 : 0.0;
            } else {
                const double t1 = d.sg[USE_STD ? i : 0] * m;                       // sigma_i / scaled
                const double t2 = u * (d.sg[USE_STD ? j : 0] * R[j]);              // I_i sigma_j / (r I_j^2)
                const double var = fma(t1, t1, t2 * t2);                           // :128
                // finite |d|, sigma != 0, weight 1/sigma not NaN (:134-135, gf.nanaverage)
                const bool ok = (a < inf) && (var > 0.0);
                const double w = ok ? rsqrt(var) : 0.0;
                num[q] = fma(ok ? a : 0.0, w, num[q]);
                den[q] += w;
            }
        }
    }
}

template <int N, bool USE_STD, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
energy_partial_kernel(const double2* __restrict__ tables, int D, const uint8_t* __restrict__ dn,
                      const double* __restrict__ sd, int64_t n_pixels, int64_t px_per_cta,
                      const __grid_constant__ PairRatios inv_ratio /* [pairs] = t_j / t_i */,
                      double* __restrict__ cta_partial /* [cta][S][pairs][2] */, int S) {
    constexpr int P = N * (N - 1) / 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* tab = reinterpret_cast<double2*>(smem_raw);           // [D][32]
    const int group = blockIdx.y;
    const double2* src = tables + (int64_t)group * D * kGroup;
    for (int i = threadIdx.x; i < D * kGroup; i += WARPS * 32) tab[i] = src[i];
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double num[P], den[P];
#pragma unroll
    for (int q = 0; q < P; ++q) { num[q] = 0.0; den[q] = 0.0; }
    const int64_t first = (int64_t)blockIdx.x * px_per_cta;
    const int64_t last = min(n_pixels, first + px_per_cta);
    int64_t px = first + warp;
    PixelData<N, USE_STD> cur, nxt;
    if (px < last) load_pixel<N, USE_STD>(cur, dn, sd, px, D);
    for (; px < last; px += WARPS) {
        const int64_t pn = px + WARPS;
        if (pn < last) load_pixel<N, USE_STD>(nxt, dn, sd, pn, D);     // prefetch
        accumulate_pixel<N, USE_STD, P>(cur, tab, lane, inv_ratio, num, den);
        cur = nxt;
    }

    // combine the CTA's warps with a fixed-order tree through shared memory (table space reused)
    double* red = reinterpret_cast<double*>(smem_raw);             // [half][2P][32]
    for (int half = WARPS / 2; half >= 1; half >>= 1) {
        __syncthreads();
        if (warp >= half && warp < 2 * half) {
#pragma unroll
            for (int q = 0; q < P; ++q) {
                red[(((warp - half) * 2 * P) + 2 * q) * 32 + lane] = num[q];
                red[(((warp - half) * 2 * P) + 2 * q + 1) * 32 + lane] = den[q];
            }
        }
        __syncthreads();
        if (warp < half) {
#pragma unroll
            for (int q = 0; q < P; ++q) {
                num[q] += red[((warp * 2 * P) + 2 * q) * 32 + lane];
                den[q] += red[((warp * 2 * P) + 2 * q + 1) * 32 + lane];
            }
        }
    }
    if (warp == 0) {
        const int s = group * kGroup + lane;
        double* out = cta_partial + (((int64_t)blockIdx.x * S + s) * P) * 2;
#pragma unroll
        for (int q = 0; q < P; ++q) {
            out[2 * q] = num[q];
            out[2 * q + 1] = den[q];
        }
    }
}

// One warp per output: lanes stride over the CTAs, then an xor tree -- a fixed order, so the result is
// deterministic (a single thread walking all CTAs' partials took 13 us of a 150 us evaluation).
__global__ void reduce_ctas_kernel(const double* __restrict__ cta_partial, int n_ctas, int64_t per_cta,
                                   double* __restrict__ out) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= per_cta) return;
    double s = 0.0;
#pragma unroll 4
    for (int c = lane; c < n_ctas; c += 32) s += cta_partial[(int64_t)c * per_cta + i];
    s = warp_sum(s);
    if (lane == 0) out[i] = s;
}

__global__ void finalize_kernel(const double* __restrict__ pair_acc, const int32_t* __restrict__ valid,
                                int S, int P, double* __restrict__ energy) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    double e = inf;
    if (valid[s]) {
        double sum = 0.0;
        int cnt = 0;
        for (int q = 0; q < P; ++q) {
            const double num = pair_acc[((int64_t)s * P + q) * 2], den = pair_acc[((int64_t)s * P + q) * 2 + 1];
            if (den == 0.0) continue;                    // empty pair -> NaN -> skipped by nanmean
            const double r = num / den;
            if (r != r) continue;
            sum += r;
            ++cnt;
        }
        if (cnt > 0) e = sum / (double)cnt;               // np.nanmean(linearity_data), :196
        if (e != e) e = inf;                              // :197-198
    }
    energy[s] = e;
}

struct Plan {
    int groups, chunks;
    int64_t px_per_cta;
};

inline Plan make_plan(const cl_icrf_problem& p, int64_t n_pixels) {
    Plan pl;
    pl.groups = p.n_candidates / kGroup;
    int chunks = sm_count() / (pl.groups > 0 ? pl.groups : 1);
    if (chunks < 1) chunks = 1;
    const int64_t min_px = 16 * 8;         // do not split tiny problems into idle CTAs
    if ((int64_t)chunks * min_px > n_pixels) chunks = (int)((n_pixels + min_px - 1) / min_px);
    if (chunks < 1) chunks = 1;
    pl.chunks = chunks;
    pl.px_per_cta = (n_pixels + chunks - 1) / chunks;
    return pl;
}

inline bool problem_ok(const cl_icrf_problem* p) {
    return p && p->n_candidates >= kGroup && p->n_candidates % kGroup == 0 && p->n_params >= 1 &&
           p->datapoints >= 2 && p->datapoints <= 256 && p->lower >= 0 && p->lower < p->datapoints &&
           p->upper >= 0 && p->upper < p->datapoints && p->n_exposures >= 2 &&
           p->n_exposures <= kMaxN && (p->use_mean_icrf || p->n_params >= 2);
}

template <int N>
int launch_partial(const cl_icrf_problem& p, const Plan& pl, const double2* tables, const uint8_t* dn,
                   const double* sd, int64_t n_pixels, const PairRatios& inv_ratio, double* cta_partial,
                   cudaStream_t stream) {
    constexpr int WARPS = N <= 6 ? 16 : 8;         // registers: 2P accumulators per lane
    const size_t tab_bytes = (size_t)p.datapoints * kGroup * sizeof(double2);
    const size_t red_bytes = (size_t)(WARPS / 2) * 2 * n_pairs(N) * 32 * sizeof(double);
    const size_t smem = tab_bytes > red_bytes ? tab_bytes : red_bytes;
    dim3 grid(pl.chunks, pl.groups);
    auto go = [&](auto kernel) -> int {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_status(e);
        kernel<<<grid, WARPS * 32, smem, stream>>>(tables, p.datapoints, dn, sd, n_pixels, pl.px_per_cta,
                                                   inv_ratio, cta_partial, p.n_candidates);
        return launched();
    };
    return sd ? go(energy_partial_kernel<N, true, WARPS>) : go(energy_partial_kernel<N, false, WARPS>);
}

}  // namespace
}  // namespace cl

extern "C" {

size_t cl_icrf_tables_bytes(const cl_icrf_problem* p) {
    if (!cl::problem_ok(p)) return 0;
    return (size_t)p->n_candidates * p->datapoints * sizeof(double2);
}

int cl_icrf_curves(const cl_icrf_problem* p, const double* mean_icrf, const double* pca,
                   const double* params, double* curves, int32_t* valid, void* tables, void* stream) {
    using namespace cl;
    if (!problem_ok(p)) return CL_ERR_INVALID_ARGUMENT;
    CL_REQUIRE(pca && params && curves && valid && tables);
    CL_REQUIRE(mean_icrf || !p->use_mean_icrf);
    if (!aligned(tables, 16)) return CL_ERR_ALIGNMENT;
    curves_kernel<<<p->n_candidates, 256, 0, (cudaStream_t)stream>>>(
        *p, mean_icrf, pca, params, curves, valid, reinterpret_cast<double2*>(tables));
    return launched();
}

size_t cl_icrf_energy_workspace_bytes(const cl_icrf_problem* p, int64_t n_pixels) {
    if (!cl::problem_ok(p) || n_pixels < 0) return 0;
    const cl::Plan pl = cl::make_plan(*p, n_pixels > 0 ? n_pixels : 1);
    const size_t pairs = cl::n_pairs(p->n_exposures);
    return (size_t)pl.chunks * p->n_candidates * pairs * 2 * sizeof(double);
}

int cl_icrf_energy_partial(const cl_icrf_problem* p, const void* tables, const uint8_t* dn,
                           const double* std, const double* exposure_s, int64_t n_pixels,
                           double* pair_acc, void* workspace, size_t workspace_bytes, void* stream) {
    using namespace cl;
    if (!problem_ok(p)) return CL_ERR_INVALID_ARGUMENT;
    CL_REQUIRE(tables && exposure_s && pair_acc && n_pixels >= 0);
    CL_REQUIRE(n_pixels == 0 || dn);
    CL_REQUIRE((p->use_std != 0) == (std != nullptr) || n_pixels == 0);
    if (!workspace || workspace_bytes < cl_icrf_energy_workspace_bytes(p, n_pixels)) return CL_ERR_WORKSPACE;
    if (!aligned(workspace, 16) || !aligned(tables, 16)) return CL_ERR_ALIGNMENT;
    cudaStream_t s = (cudaStream_t)stream;
    const int N = p->n_exposures, P = n_pairs(N);
    const int64_t per_cta = (int64_t)p->n_candidates * P * 2;
    if (n_pixels == 0) return cuda_status(cudaMemsetAsync(pair_acc, 0, per_cta * sizeof(double), s));

    // 1 / exposure ratio per pair (i < j), ratio = t_i / t_j (ICRF_calibration_exposure.py:101)
    PairRatios d_rr;
    int q = 0;
    for (int i = 0; i < N; ++i)
        for (int j = i + 1; j < N; ++j) d_rr.v[q++] = 1.0 / (exposure_s[i] / exposure_s[j]);
    for (; q < kMaxN * (kMaxN - 1) / 2; ++q) d_rr.v[q] = 0.0;
    double* cta_partial = reinterpret_cast<double*>(workspace);
    const Plan pl = make_plan(*p, n_pixels);
    const double2* tab = reinterpret_cast<const double2*>(tables);
    int st;
    switch (N) {
        case 2: st = launch_partial<2>(*p, pl, tab, dn, std, n_pixels, d_rr, cta_partial, s); break;
        case 3: st = launch_partial<3>(*p, pl, tab, dn, std, n_pixels, d_rr, cta_partial, s); break;
        case 4: st = launch_partial<4>(*p, pl, tab, dn, std, n_pixels, d_rr, cta_partial, s); break;
        case 5: st = launch_partial<5>(*p, pl, tab, dn, std, n_pixels, d_rr, cta_partial, s); break;
        case 6: st = launch_partial<6>(*p, pl, tab, dn, std, n_pixels, d_rr, cta_partial, s); break;
        case 7: st = launch_partial<7>(*p, pl, tab, dn, std, n_pixels, d_rr, cta_partial, s); break;
        case 8: st = launch_partial<8>(*p, pl, tab, dn, std, n_pixels, d_rr, cta_partial, s); break;
        default: return CL_ERR_UNSUPPORTED;
    }
    if (st != CL_OK) return st;
    reduce_ctas_kernel<<<(unsigned)((per_cta * 32 + 255) / 256), 256, 0, s>>>(cta_partial, pl.chunks, per_cta,
                                                                             pair_acc);
    return launched();
}

int cl_icrf_energy_finalize(const cl_icrf_problem* p, const double* pair_acc, const int32_t* valid,
                            double* energy, void* stream) {
    using namespace cl;
    if (!problem_ok(p)) return CL_ERR_INVALID_ARGUMENT;
    CL_REQUIRE(pair_acc && valid && energy);
    finalize_kernel<<<(p->n_candidates + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        pair_acc, valid, p->n_candidates, n_pairs(p->n_exposures), energy);
    return launched();
}

}  // extern "C"
