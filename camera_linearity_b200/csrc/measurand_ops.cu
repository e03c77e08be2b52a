// Measurand operators with first-order uncertainty propagation, fused.
// Replaces the NumPy expression chains of AbstractMeasurand.__add__ / __sub__ / __truediv__ / __mul__ / __pow__
// (modules/measurand.py:106-241), log_e / log_10 (:243-279), the static compute_difference (:620-655) and
// interpolate (:658-681): the reference makes 4-10 full-array passes (one temporary per ufunc) per operator; here
// every operator is ONE streaming pass: 16-32 B in, 8-16 B out per element, HBM bound.
//
// Arithmetic follows the reference operation for operation with round-to-nearest intrinsics (no FMA contraction),
// so + - * / and their uncertainties are bit-identical to NumPy (x**2 is np.square = x*x; division and sqrt are
// IEEE); pow / log differ from libm by the usual 1-2 ulp.
// Broadcasting: the second operand may be the same size, or a "suffix" operand whose shape equals the trailing
// dimensions of the first (a per-channel vector, a scalar): element i of x pairs with element i % period of y.
#include "common.cuh"

#include <cmath>

namespace cl {
namespace {

enum Op { kAdd = 0, kSub = 1, kMul = 2, kDiv = 3, kPow = 4 };

__device__ __forceinline__ double sq(double a) { return __dmul_rn(a, a); }

template <int OP, bool USE_STD>
__device__ __forceinline__ void binary_one(double x1, double s1, double x2, double s2, double& v, double& s) {
    if (OP == kAdd) {
        v = __dadd_rn(x1, x2);
        if (USE_STD) s = __dsqrt_rn(__dadd_rn(sq(s1), sq(s2)));                      // :124
    } else if (OP == kSub) {
        v = __dsub_rn(x1, x2);
        if (USE_STD) s = __dsqrt_rn(__dadd_rn(sq(s1), sq(s2)));                      // :147
    } else if (OP == kMul) {
        v = __dmul_rn(x1, x2);
        if (USE_STD) s = __dsqrt_rn(__dadd_rn(sq(__dmul_rn(x1, s2)), sq(__dmul_rn(x2, s1))));   // :209
    } else if (OP == kDiv) {
        v = __ddiv_rn(x1, x2);
        if (USE_STD) {
            const double u1 = __ddiv_rn(s1, x2);                                     // :185
            const double u2 = __ddiv_rn(__dmul_rn(x1, s2), sq(x2));                  // :186
            s = __dsqrt_rn(__dadd_rn(sq(u1), sq(u2)));
        }
    } else {
        v = pow(x1, x2);
        if (USE_STD) {
            const double u1 = __dmul_rn(x2, pow(x1, __dsub_rn(x2, 1.0)));            // :236
            const double u2 = __dmul_rn(log(x1), v);                                 // :237
            s = __dsqrt_rn(__dadd_rn(sq(__dmul_rn(u1, s1)), sq(__dmul_rn(u2, s2))));
        }
    }
}

// s1 / s2 may be null (the reference substitutes zeros, :118-122)
template <int OP, bool USE_STD>
__global__ void __launch_bounds__(256)
binary_kernel(const double* __restrict__ x1, const double* __restrict__ s1, const double* __restrict__ x2,
              const double* __restrict__ s2, int64_t n, int64_t period, double* __restrict__ out_v,
              double* __restrict__ out_s) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 2;
    const bool same = period == n;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < n; i += stride) {
        if (same && i + 1 < n) {                       // two elements per thread: 16-byte accesses
            const double2 a = __ldcs(reinterpret_cast<const double2*>(x1 + i));
            const double2 b = __ldcs(reinterpret_cast<const double2*>(x2 + i));
            double2 sa = make_double2(0.0, 0.0), sb = make_double2(0.0, 0.0);
            if (USE_STD && s1) sa = __ldcs(reinterpret_cast<const double2*>(s1 + i));
            if (USE_STD && s2) sb = __ldcs(reinterpret_cast<const double2*>(s2 + i));
            double2 v, s = make_double2(0.0, 0.0);
            binary_one<OP, USE_STD>(a.x, sa.x, b.x, sb.x, v.x, s.x);
            binary_one<OP, USE_STD>(a.y, sa.y, b.y, sb.y, v.y, s.y);
            __stcs(reinterpret_cast<double2*>(out_v + i), v);
            if (USE_STD) __stcs(reinterpret_cast<double2*>(out_s + i), s);
        } else {
            for (int64_t j = i; j < n && j < i + 2; ++j) {
                const int64_t k = same ? j : j % period;
                double v, s = 0.0;
                binary_one<OP, USE_STD>(x1[j], (USE_STD && s1) ? s1[j] : 0.0, x2[k], (USE_STD && s2) ? s2[k] : 0.0, v, s);
                out_v[j] = v;
                if (USE_STD) out_s[j] = s;
            }
        }
    }
}

// log_e: val = log(x), std = std / log(x) (the reference's literal formula, :258);  log_10: val = log10(x),
// std = std / (x * (log 5 + log 2)) (:277)
template <bool BASE10, bool USE_STD>
__global__ void __launch_bounds__(256)
log_kernel(const double* __restrict__ x, const double* __restrict__ s, int64_t n, double ln10_sum,
           double* __restrict__ out_v, double* __restrict__ out_s) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double xv = x[i];
        if (BASE10) {
            out_v[i] = log10(xv);
            if (USE_STD) out_s[i] = __ddiv_rn(s[i], __dmul_rn(xv, ln10_sum));
        } else {
            const double l = log(xv);
            out_v[i] = l;
            if (USE_STD) out_s[i] = __ddiv_rn(s[i], l);
        }
    }
}

// compute_difference (:634-653): abs = x - m*y, rel = abs / (m*y) and their uncertainties, four outputs, one pass
template <bool USE_STD>
__global__ void __launch_bounds__(256)
difference_kernel(const double* __restrict__ x, const double* __restrict__ xs, const double* __restrict__ y,
                  const double* __restrict__ ys, double m, int64_t n, double* __restrict__ abs_v,
                  double* __restrict__ abs_s, double* __restrict__ rel_v, double* __restrict__ rel_s) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double xv = x[i], yv = y[i];
        const double scale = __dmul_rn(m, yv);                       // :634
        const double a = __dsub_rn(xv, scale);
        abs_v[i] = a;
        rel_v[i] = __ddiv_rn(a, scale);
        if (USE_STD) {
            const double sx = xs ? xs[i] : 0.0, sy = ys ? ys[i] : 0.0;
            abs_s[i] = __dsqrt_rn(__dadd_rn(sq(sx), sq(__dmul_rn(m, sy))));                              // :652
            const double t1 = __ddiv_rn(sx, __dmul_rn(m, yv));
            const double t2 = __ddiv_rn(__dmul_rn(sy, xv), __dmul_rn(m, sq(yv)));                       // :653
            rel_s[i] = __dsqrt_rn(__dadd_rn(sq(t1), sq(t2)));
        }
    }
}

inline unsigned grid_for(int64_t items, int threads) {
    int64_t b = (items + threads - 1) / threads;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (b > cap) b = cap;
    return (unsigned)(b < 1 ? 1 : b);
}

template <int OP>
int launch_binary(const double* x1, const double* s1, const double* x2, const double* s2, int64_t n, int64_t period,
                  bool use_std, double* out_v, double* out_s, cudaStream_t s) {
    const unsigned grid = grid_for((n + 1) / 2, 256);
    if (use_std) binary_kernel<OP, true><<<grid, 256, 0, s>>>(x1, s1, x2, s2, n, period, out_v, out_s);
    else binary_kernel<OP, false><<<grid, 256, 0, s>>>(x1, s1, x2, s2, n, period, out_v, out_s);
    return launched();
}

}  // namespace
}  // namespace cl

extern "C" {

int cl_measurand_binary(int op, const double* x_val, const double* x_std, const double* y_val, const double* y_std,
                        int64_t n, int64_t y_period, double* out_val, double* out_std, void* stream) {
    using namespace cl;
    CL_REQUIRE(op >= 0 && op <= 4 && n >= 0);
    if (n == 0) return CL_OK;
    CL_REQUIRE(x_val && y_val && out_val && y_period >= 1 && y_period <= n && n % y_period == 0);
    const bool use_std = out_std != nullptr;
    if (y_period == n && (!aligned(x_val, 16) || !aligned(y_val, 16) || !aligned(out_val, 16) ||
                          (x_std && !aligned(x_std, 16)) || (y_std && !aligned(y_std, 16)) ||
                          (out_std && !aligned(out_std, 16))))
        return CL_ERR_ALIGNMENT;
    cudaStream_t s = (cudaStream_t)stream;
    switch (op) {
        case kAdd: return launch_binary<kAdd>(x_val, x_std, y_val, y_std, n, y_period, use_std, out_val, out_std, s);
        case kSub: return launch_binary<kSub>(x_val, x_std, y_val, y_std, n, y_period, use_std, out_val, out_std, s);
        case kMul: return launch_binary<kMul>(x_val, x_std, y_val, y_std, n, y_period, use_std, out_val, out_std, s);
        case kDiv: return launch_binary<kDiv>(x_val, x_std, y_val, y_std, n, y_period, use_std, out_val, out_std, s);
        default: return launch_binary<kPow>(x_val, x_std, y_val, y_std, n, y_period, use_std, out_val, out_std, s);
    }
}

int cl_measurand_log(int base10, const double* val, const double* std, int64_t n, double* out_val, double* out_std,
                     void* stream) {
    using namespace cl;
    CL_REQUIRE(n >= 0);
    if (n == 0) return CL_OK;
    CL_REQUIRE(val && out_val && ((std != nullptr) == (out_std != nullptr)));
    cudaStream_t s = (cudaStream_t)stream;
    volatile double l5 = log(5.0), l2 = log(2.0);          // np.log(5) + np.log(2), evaluated like the reference (:277)
    const double ln10_sum = l5 + l2;
    const unsigned grid = grid_for(n, 256);
    if (base10) {
        if (std) log_kernel<true, true><<<grid, 256, 0, s>>>(val, std, n, ln10_sum, out_val, out_std);
        else log_kernel<true, false><<<grid, 256, 0, s>>>(val, std, n, ln10_sum, out_val, out_std);
    } else {
        if (std) log_kernel<false, true><<<grid, 256, 0, s>>>(val, std, n, ln10_sum, out_val, out_std);
        else log_kernel<false, false><<<grid, 256, 0, s>>>(val, std, n, ln10_sum, out_val, out_std);
    }
    return launched();
}

int cl_measurand_difference(const double* x_val, const double* x_std, const double* y_val, const double* y_std,
                            double multiplier, int64_t n, double* abs_val, double* abs_std, double* rel_val,
                            double* rel_std, void* stream) {
    using namespace cl;
    CL_REQUIRE(n >= 0);
    if (n == 0) return CL_OK;
    CL_REQUIRE(x_val && y_val && abs_val && rel_val);
    const bool use_std = x_std != nullptr || y_std != nullptr;
    CL_REQUIRE(!use_std || (abs_std && rel_std));
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned grid = grid_for(n, 256);
    if (use_std) difference_kernel<true><<<grid, 256, 0, s>>>(x_val, x_std, y_val, y_std, multiplier, n, abs_val, abs_std, rel_val, rel_std);
    else difference_kernel<false><<<grid, 256, 0, s>>>(x_val, x_std, y_val, y_std, multiplier, n, abs_val, abs_std, rel_val, rel_std);
    return launched();
}

}  // extern "C"
