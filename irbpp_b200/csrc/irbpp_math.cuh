// irbpp_math.cuh -- float64 helpers whose results must equal NumPy's bit for bit.
#pragma once
#include <math.h>

namespace irbpp {

// NumPy float64 floor_divide a // b (npy_divmod, used at reference cvTools.py:78) for finite a and
// b > 0.  NumPy computes mod = fmod(a, b) (exact), div = (a - mod) / b, then snaps div to the nearest
// integer; the result is floor(a / b) of the REAL quotient of the two doubles (e.g. 0.03 // 0.01 == 2
// because the double 0.03 is below 3 * the double 0.01).  The same integer is obtained here without
// fmod: an estimate from a * (1/b), corrected with fused multiply-adds, whose single rounding cannot
// change the sign of the exact residual a - k*b.
__host__ __device__ __forceinline__ double floor_divide_exact(double a, double b, double inv_b) {
    double k = floor(a * inv_b);
    if (fma(-k, b, a) < 0.0) k -= 1.0;              // a < k*b  -> estimate one too high
    else if (fma(-(k + 1.0), b, a) >= 0.0) k += 1.0;  // a >= (k+1)*b -> one too low
    return k;
}

// np.round(v, 6) <= 0   <=>   rint(v * 1e6) <= 0   (np.round multiplies, rints, divides)
__host__ __device__ __forceinline__ bool round6_le0(double v) { return rint(v * 1e6) <= 0.0; }

// np.round(v, 6): multiply by 1e6, rint, divide by 1e6 (numpy/_core/fromnumeric.py round -> ndarray.round)
__host__ __device__ __forceinline__ double np_round6(double v) { return rint(v * 1e6) / 1e6; }

// NumPy's float64 add-reduce of n contiguous elements (np.sum over a fresh C-contiguous array;
// reference space.py:217 `np.sum(heightmapC_Prime)`): pairwise summation with 8 interleaved
// accumulators on blocks of at most 128 elements, halves rounded down to a multiple of 8 above
// that.  `a(i)` returns element i.  The association order is what makes the sum bit-identical.
template <class F>
__host__ __device__ inline double np_pairwise_block(F& a, int lo, int n) {      // n <= 128
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res += a(lo + i);
        return res;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a(lo + j);
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] += a(lo + i + j);
    }
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a(lo + i);
    return res;
}

template <class F>
__host__ __device__ inline double np_pairwise_sum(F& a, int n) {
    // explicit stack instead of recursion: post-order evaluation of the split tree
    constexpr int DEPTH = 12;      // n <= 128 * 2^10
    int lo_st[DEPTH], n_st[DEPTH];
    double val_st[DEPTH];
    unsigned char state[DEPTH];      // 0: not expanded, 1: left done (val holds it)
    int sp = 0;
    lo_st[0] = 0; n_st[0] = n; state[0] = 0;
    double ret = 0.0;
    bool have_ret = false;
    while (sp >= 0) {
        if (!have_ret) {
            if (n_st[sp] <= 128) { ret = np_pairwise_block(a, lo_st[sp], n_st[sp]); have_ret = true; --sp; continue; }
            int n2 = n_st[sp] / 2; n2 -= n2 % 8;
            state[sp] = 0;
            lo_st[sp + 1] = lo_st[sp]; n_st[sp + 1] = n2; state[sp + 1] = 0;
            ++sp;
        } else {
            if (state[sp] == 0) {          // left child returned: evaluate the right one
                val_st[sp] = ret; state[sp] = 1; have_ret = false;
                int n2 = n_st[sp] / 2; n2 -= n2 % 8;
                lo_st[sp + 1] = lo_st[sp] + n2; n_st[sp + 1] = n_st[sp] - n2; state[sp + 1] = 0;
                ++sp;
            } else { ret = val_st[sp] + ret; --sp; }
        }
    }
    return ret;
}

}  // namespace irbpp
