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

}  // namespace irbpp
