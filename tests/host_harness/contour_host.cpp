// TEST-ONLY harness: compiles the per-thread device routines of
// irbpp_b200/csrc/irbpp_contour.cuh (and the exact floor-divide of irbpp_math.cuh) as plain host
// C++ so their logic can be checked against the oracle without a GPU.  It is never linked into the
// product library and is not a CPU fallback: nothing under irbpp_b200/ references it.
#include <math.h>
#include <stdint.h>
#define __device__
#define __host__
#define __forceinline__ inline
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __any_sync(unsigned, int p) { return p; }   // one "lane": the lock-step variant degenerates to serial
static inline long long clock64() { return 0; }
static inline void __syncwarp() {}
#include "../../irbpp_b200/csrc/irbpp_contour.cuh"
#include "../../irbpp_b200/csrc/irbpp_math.cuh"

extern "C" int hull_bits(const uint16_t* rows16, int legacy, int mode, uint32_t* out_bits) {
    for (int i = 0; i < 8; ++i) out_bits[i] = 0;
    uint32_t rows[irbpp::ROWS_WORDS] = {0};                  // padded row form: rows[y + 1] = row y << 1
    for (int y = 0; y < 16; ++y) rows[y + 1] = (uint32_t)rows16[y] << 1;
    auto emit = [&](int x, int y) { const int b = x * 16 + y; out_bits[b >> 5] |= 1u << (b & 31); };
    if (mode == 5) {   // one task per start pixel, long buffers (the kernel's overflow path)
        static uint8_t b[2 * 1024];
        irbpp::FlatScratch<1024> sc; sc.b = b;
        return irbpp::process_level_image_mt(sc, rows, legacy != 0, emit) ? 0 : 1;
    }
    if (mode == 6) {   // one task per start pixel, 64-point fast buffers (the kernel's lane scratch)
        static uint8_t b[64];
        irbpp::StridedScratch<1, 64> sc; sc.b = b; sc.kept = 0;
        return irbpp::process_level_image_mt(sc, rows, legacy != 0, emit) ? 0 : 1;
    }
    return -1;   // modes 5 and 6 only
}

extern "C" void floor_div_many(const double* a, double b, int n, double* out) {
    const double inv = 1.0 / b;
    for (int i = 0; i < n; ++i) out[i] = irbpp::floor_divide_exact(a[i], b, inv);
}

#include "../../irbpp_b200/csrc/irbpp_heuristic.cuh"

// np.sum of n contiguous doubles as irbpp_math.cuh restates it (checked against NumPy itself)
extern "C" double pairwise_sum_host(const double* a, int n) {
    auto at = [&](int i) { return a[i]; };
    return irbpp::np_pairwise_sum(at, n);
}

// argmin pose of one bin exactly as irbpp_heuristic_kernel scores it (serial): hm row-major [32][32],
// posz / mask [R][256], top tables with -inf where maskT == 0
extern "C" int heuristic_pose_host(int method, int dir_idx, int R, const double* hm, const double* posz,
                                   const uint8_t* mask, const double* Ts, const int64_t* offs, const int32_t* w,
                                   const int32_t* h, double resA) {
    auto hm_at = [&](int x, int y) { return hm[x * 32 + y]; };
    double best = INFINITY;
    int beste = 0;
    for (int e = 0; e < R * 256; ++e) {
        const int r = e >> 8, p = e & 255;
        double s = irbpp::HEUR_INVALID;
        if (mask[e])
            s = irbpp::heuristic_score(method, dir_idx, p >> 4, p & 15, posz[e], resA, 16, 16, 2, hm_at,
                                       Ts + offs[r], w[r], h[r]);
        if (s < best) { best = s; beste = e; }
    }
    return beste;
}

// All outer contours of an image as the device follower produces them: for every start pixel whose path
// is kept (raster-first start of an outer border), n followed by n packed points (x << 4 | y), in
// raster order of the start pixels.  Returns the number of words written, -1 on overflow of `cap`.
extern "C" int outer_contours_host(const uint16_t* rows16, int32_t* out, int cap) {
    uint32_t rows[irbpp::ROWS_WORDS] = {0};
    for (int y = 0; y < 16; ++y) rows[y + 1] = (uint32_t)rows16[y] << 1;
    static uint8_t b[2 * 1024];
    irbpp::FlatScratch<1024> sc; sc.b = b;
    int w = 0;
    for (int y = 0; y < 16; ++y) {
        uint32_t c = irbpp::start_candidates_rows(rows, y);
        while (c) {
            const int x = __builtin_ctz(c);
            c &= c - 1;
            int area2;
            const int n = irbpp::follow_outer_rows(sc, rows, x, y, area2);
            if (n == -2 || area2 > 0) continue;
            if (n < 0 || w + 1 + n > cap) return -1;
            out[w++] = n;
            for (int i = 0; i < n; ++i) out[w++] = sc.pt(i);
        }
    }
    return w;
}
