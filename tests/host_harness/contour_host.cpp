// TEST-ONLY harness: compiles the per-thread device routines of
// ir-bpp_b200/csrc/irbpp_contour.cuh as plain host C++ so their integer logic can be checked
// against the oracle without a GPU.  It is never linked into the product library and is not a
// CPU fallback: nothing under ir-bpp_b200/ references it.
#include <stdint.h>
#define __device__
#define __host__
#define __forceinline__ inline
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
#include "../../ir-bpp_b200/csrc/irbpp_contour.cuh"

extern "C" int hull_bits(const uint16_t* rows, int legacy, int use_big, uint32_t* out_bits) {
    for (int i = 0; i < 8; ++i) out_bits[i] = 0;
    auto rowfn = [&](int y) { return (uint32_t)rows[y]; };
    auto emit = [&](int x, int y) { const int b = x * 16 + y; out_bits[b >> 5] |= 1u << (b & 31); };
    if (use_big) {
        static uint32_t w[34]; static uint8_t b[3 * 1024];
        irbpp::FlatScratch<1024> sc; sc.w = w; sc.b = b;
        return irbpp::process_level_image(sc, rowfn, legacy != 0, emit) ? 0 : 1;
    }
    static uint32_t w[34]; static uint8_t b[2 * 64];
    irbpp::StridedScratch<1, 64> sc; sc.w = w; sc.b = b; sc.kept = 0;
    return irbpp::process_level_image(sc, rowfn, legacy != 0, emit) ? 0 : 1;
}
