// irbpp_contour.cuh -- per-thread candidate extraction on one 16x16 level image.
//
// Device restatement of what the reference obtains from OpenCV at
//   cvTools.py:86  cv2.findContours(check, RETR_TREE, CHAIN_APPROX_SIMPLE)  + find_out_contour (:7-38)
//   cvTools.py:91  cv2.approxPolyDP(contour, 1, True)
//   cvTools.py:92  find_convex_vetex (:40-59)
// (paths relative to the reference root).  Algorithms: Suzuki-Abe border following with the
// CHAIN_APPROX_SIMPLE emission rule; Douglas-Peucker with OpenCV's closed-curve seeding and
// clean-up pass.  All arithmetic is integer (coordinates are 0..15), so there is nothing to round:
// the float comparisons OpenCV performs are between exactly representable values and are restated
// as integer cross-multiplications.
//
// The code is templated on a scratch accessor so the same routine serves the fast path (one thread
// per (rotation, level) task, <=64-point contours, kept-set in a 64-bit register) and the overflow
// path (one thread, 1024-point buffers in shared memory).
#pragma once
#include <stdint.h>

namespace irbpp {

// direction codes, image y grows downward: 0=E 1=NE 2=N 3=NW 4=W 5=SW 6=S 7=SE
__device__ __forceinline__ int ddx(int s) {
    // dx+1 = {2,2,1,0,0,0,1,2} packed 2 bits each, s = 0 in the low bits
    return (int)((0x901Au >> (2 * s)) & 3u) - 1;
}
__device__ __forceinline__ int ddy(int s) {
    // dy+1 = {1,0,0,0,1,2,2,2}
    return (int)((0xA901u >> (2 * s)) & 3u) - 1;
}

// ---- scratch accessors -------------------------------------------------------------------------
// Row words use PADDED x: bit (x+1) is column x, bits 0 and 17 are the zero frame; rows 0 and 17 of
// fg are the zero frame.  Marks hold "positive" (low half) and "negative" (high half) Suzuki labels
// per unpadded column.

template <int STRIDE, int CAP_>
struct StridedScratch {
    static constexpr int CAP = CAP_;
    uint32_t* w;   // 18 fg rows then 16 mark rows, element stride STRIDE
    uint8_t* b;    // CAP contour points then CAP result points, element stride STRIDE
    uint64_t kept;
    __device__ __forceinline__ uint32_t fg(int y) const { return w[y * STRIDE]; }
    __device__ __forceinline__ void set_fg(int y, uint32_t v) { w[y * STRIDE] = v; }
    __device__ __forceinline__ uint32_t mk(int y) const { return w[(18 + y) * STRIDE]; }
    __device__ __forceinline__ void set_mk(int y, uint32_t v) { w[(18 + y) * STRIDE] = v; }
    __device__ __forceinline__ int pt(int i) const { return b[i * STRIDE]; }
    __device__ __forceinline__ void set_pt(int i, int v) { b[i * STRIDE] = (uint8_t)v; }
    __device__ __forceinline__ int ds(int i) const { return b[(CAP + i) * STRIDE]; }
    __device__ __forceinline__ void set_ds(int i, int v) { b[(CAP + i) * STRIDE] = (uint8_t)v; }
    __device__ __forceinline__ void kept_clear(int) { kept = 0ull; }
    __device__ __forceinline__ void kept_set(int i) { kept |= (1ull << i); }
    __device__ __forceinline__ int kept_count(int) const { return __popcll(kept); }
    // next kept index strictly after i, cyclically
    __device__ __forceinline__ int kept_next(int i, int) const {
        uint64_t hi = (i >= 63) ? 0ull : (kept & ~((2ull << i) - 1ull));
        return hi ? (__ffsll((long long)hi) - 1) : (__ffsll((long long)kept) - 1);
    }
};

template <int CAP_>
struct FlatScratch {
    static constexpr int CAP = CAP_;
    uint32_t* w;     // 18 + 16 words
    uint8_t* b;      // CAP points, CAP result points, CAP kept flags
    __device__ __forceinline__ uint32_t fg(int y) const { return w[y]; }
    __device__ __forceinline__ void set_fg(int y, uint32_t v) { w[y] = v; }
    __device__ __forceinline__ uint32_t mk(int y) const { return w[18 + y]; }
    __device__ __forceinline__ void set_mk(int y, uint32_t v) { w[18 + y] = v; }
    __device__ __forceinline__ int pt(int i) const { return b[i]; }
    __device__ __forceinline__ void set_pt(int i, int v) { b[i] = (uint8_t)v; }
    __device__ __forceinline__ int ds(int i) const { return b[CAP + i]; }
    __device__ __forceinline__ void set_ds(int i, int v) { b[CAP + i] = (uint8_t)v; }
    __device__ __forceinline__ void kept_clear(int n) { for (int i = 0; i < n; ++i) b[2 * CAP + i] = 0; }
    __device__ __forceinline__ void kept_set(int i) { b[2 * CAP + i] = 1; }
    __device__ __forceinline__ int kept_count(int n) const { int c = 0; for (int i = 0; i < n; ++i) c += b[2 * CAP + i]; return c; }
    __device__ __forceinline__ int kept_next(int i, int n) const {
        int k = i;
        do { k = (k + 1 == n) ? 0 : k + 1; } while (!b[2 * CAP + k] && k != i);
        return k;
    }
};

// ---- border following (Suzuki-Abe, CHAIN_APPROX_SIMPLE) -----------------------------------------
// (x0, y0) in padded coordinates (1..16).  Marks every visited border pixel; stores the emitted
// points as (x<<4 | y) unpadded.  Returns the number of points, or -1 if it exceeded S::CAP (marks
// are still complete in that case).
template <class S>
__device__ int follow_border(S& sc, int x0, int y0, bool hole) {
    auto pix = [&](int x, int y) -> bool { return (sc.fg(y) >> x) & 1u; };
    auto mark_neg = [&](int x, int y) { sc.set_mk(y - 1, sc.mk(y - 1) | (0x10000u << (x - 1))); };
    auto mark_pos_if_unmarked = [&](int x, int y) {
        uint32_t m = sc.mk(y - 1);
        if (!((m | (m >> 16)) & (1u << (x - 1)))) sc.set_mk(y - 1, m | (1u << (x - 1)));
    };
    int s_end = hole ? 0 : 4;
    int s = s_end;
    bool found = false;
    do {
        s = (s - 1) & 7;
        if (pix(x0 + ddx(s), y0 + ddy(s))) { found = true; break; }
    } while (s != s_end);
    if (!found) {  // isolated pixel
        mark_neg(x0, y0);
        sc.set_pt(0, ((x0 - 1) << 4) | (y0 - 1));
        return 1;
    }
    const int x1 = x0 + ddx(s), y1 = y0 + ddy(s);
    int x3 = x0, y3 = y0;
    int prev_s = s ^ 4;
    int n = 0;
    bool ovf = false;
    for (;;) {
        s_end = s;
        int x4, y4;
        for (;;) {
            ++s;
            x4 = x3 + ddx(s & 7);
            y4 = y3 + ddy(s & 7);
            if (pix(x4, y4)) break;
        }
        s &= 7;
        if ((unsigned)(s - 1) < (unsigned)s_end) mark_neg(x3, y3);
        else mark_pos_if_unmarked(x3, y3);
        if (s != prev_s) {
            if (n < S::CAP) sc.set_pt(n, ((x3 - 1) << 4) | (y3 - 1));
            else ovf = true;
            ++n;
        }
        prev_s = s;
        if (x4 == x0 && y4 == y0 && x3 == x1 && y3 == y1) break;
        x3 = x4; y3 = y4;
        s = (s + 4) & 7;
    }
    return ovf ? -1 : n;
}

// ---- approxPolyDP(eps = 1, closed) + convex-vertex filter ---------------------------------------
// Emits the selected vertices through `emit(x, y)`.
template <class S, class Emit>
__device__ void approx_and_emit(S& sc, int n, bool legacy, Emit emit) {
    auto PX = [&](int i) { return sc.pt(i) >> 4; };
    auto PY = [&](int i) { return sc.pt(i) & 15; };
    // 1. seed: three rounds of "farthest point from the current one"
    int pos = 0, far = 0, maxd = 0;
    for (int it = 0; it < 3; ++it) {
        pos += far; if (pos >= n) pos -= n;
        const int sx = PX(pos), sy = PY(pos);
        maxd = 0; far = 0;
        int k = pos;
        for (int j = 1; j < n; ++j) {
            k = (k + 1 == n) ? 0 : k + 1;
            const int ex = PX(k) - sx, ey = PY(k) - sy;
            const int d = ex * ex + ey * ey;
            if (d > maxd) { maxd = d; far = j; }
        }
    }
    if (maxd <= 1) {  // whole contour within eps of one point
        emit(PX(pos), PY(pos));
        return;
    }
    int fp = pos + far; if (fp >= n) fp -= n;
    // 2. Douglas-Peucker.  A leaf slice (s,e) keeps P[s]; the kept set is order independent, so the
    // larger half is deferred and the explicit stack stays logarithmic.
    sc.kept_clear(n);
    int st_s[16], st_e[16];
    int sp = 0;
    st_s[sp] = fp; st_e[sp] = pos; ++sp;
    st_s[sp] = pos; st_e[sp] = fp; ++sp;
    while (sp > 0) {
        --sp;
        int s = st_s[sp], e = st_e[sp];
        for (;;) {
            int len = e - s; if (len <= 0) len += n;
            if (len == 1) { sc.kept_set(s); break; }
            const int sx = PX(s), sy = PY(s);
            const int dx = PX(e) - sx, dy = PY(e) - sy;
            const int seg2 = dx * dx + dy * dy;
            int best = -1, bi = s;
            int k = s;
            for (int t = 1; t < len; ++t) {
                k = (k + 1 == n) ? 0 : k + 1;
                const int vx = PX(k) - sx, vy = PY(k) - sy;
                int num;
                if (legacy) {
                    const int cr = vy * dx - vx * dy;
                    num = cr < 0 ? -cr : cr;                 // |cross| (common factor 1/|seg|)
                } else {
                    const int dot = vx * dx + vy * dy;
                    if (seg2 == 0) num = vx * vx + vy * vy;  // degenerate segment: plain distance^2
                    else if (dot <= 0) num = (vx * vx + vy * vy) * seg2;
                    else if (dot >= seg2) { const int wx = vx - dx, wy = vy - dy; num = (wx * wx + wy * wy) * seg2; }
                    else { const int cr = vy * dx - vx * dy; num = cr * cr; }   // dist^2 * seg2
                }
                if (num > best) { best = num; bi = k; }
            }
            bool le;
            if (legacy) le = (best * best <= seg2);
            else le = (seg2 == 0) ? (best <= 1) : (best <= seg2);
            if (le) { sc.kept_set(s); break; }
            int ll = bi - s; if (ll <= 0) ll += n;
            const int lr = len - ll;
            if (ll <= lr) { st_s[sp] = bi; st_e[sp] = e; ++sp; e = bi; }
            else          { st_s[sp] = s;  st_e[sp] = bi; ++sp; s = bi; }
        }
    }
    // 3. ring Q = kept points in contour order starting at pos; clean-up of nearly collinear points
    const int c = sc.kept_count(n);
    {
        int ci = pos;
        for (int k = 0; k < c; ++k) { sc.set_ds(k, sc.pt(ci)); ci = sc.kept_next(ci, n); }
    }
    int new_count = c;
    if (c > 2) {
        int last = pos;
        for (int k = 0; k + 1 < c; ++k) last = sc.kept_next(last, n);
        int start = sc.pt(last);
        int ci = pos;
        int pt = sc.pt(ci); ci = sc.kept_next(ci, n);
        int wpos = 0;
        int i = 0;
        while (i < c && new_count > 2) {
            const int end = sc.pt(ci); ci = sc.kept_next(ci, n);
            const int dx = (end >> 4) - (start >> 4), dy = (end & 15) - (start & 15);
            const int px = (pt >> 4) - (start >> 4), py = (pt & 15) - (start & 15);
            int dist = px * dy - py * dx; if (dist < 0) dist = -dist;
            const int ip = px * ((end >> 4) - (pt >> 4)) + py * ((end & 15) - (pt & 15));
            if (2 * dist * dist <= dx * dx + dy * dy && dx != 0 && dy != 0 && ip >= 0) {
                --new_count;
                start = end;
                sc.set_ds(wpos, end); wpos = (wpos + 1 == c) ? 0 : wpos + 1;
                pt = sc.pt(ci); ci = sc.kept_next(ci, n);
                i += 2;
                continue;
            }
            start = pt;
            sc.set_ds(wpos, pt); wpos = (wpos + 1 == c) ? 0 : wpos + 1;
            pt = end;
            ++i;
        }
    }
    // 4. find_convex_vetex: all points if <= 3, else strictly clockwise turns (cross < 0)
    if (new_count <= 3) {
        for (int k = 0; k < new_count; ++k) { const int p = sc.ds(k); emit(p >> 4, p & 15); }
    } else {
        int a = sc.ds(new_count - 1), bpt = sc.ds(0);
        for (int k = 0; k < new_count; ++k) {
            const int cpt = sc.ds(k + 1 == new_count ? 0 : k + 1);
            const int abx = (bpt >> 4) - (a >> 4), aby = (bpt & 15) - (a & 15);
            const int acx = (cpt >> 4) - (a >> 4), acy = (cpt & 15) - (a & 15);
            if (abx * acy - aby * acx < 0) emit(bpt >> 4, bpt & 15);
            a = bpt; bpt = cpt;
        }
    }
}

// ---- one level image: raster scan for border starts, follow, approximate, emit --------------------
// rows16[y] (y = 0..15) holds the unpadded 16-bit row y of the level image.  Returns false if some
// outer contour overflowed S::CAP points (the caller re-runs the task on the overflow path; emitting
// the other contours twice is harmless because emission is a set union).
template <class S, class RowFn, class Emit>
__device__ bool process_level_image(S& sc, RowFn rows16, bool legacy, Emit emit) {
    sc.set_fg(0, 0u);
    sc.set_fg(17, 0u);
    for (int y = 0; y < 16; ++y) { sc.set_fg(y + 1, (rows16(y) & 0xFFFFu) << 1); sc.set_mk(y, 0u); }
    bool ok = true;
    for (int y = 1; y <= 16; ++y) {
        uint32_t window = 0x1FFFEu;            // padded columns still to visit in this row
        const uint32_t f = sc.fg(y);
        if (!f) continue;
        for (;;) {
            const uint32_t m = sc.mk(y - 1);
            const uint32_t posP = (m & 0xFFFFu) << 1, negP = (m >> 16) << 1;
            const uint32_t outer = (f & ~posP & ~negP) & ~(f << 1);   // label 1 and left neighbour 0
            const uint32_t hole = (f & ~negP) & ~(f >> 1);            // label >= 1 and right neighbour 0
            const uint32_t c = (outer | hole) & window;
            if (!c) break;
            const int x = __ffs((int)c) - 1;
            const bool is_hole = !((outer >> x) & 1u);
            const int n = follow_border(sc, x, y, is_hole);
            if (!is_hole) {
                if (n < 0) ok = false;
                else approx_and_emit(sc, n, legacy, emit);
            }
            window = (x >= 16) ? 0u : (0x1FFFEu & ~((2u << x) - 1u));
        }
    }
    return ok;
}

}  // namespace irbpp
