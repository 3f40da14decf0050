// irbpp_kernels.cuh -- the fused packing-environment kernel for sm_100a.
//
// One CTA (128 threads) owns one bin for one call.  Phases, all on shared-memory resident state:
//   A  apply the chosen candidate: decode (binPhy.py:234-236), prejudge (:238-245), drop height of that
//      pose, placement test (Interface.py:365-369 semantics), heightmap update (space.py:213 closed
//      form), reward / episode bookkeeping (binPhy.py:299-327, monitor.py:58-75), item queue
//      (IRcreator.py:6-24), auto-reset on failure (shmem_vec_env.py:140-144)
//   B  scan every (rotation, X, Y) pose for the next item: drop height + feasibility
//      (space.py:98-129), one warp per rotation, lanes = poses
//   C  candidate extraction (cvTools.py:61-103): level quantisation with NumPy floor_divide semantics,
//      per-level bitmaps by warp ballots, one thread per (rotation, level) image doing border
//      following + approxPolyDP + convex filter (irbpp_contour.cuh), results OR-ed into a 256-bit set
//      per rotation (np.unique == sorted set)
//   D  select / pad (binPhy.py:205-225), observation assembly (binPhy.py:183-232) written as float32
//      (envs.py:151,163 cast) with coalesced stores; state written back
// (paths relative to the reference root)
//
// Arithmetic is IEEE float64 exactly as NumPy performs it (compile with -fmad=false); masks are
// folded into the tables as +/-inf sentinels at load time, which changes no comparison result.
#pragma once
#include <stdint.h>
#include <math.h>
#include "irbpp_contour.cuh"
#include "irbpp_math.cuh"

namespace irbpp {

constexpr int HX = 32, HY = 32;          // heightmap cells (rangeX_C, rangeY_C)
constexpr int AX = 16, AY = 16;          // action grid (rangeX_A, rangeY_A)
constexpr int STEP = 2;                  // stepSize = resolutionAct / resolutionH
constexpr int NPOSE = AX * AY;           // 256 poses per rotation
constexpr int CTA_THREADS = 128;
constexpr int CTA_WARPS = CTA_THREADS / 32;
#ifndef IRBPP_QUOTA
#define IRBPP_QUOTA 16
#endif
constexpr int QUOTA = IRBPP_QUOTA;       // level images a warp (= rotation) contributes per round
constexpr int NSLOT = CTA_WARPS * QUOTA; // level-image tasks per round (and per-thread scratch slots)
#ifndef IRBPP_TRACE_WARPS
#define IRBPP_TRACE_WARPS 2
#endif
constexpr int TRACE_WARPS = IRBPP_TRACE_WARPS;          // warps that run the level-image tasks
constexpr int TRACE_LANES = NSLOT / TRACE_WARPS;        // task lanes per tracing warp (<= 32)
constexpr int TRACE_WARPS_DIV = 1;
static_assert(TRACE_LANES <= 32 && TRACE_LANES * TRACE_WARPS == NSLOT, "task lanes must tile the slots");
constexpr int SLOT_WORDS = 9;            // 8 bitmap words + 1 pad (bank spread)
constexpr int FAST_CAP = 64;             // contour points on the fast path
constexpr int BIG_CAP = 1024;            // contour points on the overflow path
constexpr int LEVEL_OFFSET = 32;         // levels in [-32, 31] -> presence bit (level + 32)
constexpr int MAX_QUEUE = 16;            // buffer_size limit
constexpr double POSZ_INVALID = 1e3;     // space.py:101,126

enum Mode : int {
    MODE_RESET = 0,        // reset selected envs, emit observation
    MODE_STEP = 1,         // phase A then observation (online: B-D for queue[0]; buffered: order obs)
    MODE_CANDIDATES = 2,   // get_action_candidates(order): B-D for queue[order]
    MODE_ALL_OBS = 3,      // get_all_possible_observation: blockIdx.y = queue slot
    MODE_DEBUG_SCAN = 4,   // B-D for a caller-supplied item, dumps float64 views, no state change
    MODE_DEBUG_HULLS = 5,  // C-D on caller-supplied posZValid / mask
};

struct ShapeRot {          // one (shape, rotation) entry, device resident
    int32_t w, h;          // window in heightmap cells (rangeX_OH, rangeY_OH; space.py:105)
    int32_t nX, nY;        // number of X / Y positions scanned: A - ceil(ext/resA) + 1 (space.py:115-116)
    uint32_t okx, oky;     // bit lx set <=> prejudge passes in x / y for that lx (binPhy.py:240-241)
    int32_t any_zero;      // maskB has a zero cell -> the window max includes a 0 term
    int32_t pad;
    double ez;             // round(extent_z, 6)   (space.py:104,120)
    int64_t off;           // offset of Bs / Ts of this entry in the pools (doubles)
};

struct Params {
    // configuration
    int32_t N, R, sel, K;                // K = buffer_size (1 = online)
    int32_t loc_len, order_len, obs_stride;
    int32_t legacy;
    double binz, resZ, binvol;
    // shapes
    int32_t S;
    const ShapeRot* srot;                // [S*R]
    const double* Bs;                    // bottom tables, +inf where maskB == 0
    const double* Ts;                    // top tables, -inf where maskT == 0
    const double* vol;                   // [S]
    const double* reward_tab;            // [S] (vol / binvol) * 10
    // sequences
    const int32_t* seq; int32_t L;
    // per-env state
    double* hm;                          // [N][2][32][16] column-parity planes
    uint16_t* cand;                      // [N][sel] rot<<8 | x<<4 | y
    int32_t* queue;                      // [N][MAX_QUEUE]
    int32_t* cursor; int32_t* cur_item; int32_t* order_act; int32_t* packed; int32_t* ep_len;
    double* vol_sum; double* ep_rew;
    uint8_t* mask_any;
    // inputs of this call
    const int64_t* actions;              // MODE_STEP / MODE_CANDIDATES
    const uint8_t* which;                // MODE_RESET (NULL = all)
    const int32_t* dbg_items;            // MODE_DEBUG_SCAN
    const double* dbg_in_posz; const double* dbg_in_mask;   // MODE_DEBUG_HULLS
    // outputs
    float* obs;                          // [N][obs_stride] (+ slot offset in MODE_ALL_OBS)
    float* r_reward; uint8_t* r_done; uint8_t* r_valid; uint8_t* r_error;
    int32_t* r_counter; int32_t* r_eplen; double* r_ratio; double* r_eprew;
    double* dbg_posz; double* dbg_poszv; double* dbg_mask; double* dbg_cand; int32_t* dbg_nhull;
    unsigned long long* phase_cycles;    // [8] summed SM cycles per phase (thread 0 of every CTA), or NULL
    int32_t mode;
};

// ---- shared memory carve-up -----------------------------------------------------------------------
// Region X is time-multiplexed: heightmap (phases A, B) -> per-thread contour scratch (phase C) ->
// float32 observation staging (phase D).
struct SmemLayout {
    int x_size, posz_off, slots_off, maskbits_off, candbits_off, misc_off, big_off, total;
};

constexpr int SCRATCH_BYTES = NSLOT * (16 * 4 + FAST_CAP);     // marks + contour points per task thread

__host__ __device__ inline SmemLayout smem_layout(int R, int sel) {
    SmemLayout L;
    int x = 2 * HX * (HY / 2) * 8;                                     // heightmap, 8192
    if (x < SCRATCH_BYTES) x = SCRATCH_BYTES;
    const int stage_need = sel * 5 * 4 + sel * 2 + 16;
    if (x < stage_need) x = stage_need;
    x = (x + 15) & ~15;
    L.x_size = x;
    int o = x;
    L.posz_off = o; o += R * NPOSE * 8;
    L.slots_off = o; o += NSLOT * SLOT_WORDS * 4;
    L.maskbits_off = o; o += R * 8 * 4;
    L.candbits_off = o; o += R * 8 * 4;
    L.misc_off = o; o += 384;
    L.big_off = o; o += 16 * 4 + 2 * BIG_CAP;
    L.total = (o + 15) & ~15;
    return L;
}

struct Misc {                    // small CTA-wide scalars in shared memory (<= 384 bytes)
    int32_t nlev[CTA_WARPS];
    int32_t ovf_count;
    int32_t ovf_task[32];
    int32_t cnt_prefix[33];      // candidate count prefix over rotations (R <= 32)
    int32_t error;
    int32_t item;
    int32_t any_mask;
};

__device__ __forceinline__ int hm_index(int x, int y) { return ((y & 1) * HX + x) * (HY / 2) + (y >> 1); }

__device__ __forceinline__ int draw_item(const Params& P, int env, int& cursor) {
    int id = P.seq[(int64_t)env * P.L + (cursor % P.L)];
    ++cursor;
    return id;
}

// ---- phase B: one warp scans one rotation -----------------------------------------------------------
// Writes posz[r][256] and maskbits[r][8] (space.py:98-129).
__device__ __forceinline__ void scan_rotation(const Params& P, const double* hm_s, double* posz_s,
                                              uint32_t* maskbits_s, int item, int r, int lane) {
    const ShapeRot* sr = P.srot + (int64_t)item * P.R + r;
    const int w = sr->w, h = sr->h, nX = sr->nX, nY = sr->nY;
    const double ez = sr->ez;
    const double init = sr->any_zero ? 0.0 : -INFINITY;
    const double* __restrict__ B = P.Bs + sr->off;
    const int hpairs = h >> 1;
#pragma unroll 1
    for (int pass = 0; pass < 8; ++pass) {
        const int p = pass * 32 + lane;
        const int X = p >> 4, Y = p & 15;
        const bool valid = (X < nX) && (Y < nY);
        double acc = POSZ_INVALID;
        bool feas = false;
        if (valid) {
            acc = init;
            const double* h0 = hm_s + (STEP * X) * (HY / 2) + Y;        // even heightmap columns
            const double* brow = B;
            for (int i = 0; i < w; ++i) {
                const double* h1 = h0 + HX * (HY / 2);                   // odd heightmap columns
                int jj = 0;
                for (; jj < hpairs; ++jj) {
                    const double v0 = h0[jj] - __ldg(brow + 2 * jj);
                    const double v1 = h1[jj] - __ldg(brow + 2 * jj + 1);
                    acc = (v0 > acc) ? v0 : acc;
                    acc = (v1 > acc) ? v1 : acc;
                }
                if (h & 1) {
                    const double v0 = h0[jj] - __ldg(brow + 2 * jj);
                    acc = (v0 > acc) ? v0 : acc;
                }
                h0 += HY / 2;
                brow += h;
            }
            feas = round6_le0(acc + ez - P.binz);
        }
        posz_s[r * NPOSE + p] = acc;
        const uint32_t mb = __ballot_sync(0xffffffffu, feas);
        if (lane == 0) maskbits_s[r * 8 + pass] = mb;
    }
}

// level quantisation (cvTools.py:78-79): lv[pass] = posZ // resZ for feasible poses, -1 otherwise;
// `present` gets bit (level + LEVEL_OFFSET) for every level in this rotation
__device__ __forceinline__ void levels_from_maps(const Params& P, const double* posz_s, const uint32_t* maskbits_s,
                                                 int r, int lane, int (&lv)[8], uint64_t& present, int& err) {
    uint32_t pres_lo = 0, pres_hi = 0;
    const double inv = 1.0 / P.resZ;
#pragma unroll
    for (int pass = 0; pass < 8; ++pass) {
        const int p = pass * 32 + lane;
        const bool feas = (maskbits_s[r * 8 + pass] >> lane) & 1u;
        int L = -1;
        if (feas) {
            L = (int)floor_divide_exact(posz_s[r * NPOSE + p], P.resZ, inv);
            if (L < -LEVEL_OFFSET || L >= LEVEL_OFFSET) { err = 1; L = -1; }
            if (L != -1) {
                const int b = L + LEVEL_OFFSET;
                if (b < 32) pres_lo |= 1u << b; else pres_hi |= 1u << (b - 32);
            }
        }
        lv[pass] = L;
    }
    pres_lo = __reduce_or_sync(0xffffffffu, pres_lo);
    pres_hi = __reduce_or_sync(0xffffffffu, pres_hi);
    present = ((uint64_t)pres_hi << 32) | pres_lo;
}

// up to QUOTA 16x16 bitmaps (8 words, two rows each) for the next levels in `present`, by warp ballots;
// consumed levels are removed from `present`
__device__ __forceinline__ int build_level_bitmaps(uint32_t* slots_s, int warp, int lane, const int (&lv)[8],
                                                   uint64_t& present) {
    int k = 0;
    while (present && k < QUOTA) {
        const int b = __ffsll((long long)present) - 1;
        present &= present - 1;
        const int L = b - LEVEL_OFFSET;
        uint32_t mine = 0;
#pragma unroll
        for (int pass = 0; pass < 8; ++pass) {
            const uint32_t bits = __ballot_sync(0xffffffffu, lv[pass] == L);
            if (lane == pass) mine = bits;
        }
        if (lane < 8) slots_s[(warp * QUOTA + k) * SLOT_WORDS + lane] = mine;
        ++k;
    }
    return k;
}

// ---- the kernel -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CTA_THREADS, 8) irbpp_env_kernel(const Params P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int env = blockIdx.x;
    const int slot = blockIdx.y;                 // MODE_ALL_OBS only
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const SmemLayout SL = smem_layout(P.R, P.sel);
    double* hm_s = reinterpret_cast<double*>(smem_raw);                      // region X, phases A-B
    unsigned char* scratch_s = smem_raw;                                      // region X, phase C
    float* stage_f = reinterpret_cast<float*>(smem_raw);                      // region X, phase D
    uint16_t* stage_c = reinterpret_cast<uint16_t*>(smem_raw + P.sel * 5 * 4);
    double* posz_s = reinterpret_cast<double*>(smem_raw + SL.posz_off);
    uint32_t* slots_s = reinterpret_cast<uint32_t*>(smem_raw + SL.slots_off);
    uint32_t* maskbits_s = reinterpret_cast<uint32_t*>(smem_raw + SL.maskbits_off);
    uint32_t* candbits_s = reinterpret_cast<uint32_t*>(smem_raw + SL.candbits_off);
    unsigned char* big_s = smem_raw + SL.big_off;
    Misc* misc = reinterpret_cast<Misc*>(smem_raw + SL.misc_off);

    const int mode = P.mode;
    if (mode == MODE_RESET && P.which && !P.which[env]) return;
    long long t_prev = P.phase_cycles ? clock64() : 0;
    auto phase_mark = [&](int idx) {
        if (P.phase_cycles && tid == 0) {
            const long long now = clock64();
            atomicAdd(P.phase_cycles + idx, (unsigned long long)(now - t_prev));
            t_prev = now;
        }
    };

    // ---- load heightmap (column-parity planes, 8 KB) ----
    double* hm_g = P.hm + (int64_t)env * (HX * HY);
    {
        const double2* src = reinterpret_cast<const double2*>(hm_g);
        double2* dst = reinterpret_cast<double2*>(hm_s);
        if (mode == MODE_RESET) {
            for (int i = tid; i < HX * HY / 2; i += CTA_THREADS) dst[i] = make_double2(0.0, 0.0);
        } else {
            for (int i = tid; i < HX * HY / 2; i += CTA_THREADS) dst[i] = src[i];
        }
    }
    if (tid == 0) {
        misc->error = 0; misc->ovf_count = 0; misc->any_mask = 0;
        if (mode != MODE_ALL_OBS) P.r_error[env] = 0;
    }
    for (int i = tid; i < P.R * 8; i += CTA_THREADS) candbits_s[i] = 0u;
    __syncthreads();

    int32_t* queue_g = P.queue + (int64_t)env * MAX_QUEUE;
    bool emit_loc = true;        // location observation (scan + candidates) vs order observation
    bool write_state = true;

    // ---- phase A: bookkeeping / apply action ----
    if (mode == MODE_RESET) {
        if (tid == 0) {
            int cursor = P.cursor[env];
            const int nfill = P.K > 1 ? P.K : 1;
            for (int q = 0; q < nfill; ++q) queue_g[q] = draw_item(P, env, cursor);
            P.cursor[env] = cursor;
            P.packed[env] = 0; P.ep_len[env] = 0; P.vol_sum[env] = 0.0; P.ep_rew[env] = 0.0;
            P.order_act[env] = 0;
            misc->item = queue_g[0];
        }
        emit_loc = (P.K <= 1);
        __syncthreads();
    } else if (mode == MODE_STEP) {
        // decode the action (warp 0 computes the drop height of that single pose)
        __shared__ double z_sh;
        __shared__ int ok_sh, rot_sh, lx_sh, ly_sh, item_sh;
        if (warp == 0) {
            const int64_t a = P.actions[env];
            const int item = P.cur_item[env];
            int rot = 0, lx = 0, ly = 0;
            bool ok = true;
            if (a < 0 || a >= P.sel) { ok = false; if (lane == 0) misc->error = 2; }
            else {
                const uint16_t c = P.cand[(int64_t)env * P.sel + a];
                rot = c >> 8; lx = (c >> 4) & 15; ly = c & 15;
            }
            const ShapeRot* sr = P.srot + (int64_t)item * P.R + rot;
            // prejudge (binPhy.py:238-245)
            if (!((sr->okx >> lx) & 1u) || !((sr->oky >> ly) & 1u)) ok = false;
            if (!P.mask_any[env]) ok = false;
            double z = POSZ_INVALID;     // posZmap keeps 1e3 outside the scanned range (space.py:101)
            if (lx < sr->nX && ly < sr->nY) {
                const int w = sr->w, h = sr->h;
                const double* __restrict__ B = P.Bs + sr->off;
                double acc = sr->any_zero ? 0.0 : -INFINITY;
                for (int c = lane; c < w * h; c += 32) {
                    const int i = c / h, j = c - i * h;
                    const double v = hm_s[hm_index(STEP * lx + i, STEP * ly + j)] - B[c];
                    acc = (v > acc) ? v : acc;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double t = __shfl_xor_sync(0xffffffffu, acc, o);
                    acc = (t > acc) ? t : acc;
                }
                z = acc;
            }
            // Interface.simulateHeight (Interface.py:365-369): AABB top above the bin -> failure
            if (ok && !round6_le0(z + sr->ez - P.binz)) ok = false;
            if (lane == 0) { z_sh = z; ok_sh = ok; rot_sh = rot; lx_sh = lx; ly_sh = ly; item_sh = item; }
        }
        __syncthreads();
        const bool ok = ok_sh != 0;
        const int item = item_sh;
        if (ok) {
            // heightmap update: hm[win] = max(hm[win], (T + z) * maskT)   (space.py:213)
            const ShapeRot* sr = P.srot + (int64_t)item * P.R + rot_sh;
            const int w = sr->w, h = sr->h;
            const double* __restrict__ T = P.Ts + sr->off;
            const double z = z_sh;
            const int x0 = STEP * lx_sh, y0 = STEP * ly_sh;
            for (int c = tid; c < w * h; c += CTA_THREADS) {
                const int i = c / h, j = c - i * h;
                const double v = T[c] + z;
                double* cell = hm_s + hm_index(x0 + i, y0 + j);
                if (v > *cell) *cell = v;
            }
        } else {
            for (int i = tid; i < HX * HY; i += CTA_THREADS) hm_s[i] = 0.0;    // auto-reset
        }
        if (tid == 0) {
            int cursor = P.cursor[env];
            const int nfill = P.K > 1 ? P.K : 1;
            if (ok) {
                const double rew = P.reward_tab[item];
                P.r_reward[env] = (float)rew; P.r_done[env] = 0; P.r_valid[env] = 1;
                P.r_counter[env] = -1; P.r_eplen[env] = 0; P.r_ratio[env] = -1.0; P.r_eprew[env] = 0.0;
                P.packed[env] += 1; P.ep_len[env] += 1;
                P.vol_sum[env] += P.vol[item];
                P.ep_rew[env] += rew;
                // item_creator.update_item_queue(orderAction); generate_item()  (binPhy.py:324-325)
                const int oa = P.order_act[env];
                for (int q = oa; q + 1 < nfill; ++q) queue_g[q] = queue_g[q + 1];
                queue_g[nfill - 1] = draw_item(P, env, cursor);
            } else {
                P.r_reward[env] = 0.0f; P.r_done[env] = 1; P.r_valid[env] = 1;
                P.r_counter[env] = P.packed[env];
                P.r_ratio[env] = P.vol_sum[env] / P.binvol;
                P.r_eplen[env] = P.ep_len[env] + 1;
                P.r_eprew[env] = P.ep_rew[env] + 0.0;
                P.packed[env] = 0; P.ep_len[env] = 0; P.vol_sum[env] = 0.0; P.ep_rew[env] = 0.0;
                P.order_act[env] = 0;
                for (int q = 0; q < nfill; ++q) queue_g[q] = draw_item(P, env, cursor);   // reset(): clear + preview
            }
            P.cursor[env] = cursor;
            misc->item = queue_g[0];
        }
        emit_loc = (P.K <= 1);
        __syncthreads();
    } else if (mode == MODE_CANDIDATES) {
        if (tid == 0) {
            int64_t oa = P.actions[env];
            if (oa < 0 || oa >= P.K) { misc->error = 3; oa = 0; }
            P.order_act[env] = (int)oa;
            misc->item = queue_g[oa];
        }
        __syncthreads();
    } else if (mode == MODE_ALL_OBS) {
        if (tid == 0) misc->item = queue_g[slot];
        write_state = (slot == P.K - 1);
        __syncthreads();
    } else if (mode == MODE_DEBUG_SCAN) {
        if (tid == 0) misc->item = P.dbg_items[env];
        write_state = false;
        __syncthreads();
    } else {   // MODE_DEBUG_HULLS
        write_state = false;
        for (int i = tid; i < P.R * NPOSE; i += CTA_THREADS) {
            posz_s[i] = P.dbg_in_posz[(int64_t)env * P.R * NPOSE + i];
        }
        for (int wd = tid; wd < P.R * 8; wd += CTA_THREADS) {
            uint32_t bits = 0;
            for (int b = 0; b < 32; ++b)
                if (P.dbg_in_mask[(int64_t)env * P.R * NPOSE + wd * 32 + b] != 0.0) bits |= 1u << b;
            maskbits_s[wd] = bits;
        }
        if (tid == 0) misc->item = 0;
        __syncthreads();
    }

    phase_mark(0);   // load + phase A
    float* obs_g = P.obs + (int64_t)env * P.obs_stride + (mode == MODE_ALL_OBS ? slot * P.loc_len : 0);
    const int item = misc->item;
    const int ncand = P.sel * 5;

    // the heightmap is final for this call: emit its float32 copy and write the state back now, so
    // region X can be recycled after the scan
    {
        const int hm_obs_off = emit_loc ? ncand + 9 : P.K;
        for (int i = tid; i < HX * HY; i += CTA_THREADS)
            obs_g[hm_obs_off + i] = (float)hm_s[hm_index(i >> 5, i & 31)];
        if (mode == MODE_STEP || mode == MODE_RESET) {
            double2* dst = reinterpret_cast<double2*>(hm_g);
            const double2* src = reinterpret_cast<const double2*>(hm_s);
            for (int i = tid; i < HX * HY / 2; i += CTA_THREADS) dst[i] = src[i];
        }
    }
    if (!emit_loc) {
        // order observation: [next k item ids | heightmap]  (binPhy.py:229-230)
        for (int i = tid; i < P.K; i += CTA_THREADS) obs_g[i] = (float)queue_g[i];
        if (tid == 0 && misc->error) P.r_error[env] = (uint8_t)misc->error;
        return;
    }

    // ---- phase B: drop height + feasibility of every pose, one warp per rotation ----
    if (mode != MODE_DEBUG_HULLS) {
        for (int r = warp; r < P.R; r += CTA_WARPS) scan_rotation(P, hm_s, posz_s, maskbits_s, item, r, lane);
    }
    __syncthreads();                       // heightmap dead from here on: region X becomes contour scratch
    phase_mark(1);   // heightmap write-back + scan

    // ---- phase C: candidate extraction, rotations in groups of CTA_WARPS ----
    const int ngroups = (P.R + CTA_WARPS - 1) / CTA_WARPS;
    for (int g = 0; g < ngroups; ++g) {
        const int r = g * CTA_WARPS + warp;
        int lv[8];
        uint64_t present = 0;
        if (r < P.R) {
            int err = 0;
            levels_from_maps(P, posz_s, maskbits_s, r, lane, lv, present, err);
            if (__any_sync(0xffffffffu, err) && lane == 0) misc->error = 4;
            present &= ~(1ull << (LEVEL_OFFSET - 1));        // level -1 is skipped (cvTools.py:84)
        }
        for (;;) {                                           // rounds of at most QUOTA levels per rotation
            int nl = 0;
            if (r < P.R) nl = build_level_bitmaps(slots_s, warp, lane, lv, present);
            if (lane == 0) misc->nlev[warp] = nl;
            __syncthreads();
            int pre[CTA_WARPS + 1];
            pre[0] = 0;
#pragma unroll
            for (int q = 0; q < CTA_WARPS; ++q) pre[q + 1] = pre[q] + misc->nlev[q];
            const int ntask = pre[CTA_WARPS];
            if (ntask == 0) break;                           // uniform: every warp sees the same counts
            // one lane per (rotation, level) image; the first TRACE_WARPS warps take the tasks, all their
            // lanes enter the lock-step routine together
            if (warp < TRACE_WARPS) {
                const int li = lane / TRACE_WARPS_DIV;                 // (unused when TRACE_LANES == 32)
                (void)li;
                const int t = lane * TRACE_WARPS + warp;
                const bool has = (lane < TRACE_LANES) && (t < ntask);
                int wq = 0;
#pragma unroll
                for (int q = 1; q < CTA_WARPS; ++q) if (has && t >= pre[q]) wq = q;
                const int sl = has ? wq * QUOTA + (t - pre[wq]) : 0;
                const int sidx = warp * TRACE_LANES + (lane < TRACE_LANES ? lane : 0);
                StridedScratch<NSLOT, FAST_CAP> sc;
                sc.w = reinterpret_cast<uint32_t*>(scratch_s) + sidx;
                sc.b = scratch_s + NSLOT * 16 * 4 + sidx;
                sc.kept = 0;
                const uint32_t* bm = slots_s + sl * SLOT_WORDS;
                uint32_t* cb = candbits_s + (g * CTA_WARPS + wq) * 8;
                const bool okc = process_level_image_lockstep(
                    sc, bm, has, P.legacy != 0,
                    [&](int x, int y) { const int b = x * 16 + y; atomicOr(cb + (b >> 5), 1u << (b & 31)); });
                if (!okc) {
                    const int k = atomicAdd(&misc->ovf_count, 1);
                    if (k < 32) misc->ovf_task[k] = (wq << 16) | sl;
                }
            }
            __syncthreads();
            if (misc->ovf_count > 0) {          // rare: contours longer than FAST_CAP points, serial path
                if (tid == 0) {
                    FlatScratch<BIG_CAP> bs;
                    bs.w = reinterpret_cast<uint32_t*>(big_s);
                    bs.b = big_s + 16 * 4;
                    const int n_ovf = misc->ovf_count;
                    if (n_ovf > 32) misc->error = 5;
                    for (int k = 0; k < (n_ovf < 32 ? n_ovf : 32); ++k) {
                        const int wq = misc->ovf_task[k] >> 16, sl = misc->ovf_task[k] & 0xFFFF;
                        const uint32_t* bm = slots_s + sl * SLOT_WORDS;
                        uint32_t* cb = candbits_s + (g * CTA_WARPS + wq) * 8;
                        const bool okc = process_level_image(
                            bs, bm, P.legacy != 0,
                            [&](int x, int y) { const int b = x * 16 + y; cb[b >> 5] |= 1u << (b & 31); });
                        if (!okc) misc->error = 6;
                    }
                    misc->ovf_count = 0;
                }
                __syncthreads();
            }
        }
        __syncthreads();     // nobody may still be reading misc->nlev when the next group rewrites it
    }

    phase_mark(2);   // levels, bitmaps, contour tasks
    // ---- phase D: select / pad, observation assembly ----
    if (tid < 32) {
        int c = 0;
        if (tid < P.R) { for (int q = 0; q < 8; ++q) c += __popc(candbits_s[tid * 8 + q]); }
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        misc->cnt_prefix[tid + 1] = incl;
        if (tid == 0) misc->cnt_prefix[0] = 0;
        uint32_t any = 0;
        for (int q = tid; q < P.R * 8; q += 32) any |= maskbits_s[q];
        any = __reduce_or_sync(0xffffffffu, any);
        if (tid == 0) misc->any_mask = any != 0;
    }
    __syncthreads();
    const int Ktot = misc->cnt_prefix[P.R];
    const int sel = P.sel;
    double* dbg_cand = P.dbg_cand ? P.dbg_cand + (int64_t)env * sel * 5 : nullptr;
    auto put_row = [&](int row, int rot, int x, int y, double H, double V) {
        float* d = stage_f + row * 5;
        d[0] = (float)rot; d[1] = (float)x; d[2] = (float)y; d[3] = (float)H; d[4] = (float)V;
        stage_c[row] = (uint16_t)((rot << 8) | (x << 4) | y);
        if (dbg_cand) {
            double* q = dbg_cand + row * 5;
            q[0] = rot; q[1] = x; q[2] = y; q[3] = H; q[4] = V;
        }
    };
    auto zero_row = [&](int row) {
        float* d = stage_f + row * 5;
        d[0] = d[1] = d[2] = d[3] = d[4] = 0.0f;
        stage_c[row] = 0;
        if (dbg_cand) { double* q = dbg_cand + row * 5; q[0] = q[1] = q[2] = q[3] = q[4] = 0.0; }
    };
    // (the staging area is region X: heightmap and contour scratch are dead from here on)
    if (Ktot == 0) {
        // no hull candidate at all (binPhy.py:217-225): the `sel` smallest posZValid, stable order
        const int total = P.R * NPOSE;
        if (!misc->any_mask) {
            for (int i = tid; i < sel; i += CTA_THREADS) {
                if (i < total) put_row(i, i >> 8, (i >> 4) & 15, i & 15, P.binz, 0.0);
                else zero_row(i);   // reference would produce a short table; R*256 >= sel is enforced at create
            }
        } else {
            for (int i = tid; i < total; i += CTA_THREADS) {
                const bool mi = (maskbits_s[i >> 5] >> (i & 31)) & 1u;
                const double vi = mi ? posz_s[i] : POSZ_INVALID;
                int rank = 0;
                for (int j = 0; j < total; ++j) {
                    const bool mj = (maskbits_s[j >> 5] >> (j & 31)) & 1u;
                    const double vj = mj ? posz_s[j] : POSZ_INVALID;
                    rank += (vj < vi) || (vj == vi && j < i);
                }
                if (rank < sel) put_row(rank, i >> 8, (i >> 4) & 15, i & 15, P.binz, mi ? 1.0 : 0.0);
            }
            for (int i = total + tid; i < sel; i += CTA_THREADS) zero_row(i);
        }
    } else {
        // rows in rotation order, then (col, row) ascending == bit order of the per-rotation sets
        for (int idx = tid; idx < P.R * NPOSE; idx += CTA_THREADS) {
            const int r = idx >> 8, b = idx & 255;
            const uint32_t* cb = candbits_s + r * 8;
            if (!((cb[b >> 5] >> (b & 31)) & 1u)) continue;
            int ord = misc->cnt_prefix[r];
            for (int q = 0; q < (b >> 5); ++q) ord += __popc(cb[q]);
            ord += __popc(cb[b >> 5] & ((1u << (b & 31)) - 1u));
            const int col = b >> 4, row = b & 15;
            const int cell = r * NPOSE + row * 16 + col;
            const bool m = (maskbits_s[cell >> 5] >> (cell & 31)) & 1u;
            const double H = m ? posz_s[cell] : POSZ_INVALID;
            int dest = ord;
            if (Ktot > sel) {
                // truncate to the `sel` lowest heights, ties by original order (stable argsort; binPhy.py:209-212)
                int rank = 0;
                for (int r2 = 0; r2 < P.R; ++r2) {
                    const uint32_t* cb2 = candbits_s + r2 * 8;
                    int ord2 = misc->cnt_prefix[r2];
                    for (int q = 0; q < 8; ++q) {
                        uint32_t wbits = cb2[q];
                        while (wbits) {
                            const int bb = q * 32 + __ffs((int)wbits) - 1;
                            wbits &= wbits - 1;
                            const int cell2 = r2 * NPOSE + (bb & 15) * 16 + (bb >> 4);
                            const bool m2 = (maskbits_s[cell2 >> 5] >> (cell2 & 31)) & 1u;
                            const double H2 = m2 ? posz_s[cell2] : POSZ_INVALID;
                            rank += (H2 < H) || (H2 == H && ord2 < ord);
                            ++ord2;
                        }
                    }
                }
                dest = rank;
            }
            if (dest < sel) put_row(dest, r, row, col, H, m ? 1.0 : 0.0);
        }
        for (int i = Ktot + tid; i < sel; i += CTA_THREADS) zero_row(i);
    }
    __syncthreads();

    phase_mark(3);   // select / pad into the staging area
    // observation: [candidates sel*5 | next_item_vec 9 | heightmap]  (binPhy.py:196-227)
    for (int i = tid; i < ncand; i += CTA_THREADS) obs_g[i] = stage_f[i];
    if (tid < 9) obs_g[ncand + tid] = (tid == 0) ? (float)item : 0.0f;

    if (write_state) {
        uint16_t* cg = P.cand + (int64_t)env * sel;
        for (int i = tid; i < sel; i += CTA_THREADS) cg[i] = stage_c[i];
        if (tid == 0) { P.cur_item[env] = item; P.mask_any[env] = (uint8_t)misc->any_mask; }
    }
    if (tid == 0 && misc->error) P.r_error[env] = (uint8_t)misc->error;

    // float64 parity views
    if (P.dbg_posz) {
        const int64_t base = (int64_t)env * P.R * NPOSE;
        for (int i = tid; i < P.R * NPOSE; i += CTA_THREADS) {
            const bool m = (maskbits_s[i >> 5] >> (i & 31)) & 1u;
            P.dbg_posz[base + i] = posz_s[i];
            P.dbg_poszv[base + i] = m ? posz_s[i] : POSZ_INVALID;
            P.dbg_mask[base + i] = m ? 1.0 : 0.0;
        }
    }
    if (P.dbg_nhull && tid == 0) P.dbg_nhull[env] = Ktot;
    phase_mark(4);   // observation / state stores
}

}  // namespace irbpp
