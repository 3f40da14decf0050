// irbpp_heuristic.cuh -- placement heuristics of the reference's Space over the drop-height / feasibility
// grid of the last scan (reference environment/physics0/space.py:162-227, `get_heuristic_action`):
// MINZ, DBLF, FIRSTFIT and HM scores, np.round(score, 6), np.argmin (first index on ties, flat
// (rotation, lx, ly) order).  The reference's RANDOM branch raises on every call (np.random.choice of a
// tuple) and is not reproduced.  Host-compilable so the CPU test harness can check the arithmetic.
#pragma once
#include "irbpp_math.cuh"

namespace irbpp {

enum Heuristic : int { HEUR_MINZ = 0, HEUR_DBLF = 1, HEUR_FIRSTFIT = 2, HEUR_HM = 3, HEUR_COUNT = 4 };
constexpr double HEUR_INVALID = 1e6;     // score of infeasible poses (space.py:171,182,192,201)

// Score of ONE feasible pose (rounded).  `hm(x, y)` reads the heightmap; `Ts` is the top table of the
// item's rotation with -inf where maskT == 0 (max(hm, -inf) == max(hm, (T + z) * 0) since hm >= 0);
// w x h its window; z the pose's drop height.
template <class HM>
__host__ __device__ inline double heuristic_score(int method, int dir_idx, int lx, int ly, double z, double resA,
                                                  int ax, int ay, int step, const HM& hm, const double* Ts,
                                                  int w, int h) {
    const bool xflip = dir_idx >= 2, yflip = (dir_idx & 1) != 0;          // space.py:163-166
    const double cx = xflip ? (double)(ax - lx) : (double)lx;
    const double cy = yflip ? (double)(ay - ly) : (double)ly;
    double s;
    if (method == HEUR_MINZ) s = z;                                          // space.py:170
    else if (method == HEUR_DBLF) s = (cx + cy) * resA + 100.0 * z;       // space.py:178-181
    else if (method == HEUR_FIRSTFIT) s = cx + cy;                           // space.py:189-191
    else {                                                                   // HM, space.py:199-219
        auto elem = [&](int c) {
            const int i = c / h, j = c - i * h;
            const double t = Ts[c] + z, x = hm(step * lx + i, step * ly + j);
            return (t > x) ? t : x;                                          // heightmapC_Prime
        };
        const double map_sum = np_pairwise_sum(elem, w * h);
        s = (cx + cy) * resA;
        s += map_sum * 100.0;
    }
    return np_round6(s);
}

}  // namespace irbpp
