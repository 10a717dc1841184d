// Static block / assembly plan (host side, no CUDA).
//
// The reference re-derives the block layout every step (SMC:461-479) and walks the blocks
// sequentially during re-assembly (SMC:221-348, GRAD:282-356).  Everything in those loops
// except the predicted values depends only on the grid size and the static flow mask, so it
// is compiled once per mesh into:
//   * extraction origins + (idx_i, idx_j) tags, in the reference's order,
//   * the placement sub-rectangle of every block and the last-writer ("owner") map,
//   * a de-duplicated list of masked rectangle means ("tasks"),
//   * per field and block the recurrence  c_k = m[a] - (m[b] - c[parent])  (or  m[a] - Ref_BC),
//     with every NaN branch of the reference (np.isnan(BC_ups[j]), SMC:252) resolved
//     statically: a masked mean is NaN iff its mask rectangle is empty.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace psm {

struct Task {
    int32_t src, msk, ch;        // values from block `src`, channel `ch`; mask of block `msk` (-1: no mask, plain mean)
    int32_t y0, y1, x0, x1;      // block-local rectangle [y0,y1) x [x0,x1)
    int32_t count;               // mask pixels in the rectangle (0 -> mean is NaN)
};

// One maximal run of a shift line (SMC:350 / GRAD:358-361) owned by a single block: the plain sum of the
// UNCORRECTED block values over `rect` enters the global shift as  coef * (sum - n * c[block]).
struct LineTask {
    int32_t src, ch;             // block, channel (= field)
    int32_t y0, y1, x0, x1;      // block-local rectangle (one pixel wide or high)
    int32_t coef, n;             // +3 (line a) or -1 (line b); number of pixels
};

struct Rec {
    int32_t ta, tb, parent;      // c_k = m[ta] - (tb >= 0 ? m[tb] - c[parent] : ref_bc)
    int32_t is_nan;              // statically known NaN
};

struct Plan {
    int variant = 0, H = 0, W = 0, S = 0, ov = 0, stride = 0;
    int n_x = 0, n_y = 0, p_i = 0, p_j = 0, B = 0, F = 0, C = 0;
    int ncolb = 0;                                   // blocks per block row (n_x + 1; n_x + 2 for the thesis plan)
    std::vector<int32_t> y0, x0, idx_i, idx_j;       // [B]
    std::vector<int32_t> py0, py1, px0, px1;         // [B] placement rectangle, block-local
    std::vector<int32_t> owner;                      // [H*W] last writer, -1 if none
    std::vector<Task> tasks;
    std::vector<Rec> rec;                            // [F*B]
    int shift_axis[2] = {0, 0};                      // 0: two columns (mean over rows); 1: two rows
    int shift_a[2] = {0, 0}, shift_b[2] = {0, 0};    // result -= mean(3*line_a - line_b)/3
    std::vector<LineTask> lines[2];                  // per field: runs of the two shift lines through the owner map
    int shift_len[2] = {0, 0};                       // pixels per shift line (H or W)
    int max_depth = 0;                               // longest parent chain
    std::string error;
};

// Returns 0, or a negative psm_status_code with plan.error set.
int compile_plan(int variant, int H, int W, int S, int ov, const uint8_t* mask, Plan& plan);

}  // namespace psm
