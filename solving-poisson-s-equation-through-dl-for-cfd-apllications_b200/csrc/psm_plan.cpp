// Plan compiler: see psm_plan.h.  Reference semantics followed:
//   deltaU_to_deltaP : SMC:461-479 (plan), SMC:203-348 (corrections + placement), SMC:350 (shift)
//   U_to_gradP       : GRAD:479-500 (plan), GRAD:269-356, GRAD:358-361
//   thesis U -> p    : PMP:303-332 (plan with the extra -1 column), PMP:372-467 (corrections + placement), PMP:472
#include "psm_plan.h"

#include <cmath>
#include <map>
#include <tuple>

#include "../../include/psm_b200.h"

namespace psm {
namespace {

struct Builder {
    Plan& P;
    const uint8_t* mask;
    std::vector<int32_t> psum;   // (H+1)*(W+1) summed-area table of the mask
    std::map<std::tuple<int, int, int, int, int, int, int>, int> dedupe;

    Builder(Plan& p, const uint8_t* m) : P(p), mask(m) {
        const int H = P.H, W = P.W;
        psum.assign((size_t)(H + 1) * (W + 1), 0);
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x)
                psum[(size_t)(y + 1) * (W + 1) + x + 1] = (mask[(size_t)y * W + x] ? 1 : 0) +
                    psum[(size_t)y * (W + 1) + x + 1] + psum[(size_t)(y + 1) * (W + 1) + x] -
                    psum[(size_t)y * (W + 1) + x];
    }
    int count(int blk, int y0, int y1, int x0, int x1) const {
        const int W1 = P.W + 1;
        const int Y0 = P.y0[blk] + y0, Y1 = P.y0[blk] + y1, X0 = P.x0[blk] + x0, X1 = P.x0[blk] + x1;
        return psum[(size_t)Y1 * W1 + X1] - psum[(size_t)Y0 * W1 + X1] - psum[(size_t)Y1 * W1 + X0] +
               psum[(size_t)Y0 * W1 + X0];
    }
    int task(int src, int msk, int ch, int y0, int y1, int x0, int x1) {
        auto key = std::make_tuple(src, msk, ch, y0, y1, x0, x1);
        auto it = dedupe.find(key);
        if (it != dedupe.end()) return it->second;
        Task t{src, msk, ch, y0, y1, x0, x1, msk >= 0 ? count(msk, y0, y1, x0, x1) : (y1 - y0) * (x1 - x0)};
        P.tasks.push_back(t);
        int id = (int)P.tasks.size() - 1;
        dedupe.emplace(key, id);
        return id;
    }
    bool tnan(int t) const { return P.tasks[t].count == 0; }
};

// Per-field walking state: BC_ups[j] of the reference as (task, block) pairs.
struct Ups {
    std::vector<int> task, block;
    std::vector<char> nan;
    explicit Ups(int n) : task(n, -1), block(n, -1), nan(n, 0) {}
};

}  // namespace

// Runs of the two shift lines through the owner map (shared by all variants).
static void build_shift_lines(Plan& P) {
    const int H = P.H, W = P.W;
    for (int f = 0; f < P.F; ++f) {
        const int axis = P.shift_axis[f];
        const int len = axis == 0 ? H : W;
        P.shift_len[f] = len;
        for (int s2 = 0; s2 < 2; ++s2) {
            const int line = s2 == 0 ? P.shift_a[f] : P.shift_b[f];
            const int coef = s2 == 0 ? 3 : -1;
            int i = 0;
            while (i < len) {
                const int o = P.owner[axis == 0 ? (size_t)i * W + line : (size_t)line * W + i];
                int e = i + 1;
                while (e < len && P.owner[axis == 0 ? (size_t)e * W + line : (size_t)line * W + e] == o) ++e;
                LineTask t{};
                t.src = o; t.ch = f; t.coef = coef; t.n = e - i;
                if (axis == 0) { t.y0 = i - P.y0[o]; t.y1 = e - P.y0[o]; t.x0 = line - P.x0[o]; t.x1 = t.x0 + 1; }
                else           { t.x0 = i - P.x0[o]; t.x1 = e - P.x0[o]; t.y0 = line - P.y0[o]; t.y1 = t.y0 + 1; }
                P.lines[f].push_back(t);
                i = e;
            }
        }
    }
}

// Thesis solver module (PMP:303-332, 372-472): stride S - avance, blocks right -> left and, after the last regular
// block of every row, the block at x = 0 tagged -1; its own correction chain (see oracle/assemble.py:assemble_thesis).
static int compile_thesis_plan(int H, int W, int S, int av, const uint8_t* mask, Plan& P) {
    const int st = P.stride;
    P.C = P.F = 1;
    P.n_x = (W - S) / st;                                      // PMP:306  int()
    P.n_y = (H - S) / st;                                      // PMP:307
    const int n_x = P.n_x, n_y = P.n_y;
    P.p_i = H - (st * n_y + S);                                // `p`, PMP:407
    P.p_j = (W - S) - n_x * st;                                // PMP:391
    const int p = P.p_i, p_j = P.p_j;
    if (S - p - av < 0) { P.error = "thesis plan: bottom strip outside the block"; return PSM_ERR_GEOMETRY; }
    P.ncolb = n_x + 2;
    P.B = (n_y + 2) * P.ncolb;
    if (P.B >= 65535) { P.error = "too many blocks for the 16-bit owner map"; return PSM_ERR_GEOMETRY; }
    const int B = P.B;
    for (int i = 0; i < n_y + 2; ++i) {
        const int y_0 = (i == n_y + 1) ? H - S : i * st;
        for (int j = 0; j <= n_x; ++j) {
            P.idx_i.push_back(i); P.idx_j.push_back(n_x - j); P.y0.push_back(y_0); P.x0.push_back((W - S) - j * st);
        }
        P.idx_i.push_back(i); P.idx_j.push_back(-1); P.y0.push_back(y_0); P.x0.push_back(0);       // PMP:323-329
    }
    // placement (PMP:453-467)
    P.py0.assign(B, 0); P.py1.assign(B, S); P.px0.assign(B, 0); P.px1.assign(B, S);
    for (int k = 0; k < B; ++k) {
        if (P.idx_i[k] == n_y + 1) {
            P.py0[k] = av;                                                         // res[avance:shape]
            if (P.idx_j[k] == -1) P.px1[k] = W - (n_x + 1) * st - av;              // PMP:454
        }
    }
    P.owner.assign((size_t)H * W, -1);
    for (int k = 0; k < B; ++k)
        for (int ly = P.py0[k]; ly < P.py1[k]; ++ly) {
            int32_t* row = &P.owner[(size_t)(P.y0[k] + ly) * W + P.x0[k]];
            for (int lx = P.px0[k]; lx < P.px1[k]; ++lx) row[lx] = k;
        }
    for (size_t q = 0; q < P.owner.size(); ++q)
        if (P.owner[q] < 0) { P.error = "placement does not cover the grid"; return PSM_ERR_GEOMETRY; }

    Builder bld(P, mask);
    P.rec.assign((size_t)B, Rec{-1, -1, -1, 0});
    std::vector<int> depth((size_t)B, 0);
    Rec* rec = P.rec.data();
    auto set = [&](int k, int ta, int tb, int parent) {
        Rec r{ta, tb, parent, 0};
        r.is_nan = bld.tnan(ta) || (tb >= 0 && (bld.tnan(tb) || rec[parent].is_nan));
        rec[k] = r;
        depth[k] = (parent >= 0) ? depth[parent] + 1 : 0;
        if (depth[k] > P.max_depth) P.max_depth = depth[k];
    };
    Ups ups(n_x + 1);
    int up_task = -1, up_block = -1; bool up_nan = false;      // BC_up_ : the chain of the -1 column
    auto set_ups = [&](int k, int j, int y0, int y1, int x0, int x1) {
        ups.task[j] = bld.task(k, k, 0, y0, y1, x0, x1); ups.block[j] = k;
        ups.nan[j] = bld.tnan(ups.task[j]) || rec[k].is_nan;
    };
    auto left_of_prev = [&](int k) { return bld.task(k - 1, k - 1, 0, 0, S, 0, av); };   // BC_ant_0 / BC_alter: previous block's own left strip
    for (int k = 0; k < B; ++k) {
        const int ii = P.idx_i[k], jj = P.idx_j[k];
        if (ii == 0) {
            if (jj == n_x) {                                                       // PMP:387-391
                set(k, bld.task(k, k, 0, 0, S, S - av, S), -1, -1);                // - BC_up (= 0)
                set_ups(k, jj, S - av, S, S - av, S);
            } else if (jj == -1) {                                                 // PMP:394-398
                set(k, bld.task(k, k, 0, 0, S, p_j, p_j + av), left_of_prev(k), k - 1);
                up_task = bld.task(k, k, 0, S - av, S, p_j, p_j + av); up_block = k;
                up_nan = bld.tnan(up_task) || rec[k].is_nan;
            } else {                                                               // PMP:399-402
                set(k, bld.task(k, k, 0, 0, S, S - av, S), left_of_prev(k), k - 1);
                set_ups(k, jj, S - av, S, 0, S);
            }
        } else if (ii == n_y + 1) {
            if (jj == -1) set(k, bld.task(k, k, 0, S - p - av, S - p, p_j, p_j + av), up_task, up_block);   // PMP:408-410
            else if (ups.nan[jj]) set(k, bld.task(k, k, 0, 0, S, S - av, S), left_of_prev(k), k - 1);      // PMP:415-416
            else set(k, bld.task(k, k, 0, S - p - av, S - p, 0, S), ups.task[jj], ups.block[jj]);          // PMP:418
        } else {
            if (jj == -1) {                                                        // PMP:432-437
                set(k, bld.task(k, k, 0, 0, av, p_j, p_j + av), up_task, up_block);
                up_task = bld.task(k, -1, 0, S - av, S, p_j, p_j + av); up_block = k;   // np.mean WITHOUT the mask
                up_nan = rec[k].is_nan;
            } else {
                if (ups.nan[jj]) set(k, bld.task(k, k, 0, 0, S, S - av, S), left_of_prev(k), k - 1);       // PMP:441-442
                else set(k, bld.task(k, k, 0, 0, av, 0, S), ups.task[jj], ups.block[jj]);                  // PMP:444
                set_ups(k, jj, S - av, S, 0, S);                                                          // PMP:446
            }
        }
    }
    (void)up_nan;
    P.shift_axis[0] = 0; P.shift_a[0] = W - 1; P.shift_b[0] = W - 2;              // PMP:472
    build_shift_lines(P);
    return 0;
}

int compile_plan(int variant, int H, int W, int S, int ov, const uint8_t* mask, Plan& P) {
    P = Plan();
    P.variant = variant; P.H = H; P.W = W; P.S = S; P.ov = ov; P.stride = S - ov;
    if (variant != PSM_DELTAU_TO_DELTAP && variant != PSM_U_TO_GRADP && variant != PSM_THESIS_U_TO_P) { P.error = "unknown variant"; return PSM_ERR_INVALID; }
    if (S != 128) { P.error = "only shape == 128 is supported (the reference hard-codes 128**2 at SMC:307)"; return PSM_ERR_INVALID; }
    if (ov <= 0 || ov >= S) { P.error = "overlap must be in (0, shape)"; return PSM_ERR_INVALID; }
    if (H < S || W <= S) { P.error = "grid smaller than one block"; return PSM_ERR_GEOMETRY; }
    if (!mask) { P.error = "mask is NULL"; return PSM_ERR_INVALID; }
    const int st = P.stride;
    if (variant == PSM_THESIS_U_TO_P) return compile_thesis_plan(H, W, S, ov, mask, P);
    const bool grad = (variant == PSM_U_TO_GRADP);
    P.C = P.F = grad ? 2 : 1;
    P.n_x = (int)std::ceil((double)(W - S) / (double)st);      // SMC:461 / GRAD:479
    P.n_y = (H - S) / st;                                      // SMC:462 / GRAD:480
    P.p_i = H - (st * P.n_y + S);                              // SMC:213 / GRAD:277
    P.p_j = W - (st * P.n_x + S);                              // SMC:216 / GRAD:278
    const int n_x = P.n_x, n_y = P.n_y, p_i = P.p_i, p_j = P.p_j;
    if (p_i == 0) {
        P.error = "(H - shape) is a multiple of the stride: the reference is undefined there "
                  "(empty slice at SMC:292, shape mismatch at SMC:335)";
        return PSM_ERR_GEOMETRY;
    }
    if (n_x < 1) { P.error = "n_x == 0: the reference reads an unbound block (SMC:239)"; return PSM_ERR_GEOMETRY; }
    P.B = (n_y + 2) * (n_x + 1);
    P.ncolb = n_x + 1;
    if (P.B >= 65535) { P.error = "too many blocks for the 16-bit owner map"; return PSM_ERR_GEOMETRY; }
    const int B = P.B;

    // ---- extraction plan -----------------------------------------------------------------
    for (int i = 0; i < n_y + 2; ++i)
        for (int j = 0; j < n_x + 1; ++j) {
            int x_0, y_0 = i * st;
            if (i == n_y + 1) y_0 = H - S;
            if (!grad) {                                       // right -> left, SMC:468-469
                x_0 = W - j * st - S;
                if (j == n_x) x_0 = 0;
                P.idx_j.push_back(n_x - j);
            } else {                                           // left -> right, GRAD:489-490
                x_0 = j * st;
                if (j == n_x) x_0 = W - S;
                P.idx_j.push_back(j);
            }
            P.idx_i.push_back(i);
            P.y0.push_back(y_0);
            P.x0.push_back(x_0);
        }

    // ---- placement rectangles + owner map --------------------------------------------------
    const int wide = ov - p_j;                                 // intersect_zone_limit, SMC:238 / GRAD:308
    P.py0.assign(B, 0); P.py1.assign(B, S); P.px0.assign(B, 0); P.px1.assign(B, S);
    for (int k = 0; k < B; ++k) {
        const int ii = P.idx_i[k], jj = P.idx_j[k];
        if (!grad) {
            if (ii == n_y + 1) P.py0[k] = S - p_i;             // SMC:334-335,339-342: bottom p_i rows only
        } else {
            if (ii == n_y + 1) P.py0[k] = ov;                  // GRAD:346,353: rows [avance, shape)
            if (jj == n_x) P.px0[k] = S - wide;                // GRAD:346,350: last `wide` columns
        }
    }
    P.owner.assign((size_t)H * W, -1);
    for (int k = 0; k < B; ++k)
        for (int ly = P.py0[k]; ly < P.py1[k]; ++ly) {
            int32_t* row = &P.owner[(size_t)(P.y0[k] + ly) * W + P.x0[k]];
            for (int lx = P.px0[k]; lx < P.px1[k]; ++lx) row[lx] = k;
        }
    for (size_t q = 0; q < P.owner.size(); ++q)
        if (P.owner[q] < 0) { P.error = "placement does not cover the grid"; return PSM_ERR_GEOMETRY; }

    // ---- correction recurrence ---------------------------------------------------------------
    Builder bld(P, mask);
    P.rec.assign((size_t)P.F * B, Rec{-1, -1, -1, 0});
    std::vector<int> depth((size_t)P.F * B, 0);
    for (int f = 0; f < P.F; ++f) {
        Ups ups(n_x + 1);
        Rec* rec = &P.rec[(size_t)f * B];
        auto set = [&](int k, int ta, int tb, int parent) {
            Rec r{ta, tb, parent, 0};
            r.is_nan = bld.tnan(ta) || (tb >= 0 && (bld.tnan(tb) || rec[parent].is_nan));
            rec[k] = r;
            depth[(size_t)f * B + k] = (parent >= 0) ? depth[(size_t)f * B + parent] + 1 : 0;
            if (depth[(size_t)f * B + k] > P.max_depth) P.max_depth = depth[(size_t)f * B + k];
        };
        // side-strip correction against the previous block in order (SMC:235-236 / GRAD:305-306):
        // the previous block's strip is averaged under THIS block's mask.
        auto side = [&](int k, int width) {
            if (!grad) set(k, bld.task(k, k, f, 0, S, S - width, S), bld.task(k - 1, k, f, 0, S, 0, width), k - 1);
            else       set(k, bld.task(k, k, f, 0, S, 0, width), bld.task(k - 1, k, f, 0, S, S - width, S), k - 1);
        };
        auto set_ups = [&](int k, int j, int y0, int y1) {
            ups.task[j] = bld.task(k, k, f, y0, y1, 0, S);
            ups.block[j] = k;
            ups.nan[j] = bld.tnan(ups.task[j]) || rec[k].is_nan;
        };
        const int edge_j = grad ? n_x : 0;                    // column where the side strip is `wide`
        for (int k = 0; k < B; ++k) {
            const int ii = P.idx_i[k], jj = P.idx_j[k];
            if (ii == 0) {                                     // first row
                if (k == 0) {
                    int ta;
                    if (!grad) ta = bld.task(0, 0, f, 0, S, S - 1, S);             // SMC:233 outlet column
                    else if (f == 0) {                                            // GRAD:294-300
                        int col = 0;
                        while (col < S && bld.count(0, 0, S, col, col + 1) == 0) ++col;
                        if (col >= S) { P.error = "first block has no flow pixel (GRAD:297 assert)"; return PSM_ERR_GEOMETRY; }
                        ta = bld.task(0, 0, f, 0, S, col, col + 1);
                    } else ta = bld.task(0, 0, f, 1, 2, 0, S);                     // GRAD:303 row 1
                    set(k, ta, -1, -1);
                } else if (jj == edge_j) side(k, wide);                           // SMC:237-240 / GRAD:307-310 (overrides)
                else side(k, ov);
                set_ups(k, jj, S - ov, S);                                        // SMC:246
            } else if (ii != n_y + 1) {                        // middle rows
                if (ups.nan[jj]) {                                                // SMC:252-263 / GRAD:315-322
                    if (jj == edge_j) side(k, wide);
                    else if (!grad && jj == n_x) set(k, bld.task(k, k, f, 0, ov, 0, S), ups.task[jj], ups.block[jj]);
                    else side(k, ov);
                } else set(k, bld.task(k, k, f, 0, ov, 0, S), ups.task[jj], ups.block[jj]);   // SMC:265
                if (ii == n_y) set_ups(k, jj, p_i, S);                            // SMC:282-283
                else set_ups(k, jj, S - ov, S);                                   // SMC:279
            } else {                                           // last row
                if (!grad) {
                    if (jj == n_x) set(k, bld.task(k, k, f, S - p_i - ov, S - p_i, 0, S), ups.task[jj], ups.block[jj]);  // SMC:292
                    else {
                        const int n_up = bld.count(k, S - p_i - ov, S - p_i, 0, S);
                        if ((double)n_up / (128.0 * 128.0) > 0.9) side(k, jj == 0 ? wide : ov);   // SMC:307-314
                        else set(k, bld.task(k, k, f, 0, S - p_i, 0, S), ups.task[jj], ups.block[jj]);  // SMC:316
                    }
                } else {
                    if (ups.nan[jj]) side(k, jj == n_x ? wide : ov);               // GRAD:331-338
                    else set(k, bld.task(k, k, f, S - p_i - ov, S - p_i, 0, S), ups.task[jj], ups.block[jj]);  // GRAD:340
                }
            }
        }
    }

    // ---- global shift ---------------------------------------------------------------------------
    if (!grad) { P.shift_axis[0] = 0; P.shift_a[0] = W - 1; P.shift_b[0] = W - 2; }         // SMC:350
    else {
        P.shift_axis[0] = 0; P.shift_a[0] = 0; P.shift_b[0] = 1;                            // GRAD:359
        P.shift_axis[1] = 1; P.shift_a[1] = 1; P.shift_b[1] = 2;                            // GRAD:361
    }
    // result -= mean(3*line_a - line_b)/3 through the owner map, as per-block runs: the placed value is
    // raw - c[owner], so every run contributes coef * (sum(raw) - n * c[owner]).
    build_shift_lines(P);
    return 0;
}

}  // namespace psm

// ------------------------------------------------------------------------------------------------
// C-ABI: host-only entry points (usable without a GPU).
extern "C" int psm_plan_sizes(int32_t variant, int32_t grid_h, int32_t grid_w, int32_t shape, int32_t overlap,
                              const uint8_t* mask, int32_t* n_blocks, int32_t* n_fields, int32_t* n_tasks) {
    psm::Plan P;
    int rc = psm::compile_plan(variant, grid_h, grid_w, shape, overlap, mask, P);
    if (rc != 0) return rc;
    if (n_blocks) *n_blocks = P.B;
    if (n_fields) *n_fields = P.F;
    if (n_tasks) *n_tasks = (int32_t)P.tasks.size();
    return 0;
}

extern "C" int psm_plan_compile(int32_t variant, int32_t grid_h, int32_t grid_w, int32_t shape, int32_t overlap,
                                const uint8_t* mask, int32_t* origins, int32_t* indices_list, int32_t* owner,
                                int32_t* rec, int32_t* tasks) {
    psm::Plan P;
    int rc = psm::compile_plan(variant, grid_h, grid_w, shape, overlap, mask, P);
    if (rc != 0) return rc;
    for (int k = 0; k < P.B; ++k) {
        if (origins) { origins[2 * k] = P.y0[k]; origins[2 * k + 1] = P.x0[k]; }
        if (indices_list) { indices_list[2 * k] = P.idx_i[k]; indices_list[2 * k + 1] = P.idx_j[k]; }
    }
    if (owner) for (size_t q = 0; q < P.owner.size(); ++q) owner[q] = P.owner[q];
    if (rec) for (size_t q = 0; q < P.rec.size(); ++q) {
        rec[4 * q] = P.rec[q].ta; rec[4 * q + 1] = P.rec[q].tb; rec[4 * q + 2] = P.rec[q].parent; rec[4 * q + 3] = P.rec[q].is_nan;
    }
    if (tasks) for (size_t q = 0; q < P.tasks.size(); ++q) {
        const psm::Task& t = P.tasks[q];
        int32_t* o = tasks + 8 * q;
        o[0] = t.src; o[1] = t.msk; o[2] = t.ch; o[3] = t.y0; o[4] = t.y1; o[5] = t.x0; o[6] = t.x1; o[7] = t.count;
    }
    return 0;
}

// Shift-line runs (SMC:350 / GRAD:358-361 through the owner map): int32[n][8] =
// (field, block, y0, y1, x0, x1, coef, n_pixels); `lines` may be NULL to query the count.
extern "C" int psm_plan_shift_lines(int32_t variant, int32_t grid_h, int32_t grid_w, int32_t shape, int32_t overlap,
                                    const uint8_t* mask, int32_t* n_lines, int32_t* lines) {
    psm::Plan P;
    int rc = psm::compile_plan(variant, grid_h, grid_w, shape, overlap, mask, P);
    if (rc != 0) return rc;
    int n = 0;
    for (int f = 0; f < P.F; ++f)
        for (const psm::LineTask& t : P.lines[f]) {
            if (lines) {
                int32_t* o = lines + 8 * n;
                o[0] = f; o[1] = t.src; o[2] = t.y0; o[3] = t.y1; o[4] = t.x0; o[5] = t.x1; o[6] = t.coef; o[7] = t.n;
            }
            ++n;
        }
    if (n_lines) *n_lines = n;
    return 0;
}
