// Block-row partitioner in C++ (host only): one rank's psm_shard from the GLOBAL once-per-mesh tables, so that a C / C++ caller
// (the OpenFOAM adapter) can shard a mesh without the Python shim.  Same arithmetic, same orderings and the same outputs as
// psm_b200/shard.py `partition(..., halo='cells')` -- tests/test_partition_cpu.py compares every array.
//
// Replaces the gather-to-root of the reference (PMP:179-185, 258, 501-511): the grid is cut by BLOCK ROWS of the extraction plan
// (SMC:461-479 / GRAD:479-500); rank g gathers and places pixel rows [row0, row1), evaluates the blocks of its block rows and owns
// the cells whose centre falls into its pixel rows.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <exception>
#include <new>
#include <vector>

#include "../../include/psm_b200.h"

struct psm_shard_owned {
    psm_shard view{};
    std::vector<uint8_t> mask;
    std::vector<int32_t> vert, vert_back, cell_send_idx, pix_send_idx, cell_rank;
    std::vector<double> weights, sdf, weights_back;
    std::vector<int64_t> cell_send_ptr, cell_recv_ptr, pix_send_ptr, pix_recv_ptr, ghost_pix, owned_ids, ghost_ids;
};

namespace {

// psm_b200/shard.py block_row_split
bool block_row_split(int H, int shape, int overlap, int world, std::vector<std::pair<int, int>>& ranges, std::vector<std::pair<int, int>>& rows) {
    const int stride = shape - overlap;
    const int n_y = (H - shape) / stride;
    const int n_reg = n_y + 1;
    const int need = (overlap + stride - 1) / stride;
    if (world > n_reg || (world > 1 && n_reg / world < need)) return false;
    std::vector<int> cuts(world + 1);
    for (int g = 0; g <= world; ++g) cuts[g] = (int)std::nearbyint((double)g * n_reg / world);      // Python round(): half to even
    for (int g = 0; g < world; ++g) {
        const bool last = g == world - 1;
        ranges.push_back({cuts[g], last ? cuts[g + 1] + 1 : cuts[g + 1]});
        rows.push_back({cuts[g] * stride, last ? H : cuts[g + 1] * stride});
    }
    return true;
}

// ghosts of rank g: cells its forward rows reference but does not own, ordered by (owner, id), small holes closed
// (psm_b200/shard.py _fill_ghost_gaps)
void fill_ghost_gaps(std::vector<int64_t>& ghost, const std::vector<int32_t>& cell_rank, int max_gap = 16, double slack = 0.25) {
    std::vector<int64_t> out;
    size_t a = 0;
    while (a < ghost.size()) {
        const int o = cell_rank[ghost[a]];
        size_t b = a;
        while (b < ghost.size() && cell_rank[ghost[b]] == o) ++b;
        long long extra = 0; bool any = false;
        for (size_t k = a; k + 1 < b; ++k) { const long long d = ghost[k + 1] - ghost[k]; if (d > 1 && d <= max_gap + 1) { extra += d - 1; any = true; } }
        if (!any || (double)extra > slack * (double)(b - a) + 64.0) { out.insert(out.end(), ghost.begin() + a, ghost.begin() + b); a = b; continue; }
        for (size_t k = a; k < b; ++k) {
            out.push_back(ghost[k]);
            if (k + 1 < b) {
                const long long d = ghost[k + 1] - ghost[k];
                if (d > 1 && d <= max_gap + 1)
                    for (long long c = ghost[k] + 1; c < ghost[k + 1]; ++c) if (cell_rank[c] == o) out.push_back(c);
            }
        }
        a = b;
    }
    ghost.swap(out);
}

}  // namespace

extern "C" int psm_shard_build(const psm_tables* t, const double* cells_xy, int32_t xy_stride, double y_min, double delta, int32_t variant,
                               int32_t shape, int32_t overlap, double near_wall_sdf, int32_t rank, int32_t world, psm_shard_owned** out) {
    if (!t || !cells_xy || !out || !t->vert || !t->weights || !t->indices || !t->sdfunct || xy_stride < 2 || world < 1 || rank < 0 || rank >= world ||
        shape <= overlap || !(delta > 0))
        return PSM_ERR_INVALID;
    *out = nullptr;
    try {
        const int H = t->grid_h, W = t->grid_w;
        const long long G = (long long)H * W, N = t->n_cells;
        std::vector<std::pair<int, int>> ranges, rows;
        if (!block_row_split(H, shape, overlap, world, ranges, rows)) return PSM_ERR_GEOMETRY;
        psm_shard_owned* S = new psm_shard_owned();
        // ---- fold_forward_table: what pixel q receives (last source point wins; negative weight -> zero)
        std::vector<long long> src(G, -1);
        for (long long m = 0; m < G; ++m) {
            const long long ii = t->indices[2 * m], jj = t->indices[2 * m + 1];
            if (ii < 0 || ii >= H || jj < 0 || jj >= W) { delete S; return PSM_ERR_INVALID; }
            src[ii * W + jj] = m;
        }
        std::vector<int32_t> fv((size_t)G * 3, 0); std::vector<double> fw((size_t)G * 3, 0.0);
        for (long long q = 0; q < G; ++q) {
            const long long m = src[q];
            if (m < 0) continue;
            const double* wm = t->weights + 3 * m;
            const bool neg = wm[0] < 0 || wm[1] < 0 || wm[2] < 0;
            for (int j = 0; j < 3; ++j) { fv[3 * q + j] = t->vert[3 * m + j]; fw[3 * q + j] = neg ? 0.0 : wm[j]; }
        }
        // ---- owner of every cell: the rank whose pixel rows contain it
        S->cell_rank.resize(N);
        std::vector<std::vector<int64_t>> owned(world);
        std::vector<int64_t> local_of(N);
        for (long long c = 0; c < N; ++c) {
            long long r = (long long)std::floor((cells_xy[c * xy_stride + 1] - y_min) / delta);
            r = r < 0 ? 0 : (r > H - 1 ? H - 1 : r);
            int g = 0;
            while (g + 1 < world && rows[g + 1].first <= r) ++g;
            S->cell_rank[c] = g;
            local_of[c] = (int64_t)owned[g].size();
            owned[g].push_back(c);
        }
        // ---- hop_back_table
        const bool have_back = t->vert_back && t->weights_back;
        std::vector<long long> bv; std::vector<uint8_t> keep;
        if (have_back) {
            bv.resize((size_t)N * 3); keep.assign(N, 0);
            for (long long c = 0; c < N; ++c) {
                const double* wc = t->weights_back + 3 * c;
                bool k = wc[0] < 0 || wc[1] < 0 || wc[2] < 0;
                double sdf_mesh = 0.0;
                for (int j = 0; j < 3; ++j) {
                    const long long gq = t->vert_back[3 * c + j];
                    if (gq < 0 || gq >= G) { delete S; return PSM_ERR_INVALID; }
                    bv[3 * c + j] = t->indices[2 * gq] * W + t->indices[2 * gq + 1];
                    sdf_mesh += t->sdfunct[gq] * wc[j];
                }
                if (near_wall_sdf > 0 && !k && sdf_mesh < near_wall_sdf) k = true;
                keep[c] = k;
            }
        }
        auto pix_rank = [&](long long q) { const long long r = q / W; int g = 0; while (g + 1 < world && rows[g + 1].first <= r) ++g; return g; };
        // ---- ghost cells and ghost pixels of EVERY rank (the send lists of this rank follow from the others' ghost lists)
        std::vector<std::vector<int64_t>> ghosts(world), gpix(world);
        std::vector<uint8_t> mark(N, 0);
        for (int g = 0; g < world; ++g) {
            const long long q0 = (long long)rows[g].first * W, q1 = (long long)rows[g].second * W;
            const int ext = (g == world - 1) ? 0 : overlap;
            std::vector<int64_t> need;
            for (long long q = q0; q < q1 + (long long)ext * W; ++q) {
                if (fw[3 * q] == 0.0 && fw[3 * q + 1] == 0.0 && fw[3 * q + 2] == 0.0) continue;       // not live
                for (int j = 0; j < 3; ++j) { const int32_t c = fv[3 * q + j]; if (!mark[c]) { mark[c] = 1; need.push_back(c); } }
            }
            for (int64_t c : need) mark[c] = 0;
            std::vector<int64_t>& gh = ghosts[g];
            for (int64_t c : need) if (S->cell_rank[c] != g) gh.push_back(c);
            std::sort(gh.begin(), gh.end(), [&](int64_t a, int64_t b) { return S->cell_rank[a] != S->cell_rank[b] ? S->cell_rank[a] < S->cell_rank[b] : a < b; });
            fill_ghost_gaps(gh, S->cell_rank);
            if (have_back) {
                std::vector<int64_t>& gp = gpix[g];
                for (int64_t c : owned[g]) {
                    if (keep[c]) continue;
                    for (int j = 0; j < 3; ++j) { const long long q = bv[3 * c + j]; if (q < q0 || q >= q1) gp.push_back(q); }
                }
                std::sort(gp.begin(), gp.end());
                gp.erase(std::unique(gp.begin(), gp.end()), gp.end());
            }
        }
        // ---- this rank
        const int g = rank;
        const int r0 = rows[g].first, r1 = rows[g].second;
        const long long q0 = (long long)r0 * W, q1 = (long long)r1 * W;
        const int ext = (g == world - 1) ? 0 : overlap;
        const std::vector<int64_t>& gh = ghosts[g];
        std::vector<int64_t> lut(N, 0);
        for (size_t k = 0; k < owned[g].size(); ++k) lut[owned[g][k]] = (int64_t)k;
        for (size_t k = 0; k < gh.size(); ++k) lut[gh[k]] = (int64_t)(owned[g].size() + k);
        const long long nq = q1 - q0 + (long long)ext * W;
        S->vert.resize((size_t)nq * 3); S->weights.resize((size_t)nq * 3);
        for (long long q = 0; q < nq; ++q) {
            const double* w = &fw[3 * (q0 + q)];
            const bool live = w[0] != 0.0 || w[1] != 0.0 || w[2] != 0.0;
            for (int j = 0; j < 3; ++j) { S->vert[3 * q + j] = live ? (int32_t)lut[fv[3 * (q0 + q) + j]] : 0; S->weights[3 * q + j] = w[j]; }
        }
        S->mask.resize(G);
        for (long long q = 0; q < G; ++q) S->mask[q] = t->sdfunct[q] != 0.0;
        S->sdf.assign(t->sdfunct + q0, t->sdfunct + q1 + (long long)ext * W);
        S->owned_ids = owned[g]; S->ghost_ids = gh;
        S->cell_recv_ptr.assign(world + 1, 0);
        for (int64_t c : gh) ++S->cell_recv_ptr[S->cell_rank[c] + 1];
        for (int p = 0; p < world; ++p) S->cell_recv_ptr[p + 1] += S->cell_recv_ptr[p];
        S->pix_recv_ptr.assign(world + 1, 0);
        if (have_back) {
            const std::vector<int64_t>& gp = gpix[g];
            const long long no = (long long)owned[g].size();
            S->vert_back.resize((size_t)no * 3); S->weights_back.resize((size_t)no * 3);
            for (long long k = 0; k < no; ++k) {
                const int64_t c = owned[g][k];
                for (int j = 0; j < 3; ++j) {
                    const long long q = bv[3 * c + j];
                    long long lb;
                    if (q >= q0 && q < q1) lb = q - q0;
                    else if (!gp.empty()) {
                        long long pos = std::lower_bound(gp.begin(), gp.end(), q) - gp.begin();
                        if (pos > (long long)gp.size() - 1) pos = (long long)gp.size() - 1;
                        lb = (q1 - q0) + pos;
                    } else lb = 0;
                    if (keep[c]) lb = (j == 0) ? -1 : 0;
                    S->vert_back[3 * k + j] = (int32_t)lb;
                    S->weights_back[3 * k + j] = t->weights_back[3 * c + j];
                }
            }
            S->ghost_pix = gp;
            for (int64_t q : gp) ++S->pix_recv_ptr[pix_rank(q) + 1];
            for (int p = 0; p < world; ++p) S->pix_recv_ptr[p + 1] += S->pix_recv_ptr[p];
        }
        // send lists: rank `rank` sends to g2 exactly g2's ghosts it owns, in g2's ghost order
        S->cell_send_ptr.assign(world + 1, 0); S->pix_send_ptr.assign(world + 1, 0);
        for (int g2 = 0; g2 < world; ++g2) {
            if (g2 != rank) {
                for (int64_t c : ghosts[g2]) if (S->cell_rank[c] == rank) S->cell_send_idx.push_back((int32_t)local_of[c]);
                for (int64_t q : gpix[g2]) if (q >= q0 && q < q1) S->pix_send_idx.push_back((int32_t)(q - q0));
            }
            S->cell_send_ptr[g2 + 1] = (int64_t)S->cell_send_idx.size();
            S->pix_send_ptr[g2 + 1] = (int64_t)S->pix_send_idx.size();
        }
        psm_shard& v = S->view;
        v.rank = rank; v.world = world; v.grid_h = H; v.grid_w = W; v.row0 = r0; v.row1 = r1;
        v.ext_rows = 0; v.send_rows = 0; v.local_ext_rows = ext;                         // halo 'cells': the overlap rows are gathered locally
        v.blk_row0 = ranges[g].first; v.blk_row1 = ranges[g].second;
        v.mask_global = S->mask.data();
        v.n_owned = (int64_t)owned[g].size(); v.n_ghost = (int64_t)gh.size(); v.n_ghost_pix = (int64_t)S->ghost_pix.size();
        v.vert = S->vert.data(); v.weights = S->weights.data(); v.sdfunct = S->sdf.data();
        v.vert_back = have_back ? S->vert_back.data() : nullptr; v.weights_back = have_back ? S->weights_back.data() : nullptr;
        v.cell_send_ptr = S->cell_send_ptr.data(); v.cell_send_idx = S->cell_send_idx.data(); v.cell_recv_ptr = S->cell_recv_ptr.data();
        v.pix_send_ptr = S->pix_send_ptr.data(); v.pix_send_idx = S->pix_send_idx.data(); v.pix_recv_ptr = S->pix_recv_ptr.data();
        v.ghost_pix = S->ghost_pix.empty() ? nullptr : S->ghost_pix.data();
        (void)variant;
        *out = S;
        return PSM_OK;
    } catch (const std::exception&) {
        return PSM_ERR_INVALID;
    } catch (...) {
        return PSM_ERR_INVALID;
    }
}

extern "C" const psm_shard* psm_shard_view(const psm_shard_owned* s) { return s ? &s->view : nullptr; }
extern "C" int psm_shard_cells(const psm_shard_owned* s, const int64_t** owned_ids, const int32_t** cell_rank) {
    if (!s) return PSM_ERR_INVALID;
    if (owned_ids) *owned_ids = s->owned_ids.data();
    if (cell_rank) *cell_rank = s->cell_rank.data();
    return PSM_OK;
}
extern "C" int psm_shard_free(psm_shard_owned* s) { delete s; return PSM_OK; }
