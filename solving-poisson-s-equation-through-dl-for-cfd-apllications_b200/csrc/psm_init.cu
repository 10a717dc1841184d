// Once-per-mesh initialisation behind the C-ABI without Python: everything init_func does (PMP:172-247,
// SMC:89-180) except the cells -> grid Delaunay, which stays with Qhull (the library the reference calls, through SciPy) and is
// handed in as tables -- or comes from the on-disk table cache keyed by a hash of the mesh, so that only the FIRST run of a case
// needs an interpreter.
//
//   I1  bounding box rounded like Python's round(x, nd) (SMC:102-106: 3 decimals; PMP:197-201 / GRAD:174-178: 2), uniform grid
//       like np.linspace (UTL:111-125)                                                              host, bit-identical
//   I3  flow mask (inside the bbox test of SMC:120-126 / PMP:76-84, outside the obstacle's convex hull SMC:128-136) and distance
//       to the nearest sub-sampled wall point (SMC:138-143: cdist + min)                            one CUDA kernel, FP64
//   I4  index raster + distance field (SMC:161-178 / PMP:225-243)                                   host loop over the pixels
//   back tables in closed form (psm_b200/tables.py regular_grid_back_tables, opt-in like there)     host, bit-identical
//   block-row partition of global tables into one rank's psm_shard (psm_b200/shard.py partition)   host
#include <cuda_runtime.h>
#include <math_constants.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/psm_b200.h"
#include "psm_internal.h"

using psm::handle_fail;

namespace {

// Python's round(x, nd) for a float: the correctly rounded decimal string, parsed back (glibc printf is exact).
double py_round(double x, int nd) {
    char b[64];
    snprintf(b, sizeof b, "%.*f", nd, x);
    return strtod(b, nullptr);
}
// np.linspace(start, stop, num): arange(num) * step + start, last element = stop
void linspace(double start, double stop, int num, std::vector<double>& out) {
    out.resize(num);
    if (num == 1) { out[0] = start; return; }
    const double step = (stop - start) / (double)(num - 1);
    for (int k = 0; k < num; ++k) out[k] = (double)k * step + start;
    out[num - 1] = stop;
}

// Andrew's monotone chain; counter-clockwise, collinear points dropped (the polygon is what matters: SMC:128-136)
void convex_hull(const double* pts, long long n, std::vector<double>& hull) {
    std::vector<std::pair<double, double>> p(n);
    for (long long i = 0; i < n; ++i) p[i] = {pts[2 * i], pts[2 * i + 1]};
    std::sort(p.begin(), p.end());
    p.erase(std::unique(p.begin(), p.end()), p.end());
    const long long m = (long long)p.size();
    if (m < 3) { hull.clear(); for (auto& q : p) { hull.push_back(q.first); hull.push_back(q.second); } return; }
    std::vector<std::pair<double, double>> h(2 * m);
    auto cross = [](const std::pair<double, double>& o, const std::pair<double, double>& a, const std::pair<double, double>& b) {
        return (a.first - o.first) * (b.second - o.second) - (a.second - o.second) * (b.first - o.first);
    };
    long long k = 0;
    for (long long i = 0; i < m; ++i) { while (k >= 2 && cross(h[k - 2], h[k - 1], p[i]) <= 0) --k; h[k++] = p[i]; }
    for (long long i = m - 2, t = k + 1; i >= 0; --i) { while (k >= t && cross(h[k - 2], h[k - 1], p[i]) <= 0) --k; h[k++] = p[i]; }
    hull.clear();
    for (long long i = 0; i + 1 < k; ++i) { hull.push_back(h[i].first); hull.push_back(h[i].second); }
}

// FNV-1a, 64 bit
struct Fnv {
    unsigned long long h = 1469598103934665603ull;
    void add(const void* p, size_t n) { const unsigned char* b = (const unsigned char*)p; for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; } }
};

}  // namespace

// ---- I3 on the GPU ------------------------------------------------------------------------------------------
// One thread per grid point: bbox test, strict point-in-convex-polygon (CCW hull: every edge cross product > 0), then the
// minimum Euclidean distance over the two sub-sampled boundary point sets staged through shared memory.  The squares and the sum
// are rounded separately (no FMA) like cdist / cKDTree.
__global__ void __launch_bounds__(256) init_mask_sdf_kernel(const double* __restrict__ X0, const double* __restrict__ Y0, int H, int W,
                                                            double min_x, double max_x, double min_y, double max_y,
                                                            const double* __restrict__ hull, int n_hull,
                                                            const double* __restrict__ pts, int n_pts,       // obst[::step] then top[::step]
                                                            unsigned char* __restrict__ domain, double* __restrict__ sdf) {
    extern __shared__ double sh[];
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long G = (long long)H * W;
    const bool live = q < G;
    double x = 0.0, y = 0.0;
    if (live) { x = X0[q % W]; y = Y0[q / W]; }
    bool inside = live && (x <= max_x) && (x >= min_x) && (y <= max_y) && (y >= min_y);
    bool in_obst = n_hull >= 3;
    for (int k0 = 0; k0 < n_hull; k0 += 512) {
        const int m = min(512, n_hull - k0);
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * (m + 1); i += blockDim.x) sh[i] = hull[(2 * k0 + i) % (2 * n_hull)];     // + the closing vertex
        __syncthreads();
        for (int k = 0; k < m; ++k) {
            const double ax = sh[2 * k], ay = sh[2 * k + 1], bx = sh[2 * k + 2], by = sh[2 * k + 3];
            const double cr = __dsub_rn(__dmul_rn(bx - ax, y - ay), __dmul_rn(by - ay, x - ax));
            in_obst = in_obst && (cr > 0.0);
        }
    }
    double best = CUDART_INF;
    for (int k0 = 0; k0 < n_pts; k0 += 1024) {
        const int m = min(1024, n_pts - k0);
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * m; i += blockDim.x) sh[i] = pts[2 * k0 + i];
        __syncthreads();
        for (int k = 0; k < m; ++k) {
            const double dx = x - sh[2 * k], dy = y - sh[2 * k + 1];
            const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
            best = fmin(best, d2);
        }
    }
    if (live) {
        const bool dom = inside && !in_obst;
        domain[q] = dom ? 1 : 0;
        sdf[q] = dom ? sqrt(best) : 0.0;
    }
}

// ---- host-only pieces (exported so that the CPU test-suite can check them against the NumPy shim) ----------------------

extern "C" int psm_mesh_hash(int32_t variant, double delta, const double* cells_xy, int32_t xy_stride, int64_t n_cells,
                             const double* top, int64_t n_top, const double* obst, int64_t n_obst, const double* probe, char out[17]) {
    if (!cells_xy || !top || !obst || !out || n_cells < 1 || xy_stride < 2) return PSM_ERR_INVALID;
    Fnv f;
    f.add("PSMMESH1", 8); f.add(&variant, 4); f.add(&delta, 8); f.add(&n_cells, 8); f.add(&n_top, 8); f.add(&n_obst, 8);
    for (int64_t i = 0; i < n_cells; ++i) f.add(cells_xy + i * xy_stride, 16);
    f.add(top, (size_t)n_top * 16); f.add(obst, (size_t)n_obst * 16);
    // the probe field only decides validity where its interpolation is NaN (SMC:165-169): a finite probe leaves the tables
    // untouched, so its VALUES (the solver's initial pressure, different from run to run) stay out of the key
    bool probe_nan = false;
    if (probe) for (int64_t i = 0; i < n_cells && !probe_nan; ++i) probe_nan = probe[i] != probe[i];
    if (probe_nan) f.add(probe, (size_t)n_cells * 8);
    snprintf(out, 17, "%016llx", f.h);
    return PSM_OK;
}

extern "C" int psm_mesh_grid(int32_t variant, double delta, const double* cells_xy, int32_t xy_stride, int64_t n_cells,
                             double bbox[4], int32_t* grid_h, int32_t* grid_w) {
    if (!cells_xy || !bbox || !grid_h || !grid_w || n_cells < 1 || xy_stride < 2 || !(delta > 0)) return PSM_ERR_INVALID;
    const int nd = (variant == PSM_DELTAU_TO_DELTAP) ? 3 : 2;                    // SMC:102-106 vs GRAD:174-178 / PMP:197-201
    double xmin = cells_xy[0], xmax = xmin, ymin = cells_xy[1], ymax = ymin;
    for (int64_t i = 1; i < n_cells; ++i) {
        const double x = cells_xy[i * xy_stride], y = cells_xy[i * xy_stride + 1];
        xmin = x < xmin ? x : xmin; xmax = x > xmax ? x : xmax; ymin = y < ymin ? y : ymin; ymax = y > ymax ? y : ymax;
    }
    bbox[0] = py_round(xmin, nd); bbox[1] = py_round(xmax, nd); bbox[2] = py_round(ymin, nd); bbox[3] = py_round(ymax, nd);
    *grid_h = (int32_t)nearbyint((bbox[3] - bbox[2]) / delta);                   // int(round(.)): half to even (SMC:148-149)
    *grid_w = (int32_t)nearbyint((bbox[1] - bbox[0]) / delta);
    return (*grid_h > 0 && *grid_w > 0) ? PSM_OK : PSM_ERR_GEOMETRY;
}

// Grid -> cell tables in closed form (psm_b200/tables.py regular_grid_back_tables, same arithmetic in the same order).
extern "C" int psm_back_tables_closed_form(const double* cells_xy, int32_t xy_stride, int64_t n_cells, const double* X0_row, int32_t W,
                                           const double* Y0_col, int32_t H, int32_t* vert_back, double* weights_back) {
    if (!cells_xy || !X0_row || !Y0_col || !vert_back || !weights_back || W < 2 || H < 2) return PSM_ERR_INVALID;
    const double x0 = X0_row[0], y0 = Y0_col[0];
    const double dx = (X0_row[W - 1] - X0_row[0]) / (double)(W - 1), dy = (Y0_col[H - 1] - Y0_col[0]) / (double)(H - 1);
    for (int64_t c = 0; c < n_cells; ++c) {
        const double fx = (cells_xy[c * xy_stride] - x0) / dx, fy = (cells_xy[c * xy_stride + 1] - y0) / dy;
        const bool outside = (fx < 0) || (fy < 0) || (fx > W - 1) || (fy > H - 1);
        long long gj = (long long)std::floor(fx), gi = (long long)std::floor(fy);
        gj = gj < 0 ? 0 : (gj > W - 2 ? W - 2 : gj);
        gi = gi < 0 ? 0 : (gi > H - 2 ? H - 2 : gi);
        const double u = fx - (double)gj, v = fy - (double)gi;
        const bool lower = u >= v;
        const long long p00 = gi * W + gj, p10 = p00 + 1, p11 = (gi + 1) * W + gj + 1, p01 = (gi + 1) * W + gj;
        int32_t* vb = vert_back + 3 * c; double* wb = weights_back + 3 * c;
        if (lower) { vb[0] = (int32_t)p00; vb[1] = (int32_t)p10; vb[2] = (int32_t)p11; wb[0] = 1 - u; wb[1] = u - v; wb[2] = v; }
        else { vb[0] = (int32_t)p00; vb[1] = (int32_t)p11; vb[2] = (int32_t)p01; wb[0] = 1 - v; wb[1] = u; wb[2] = v - u; }
        if (outside) { wb[0] = -1.0; wb[1] = 1.0; wb[2] = 1.0; }
    }
    return PSM_OK;
}

namespace {
struct MeshTables {
    int H = 0, W = 0; double bbox[4] = {0, 0, 0, 0};
    std::vector<double> X0, Y0;                   // one row / one column of the grid
    std::vector<int64_t> indices; std::vector<double> sdfunct;
    std::vector<int32_t> vb; std::vector<double> wb;
};

int build_mesh_tables(psm_handle* h, int variant, double delta, const psm_mesh* m, MeshTables& T) {
    int rc = psm_mesh_grid(variant, delta, m->cells_xy, m->xy_stride, m->n_cells, T.bbox, &T.H, &T.W);
    if (rc) return handle_fail(h, rc, "psm_init_mesh: degenerate bounding box");
    const int H = T.H, W = T.W;
    const long long G = (long long)H * W;
    linspace(T.bbox[0] + delta / 2, T.bbox[1] - delta / 2, W, T.X0);              // UTL:111-125
    linspace(T.bbox[2] + delta / 2, T.bbox[3] - delta / 2, H, T.Y0);
    // ---- I3: mask + distance on the GPU
    double tmin_x = m->top[0], tmax_x = tmin_x, tmin_y = m->top[1], tmax_y = tmin_y;
    for (long long i = 1; i < m->n_top; ++i) {
        tmin_x = std::min(tmin_x, m->top[2 * i]); tmax_x = std::max(tmax_x, m->top[2 * i]);
        tmin_y = std::min(tmin_y, m->top[2 * i + 1]); tmax_y = std::max(tmax_y, m->top[2 * i + 1]);
    }
    double max_x, max_y, min_x, min_y; int step;
    if (variant == PSM_DELTAU_TO_DELTAP) {                                          // SMC:120-121, literally
        max_x = std::max(tmax_x, T.bbox[1]); max_y = std::min(tmax_y, T.bbox[3]);
        min_x = std::max(tmin_x, T.bbox[0]); min_y = std::min(tmin_y, T.bbox[2]);
        step = 5;                                                                   // SMC:138-139
    } else {
        max_x = tmax_x; max_y = tmax_y; min_x = tmin_x; min_y = tmin_y;
        step = (variant == PSM_U_TO_GRADP) ? 2 : 10;                                // GRAD:207-208 / PMP:92-93
    }
    std::vector<double> hull, pts;
    convex_hull(m->obst, m->n_obst, hull);
    for (long long i = 0; i < m->n_obst; i += step) { pts.push_back(m->obst[2 * i]); pts.push_back(m->obst[2 * i + 1]); }
    for (long long i = 0; i < m->n_top; i += step) { pts.push_back(m->top[2 * i]); pts.push_back(m->top[2 * i + 1]); }
    const int n_hull = (int)(hull.size() / 2), n_pts = (int)(pts.size() / 2);
    double *dX = nullptr, *dY = nullptr, *dH = nullptr, *dP = nullptr, *dS = nullptr; unsigned char* dD = nullptr;
    std::vector<unsigned char> domain(G); std::vector<double> sdf(G);
    bool ok = cudaMalloc(&dX, W * 8) == cudaSuccess && cudaMalloc(&dY, H * 8) == cudaSuccess && cudaMalloc(&dH, std::max<size_t>(hull.size(), 2) * 8) == cudaSuccess &&
              cudaMalloc(&dP, std::max<size_t>(pts.size(), 2) * 8) == cudaSuccess && cudaMalloc(&dS, G * 8) == cudaSuccess && cudaMalloc(&dD, G) == cudaSuccess;
    if (ok) {
        cudaMemcpy(dX, T.X0.data(), W * 8, cudaMemcpyHostToDevice); cudaMemcpy(dY, T.Y0.data(), H * 8, cudaMemcpyHostToDevice);
        cudaMemcpy(dH, hull.data(), hull.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(dP, pts.data(), pts.size() * 8, cudaMemcpyHostToDevice);
        init_mask_sdf_kernel<<<(unsigned)((G + 255) / 256), 256, 2 * 1025 * sizeof(double)>>>(dX, dY, H, W, min_x, max_x, min_y, max_y, dH, n_hull, dP, n_pts, dD, dS);
        ok = cudaMemcpy(domain.data(), dD, G, cudaMemcpyDeviceToHost) == cudaSuccess && cudaMemcpy(sdf.data(), dS, G * 8, cudaMemcpyDeviceToHost) == cudaSuccess &&
             cudaGetLastError() == cudaSuccess;
    }
    cudaFree(dX); cudaFree(dY); cudaFree(dH); cudaFree(dP); cudaFree(dS); cudaFree(dD);
    if (!ok) return handle_fail(h, PSM_ERR_CUDA, "psm_init_mesh: mask / distance kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
    // ---- I4: raster (SMC:161-178): valid = inside the flow AND the probe interpolates to a number (no negative weight, UTL:89)
    T.indices.assign((size_t)G * 2, 0); T.sdfunct.assign((size_t)G, 0.0);
    const double x0m = T.X0[0], y0m = T.Y0[0];
    for (long long q = 0; q < G; ++q) {
        if (!domain[q]) continue;
        const int32_t* v = m->vert + 3 * q; const double* w = m->weights + 3 * q;
        if (w[0] < 0 || w[1] < 0 || w[2] < 0) continue;
        if (m->probe) {
            const double pv = (m->probe[v[0]] * w[0] + m->probe[v[1]] * w[1]) + m->probe[v[2]] * w[2];
            if (pv != pv) continue;
        }
        const long long jj = (long long)nearbyint((T.X0[q % W] - x0m) / delta), ii = (long long)nearbyint((T.Y0[q / W] - y0m) / delta);
        if (ii < 0 || ii >= H || jj < 0 || jj >= W) return handle_fail(h, PSM_ERR_GEOMETRY, "raster index out of the grid at point %lld", q);
        T.indices[2 * q] = ii; T.indices[2 * q + 1] = jj;
        T.sdfunct[(size_t)ii * W + jj] = sdf[q];
    }
    return PSM_OK;
}
}  // namespace

extern "C" int psm_init_mesh(psm_handle* h, const psm_mesh* m) {
    if (!h) return PSM_ERR_INVALID;
    if (!m || !m->cells_xy || !m->top || !m->obst || m->n_cells < 3 || m->n_top < 1 || m->n_obst < 3 || m->xy_stride < 2)
        return handle_fail(h, PSM_ERR_INVALID, "psm_init_mesh: bad mesh description");
    try {
        const int variant = psm::handle_variant(h);
        const double delta = psm::handle_delta(h);
        if (cudaSetDevice(psm::handle_device(h)) != cudaSuccess) return handle_fail(h, PSM_ERR_CUDA, "cudaSetDevice failed");
        std::string cache;
        if (m->cache_dir && m->cache_dir[0]) {
            char key[17];
            psm_mesh_hash(variant, delta, m->cells_xy, m->xy_stride, m->n_cells, m->top, m->n_top, m->obst, m->n_obst, m->probe, key);
            cache = std::string(m->cache_dir) + "/psm_tables_" + key + (m->back_closed_form ? "_cf" : "_qh") + ".bin";
            if (FILE* f = fopen(cache.c_str(), "rb")) {           // hit: no Delaunay, no distance field, one sequential read
                fclose(f);
                return psm_init_from_file(h, cache.c_str());
            }
        }
        if (!m->vert || !m->weights)
            return handle_fail(h, PSM_ERR_STATE, "psm_init_mesh: no cells -> grid tables were given and the table cache has no entry for this mesh "
                                                 "(build them once with the Python shim: Qhull through SciPy, as the reference does)");
        MeshTables T;
        int rc = build_mesh_tables(h, variant, delta, m, T);
        if (rc) return rc;
        psm_tables t{};
        t.n_cells = m->n_cells; t.grid_h = T.H; t.grid_w = T.W; t.vert = m->vert; t.weights = m->weights;
        t.indices = T.indices.data(); t.sdfunct = T.sdfunct.data();
        const long long G = (long long)T.H * T.W;
        for (long long q = 0; q < 3 * G; ++q)
            if (m->vert[q] < 0 || m->vert[q] >= m->n_cells) return handle_fail(h, PSM_ERR_INVALID, "psm_init_mesh: vert out of range");
        if (m->vert_back && m->weights_back) { t.vert_back = m->vert_back; t.weights_back = m->weights_back; }
        else if (m->back_closed_form) {
            T.vb.resize((size_t)m->n_cells * 3); T.wb.resize((size_t)m->n_cells * 3);
            rc = psm_back_tables_closed_form(m->cells_xy, m->xy_stride, m->n_cells, T.X0.data(), T.W, T.Y0.data(), T.H, T.vb.data(), T.wb.data());
            if (rc) return handle_fail(h, rc, "psm_init_mesh: closed-form back tables need a grid of at least 2 x 2");
            t.vert_back = T.vb.data(); t.weights_back = T.wb.data();
        }
        if (!cache.empty()) psm_save_tables(&t, cache.c_str());     // best effort: a read-only cache directory is not an error
        return psm_init_with_tables(h, &t);
    } catch (const std::exception& e) {
        return handle_fail(h, PSM_ERR_INVALID, "psm_init_mesh: %s", e.what());
    } catch (...) {
        return handle_fail(h, PSM_ERR_INVALID, "psm_init_mesh: unknown exception");
    }
}
