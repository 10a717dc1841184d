// Flat binary containers for the once-per-mesh tables and the model artefacts, so that a C/C++ caller
// (INTEGRATION.md route B) needs neither SciPy nor an interpreter at run time.  Written by
// psm_save_tables / psm_save_params (from the Python shim, offline), read by psm_init_from_file /
// psm_load_params_file.  Little-endian, no padding:
//   tables: "PSMTBL01" | i64 n_cells | i32 H | i32 W | i32 has_back | i32 0
//           | vert i32[G*3] | weights f64[G*3] | indices i64[G*2] | sdfunct f64[G]
//           | (vert_back i32[N*3] | weights_back f64[N*3])
//   params: "PSMPRM01" | i32 shape, n_out, pc_in, pc_p, standardization, n_dense | f64 maxs[5]
//           | f64 max_abs_in, max_abs_out | i32 dims[n_dense+1]
//           | f64 comp_in[pc_in*S*S*3], mean_in_pca[S*S*3], comp_out[pc_p*S*S*C], mean_out_pca[S*S*C]
//           | (f64 mean_in[pc_in], std_in[pc_in], mean_out[pc_p], std_out[pc_p])   (PSM_STD; PSM_MIN_MAX stores min / max there)
//           | per layer: f32 kernel[in*out] (Keras [in][out]), f32 bias[out]
#include <cstdio>
#include <cstring>
#include <exception>
#include <vector>

#include "../../include/psm_b200.h"
#include "psm_internal.h"

using psm::handle_fail;

namespace {
struct File {
    FILE* f = nullptr;
    File(const char* p, const char* m) { f = p ? fopen(p, m) : nullptr; }
    ~File() { if (f) fclose(f); }
    template <typename T> bool wr(const T* p, size_t n) { return n == 0 || fwrite(p, sizeof(T), n, f) == n; }
    template <typename T> bool rd(T* p, size_t n) { return n == 0 || fread(p, sizeof(T), n, f) == n; }
    template <typename T> bool rdv(std::vector<T>& v, size_t n) { v.resize(n); return rd(v.data(), n); }
    // bytes from the current position to the end of the file (a corrupt header must not drive an allocation)
    long long remaining() {
        const long cur = ftell(f);
        if (cur < 0 || fseek(f, 0, SEEK_END) != 0) return -1;
        const long end = ftell(f);
        fseek(f, cur, SEEK_SET);
        return end < cur ? -1 : (long long)(end - cur);
    }
};
constexpr long long kMaxWidth = 1 << 16;      // PCA components / Dense widths accepted from a file
}  // namespace

namespace {
int init_from_file_body(psm_handle* h, const char* path) {
    File F(path, "rb");
    char magic[8];
    psm_tables t{};
    int32_t hb = 0, zero = 0;
    if (!F.f) return handle_fail(h, PSM_ERR_INVALID, "psm_init_from_file: cannot open %s", path);
    if (!F.rd(magic, 8) || memcmp(magic, "PSMTBL01", 8) != 0) return handle_fail(h, PSM_ERR_INVALID, "%s is not a PSMTBL01 table file", path);
    if (!F.rd(&t.n_cells, 1) || !F.rd(&t.grid_h, 1) || !F.rd(&t.grid_w, 1) || !F.rd(&hb, 1) || !F.rd(&zero, 1))
        return handle_fail(h, PSM_ERR_INVALID, "%s: truncated header", path);
    if (t.n_cells < 3 || t.n_cells > (1ll << 31) - 1 || t.grid_h < 1 || t.grid_w < 1 || (long long)t.grid_h * t.grid_w > (1ll << 31) - 1 || (hb != 0 && hb != 1))
        return handle_fail(h, PSM_ERR_INVALID, "%s: header out of range (n_cells %lld, grid %d x %d)", path, (long long)t.n_cells, t.grid_h, t.grid_w);
    const size_t G = (size_t)t.grid_h * t.grid_w, N = (size_t)t.n_cells;
    const long long want = (long long)G * (3 * 4 + 3 * 8 + 2 * 8 + 8) + (hb ? (long long)N * (3 * 4 + 3 * 8) : 0);
    if (F.remaining() != want)
        return handle_fail(h, PSM_ERR_INVALID, "%s: payload is %lld bytes, the header promises %lld", path, F.remaining(), want);
    std::vector<int32_t> vert, vb; std::vector<double> w, sdf, wb; std::vector<int64_t> idx;
    if (!F.rdv(vert, G * 3) || !F.rdv(w, G * 3) || !F.rdv(idx, G * 2) || !F.rdv(sdf, G) || (hb && (!F.rdv(vb, N * 3) || !F.rdv(wb, N * 3))))
        return handle_fail(h, PSM_ERR_INVALID, "%s: short read", path);
    t.vert = vert.data(); t.weights = w.data(); t.indices = idx.data(); t.sdfunct = sdf.data();
    t.vert_back = hb ? vb.data() : nullptr; t.weights_back = hb ? wb.data() : nullptr;
    return psm_init_with_tables(h, &t);
}
}  // namespace

extern "C" int psm_save_tables(const psm_tables* t, const char* path) {
    if (!t || !path || !t->vert || !t->weights || !t->indices || !t->sdfunct) return PSM_ERR_INVALID;
    File F(path, "wb");
    if (!F.f) return PSM_ERR_INVALID;
    const size_t G = (size_t)t->grid_h * t->grid_w, N = (size_t)t->n_cells;
    const int32_t hb = (t->vert_back && t->weights_back) ? 1 : 0, zero = 0;
    bool ok = F.wr("PSMTBL01", 8) && F.wr(&t->n_cells, 1) && F.wr(&t->grid_h, 1) && F.wr(&t->grid_w, 1) && F.wr(&hb, 1) && F.wr(&zero, 1) &&
              F.wr(t->vert, G * 3) && F.wr(t->weights, G * 3) && F.wr(t->indices, G * 2) && F.wr(t->sdfunct, G);
    if (ok && hb) ok = F.wr(t->vert_back, N * 3) && F.wr(t->weights_back, N * 3);
    return ok ? PSM_OK : PSM_ERR_INVALID;
}

extern "C" int psm_init_from_file(psm_handle* h, const char* path) {
    if (!h) return PSM_ERR_INVALID;
    if (!path) return handle_fail(h, PSM_ERR_INVALID, "psm_init_from_file: NULL path");
    try {
        return init_from_file_body(h, path);
    } catch (const std::exception& e) {            // bad_alloc / length_error: nothing may cross the extern "C" boundary
        return handle_fail(h, PSM_ERR_INVALID, "psm_init_from_file(%s): %s", path, e.what());
    } catch (...) {
        return handle_fail(h, PSM_ERR_INVALID, "psm_init_from_file(%s): unknown exception", path);
    }
}

extern "C" int psm_save_params(const psm_params* p, int32_t shape, const char* path) {
    if (!p || !path || shape < 1 || !p->layer_dims || !p->dense_kernels || !p->dense_biases) return PSM_ERR_INVALID;
    File F(path, "wb");
    if (!F.f) return PSM_ERR_INVALID;
    const size_t S2 = (size_t)shape * shape, Kin = S2 * 3, Kout = S2 * p->n_out_channels;
    const int32_t hdr[6] = {shape, p->n_out_channels, p->pc_in, p->pc_p, p->standardization, p->n_dense};
    bool ok = F.wr("PSMPRM01", 8) && F.wr(hdr, 6) && F.wr(p->maxs, 5) && F.wr(&p->max_abs_input_PCA, 1) && F.wr(&p->max_abs_output_PCA, 1) &&
              F.wr(p->layer_dims, (size_t)p->n_dense + 1) && F.wr(p->pca_in_components, p->pc_in * Kin) && F.wr(p->pca_in_mean, Kin) &&
              F.wr(p->pca_out_components, p->pc_p * Kout) && F.wr(p->pca_out_mean, Kout);
    if (ok && p->standardization != PSM_MAX_ABS)          // PSM_STD: mean / std; PSM_MIN_MAX: min / max (same four slots)
        ok = F.wr(p->mean_in, p->pc_in) && F.wr(p->std_in, p->pc_in) && F.wr(p->mean_out, p->pc_p) && F.wr(p->std_out, p->pc_p);
    for (int l = 0; ok && l < p->n_dense; ++l)
        ok = F.wr(p->dense_kernels[l], (size_t)p->layer_dims[l] * p->layer_dims[l + 1]) && F.wr(p->dense_biases[l], (size_t)p->layer_dims[l + 1]);
    return ok ? PSM_OK : PSM_ERR_INVALID;
}

namespace {
int load_params_file_body(psm_handle* h, const char* path) {
    File F(path, "rb");
    char magic[8];
    int32_t hdr[6];
    psm_params p{};
    if (!F.f) return handle_fail(h, PSM_ERR_INVALID, "psm_load_params_file: cannot open %s", path);
    if (!F.rd(magic, 8) || memcmp(magic, "PSMPRM01", 8) != 0) return handle_fail(h, PSM_ERR_INVALID, "%s is not a PSMPRM01 parameter file", path);
    if (!F.rd(hdr, 6) || !F.rd(p.maxs, 5) || !F.rd(&p.max_abs_input_PCA, 1) || !F.rd(&p.max_abs_output_PCA, 1))
        return handle_fail(h, PSM_ERR_INVALID, "%s: truncated header", path);
    const int32_t shape = hdr[0];
    p.n_out_channels = hdr[1]; p.pc_in = hdr[2]; p.pc_p = hdr[3]; p.standardization = hdr[4]; p.n_dense = hdr[5];
    // psm_load_params indexes the PCA matrices with the HANDLE's block edge: a file written for another shape is rejected here
    if (shape != psm::handle_shape(h))
        return handle_fail(h, PSM_ERR_INVALID, "%s was written for shape %d, the handle uses %d", path, shape, psm::handle_shape(h));
    if (p.n_out_channels < 1 || p.n_out_channels > 2 || p.n_dense < 1 || p.n_dense > 64 || p.standardization < PSM_STD || p.standardization > PSM_MIN_MAX)
        return handle_fail(h, PSM_ERR_INVALID, "%s: header out of range (n_out %d, n_dense %d, standardization %d)", path, p.n_out_channels, p.n_dense, p.standardization);
    const long long S2 = (long long)shape * shape, Kin = S2 * 3, Kout = S2 * p.n_out_channels;
    if (p.pc_in < 1 || p.pc_in > Kin || p.pc_in > kMaxWidth || p.pc_p < 1 || p.pc_p > Kout || p.pc_p > kMaxWidth)
        return handle_fail(h, PSM_ERR_INVALID, "%s: pc_in %d / pc_p %d out of range", path, p.pc_in, p.pc_p);
    std::vector<int32_t> dims;
    if (!F.rdv(dims, (size_t)p.n_dense + 1)) return handle_fail(h, PSM_ERR_INVALID, "%s: truncated layer list", path);
    long long want = 8ll * (p.pc_in * Kin + Kin + p.pc_p * Kout + Kout);
    if (p.standardization != PSM_MAX_ABS) want += 8ll * 2 * (p.pc_in + p.pc_p);
    for (int l = 0; l <= p.n_dense; ++l)
        if (dims[l] < 1 || dims[l] > kMaxWidth) return handle_fail(h, PSM_ERR_INVALID, "%s: Dense width %d out of range", path, dims[l]);
    for (int l = 0; l < p.n_dense; ++l) want += 4ll * ((long long)dims[l] * dims[l + 1] + dims[l + 1]);
    if (F.remaining() != want)
        return handle_fail(h, PSM_ERR_INVALID, "%s: payload is %lld bytes, the header promises %lld", path, F.remaining(), want);
    std::vector<double> ci, mi, co, mo, a, b, c, d;
    if (!F.rdv(ci, (size_t)(p.pc_in * Kin)) || !F.rdv(mi, (size_t)Kin) || !F.rdv(co, (size_t)(p.pc_p * Kout)) || !F.rdv(mo, (size_t)Kout))
        return handle_fail(h, PSM_ERR_INVALID, "%s: short read", path);
    if (p.standardization != PSM_MAX_ABS && (!F.rdv(a, p.pc_in) || !F.rdv(b, p.pc_in) || !F.rdv(c, p.pc_p) || !F.rdv(d, p.pc_p)))
        return handle_fail(h, PSM_ERR_INVALID, "%s: short read", path);
    std::vector<std::vector<float>> ks(p.n_dense), bs(p.n_dense);
    std::vector<const float*> kp(p.n_dense), bp(p.n_dense);
    for (int l = 0; l < p.n_dense; ++l) {
        if (!F.rdv(ks[l], (size_t)dims[l] * dims[l + 1]) || !F.rdv(bs[l], (size_t)dims[l + 1])) return handle_fail(h, PSM_ERR_INVALID, "%s: short read", path);
        kp[l] = ks[l].data(); bp[l] = bs[l].data();
    }
    p.layer_dims = dims.data();
    p.pca_in_components = ci.data(); p.pca_in_mean = mi.data(); p.pca_out_components = co.data(); p.pca_out_mean = mo.data();
    p.mean_in = a.data(); p.std_in = b.data(); p.mean_out = c.data(); p.std_out = d.data();
    p.dense_kernels = kp.data(); p.dense_biases = bp.data();
    return psm_load_params(h, &p);
}
}  // namespace

extern "C" int psm_load_params_file(psm_handle* h, const char* path) {
    if (!h) return PSM_ERR_INVALID;
    if (!path) return handle_fail(h, PSM_ERR_INVALID, "psm_load_params_file: NULL path");
    try {
        return load_params_file_body(h, path);
    } catch (const std::exception& e) {
        return handle_fail(h, PSM_ERR_INVALID, "psm_load_params_file(%s): %s", path, e.what());
    } catch (...) {
        return handle_fail(h, PSM_ERR_INVALID, "psm_load_params_file(%s): unknown exception", path);
    }
}
