// Handle, parameter/table upload and per-step orchestration behind the C-ABI (include/psm_b200.h).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/psm_b200.h"
#include "psm_kernels.cuh"
#include "psm_plan.h"
#include "psm_internal.h"
#include <cstdarg>

using namespace psm;

extern "C" int psm_set_timings(psm_handle* h, int32_t on);

namespace {
thread_local std::string g_create_error;

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
inline bool env_on(const char* name) { const char* e = getenv(name); return e && e[0] && e[0] != '0'; }
inline long long round_up_ll(long long v, long long m) { return (v + m - 1) / m * m; }
inline long long Bp2_guard(int B) { return (B + 127) / 128 * 128; }
}  // namespace

// NCCL is bound at run time (dlopen) so that single-GPU callers need no NCCL at all and a process that
// already loaded one (e.g. the copy PyTorch bundles) shares it instead of loading a second.
namespace {
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
    bool load() {
        if (lib) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
        if (!lib) { error = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
        bool ok = true;
        auto sym = [&](const char* n) { void* p = dlsym(lib, n); if (!p) { ok = false; error = std::string("missing NCCL symbol ") + n; } return p; };
        GetUniqueId = (decltype(GetUniqueId))sym("ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))sym("ncclCommInitRank");
        CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
        AllReduce = (decltype(AllReduce))sym("ncclAllReduce");
        AllGather = (decltype(AllGather))sym("ncclAllGather");
        Send = (decltype(Send))sym("ncclSend");
        Recv = (decltype(Recv))sym("ncclRecv");
        GroupStart = (decltype(GroupStart))sym("ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))sym("ncclGroupEnd");
        GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
        if (!ok) { lib = nullptr; return false; }
        return true;
    }
};
NcclApi g_nccl;
}  // namespace

struct psm_handle {
    psm_config cfg{};
    cudaStream_t stream = nullptr;
    std::string err;
    bool params_loaded = false, initialised = false;
    std::vector<void*> allocs;

    // ---- parameters -------------------------------------------------------------------------
    int S = 128, C = 1, pc_in = 0, pc_p = 0, pc_in_pad = 0, pc_p_pad = 0, n_dense = 0, standardization = 0;
    double maxs[5] = {1, 1, 1, 1, 1};
    std::vector<int> dims, dims_pad;
    float* d_comp_u = nullptr;        // [pc_in_pad][2*S*S]    input PCA, velocity channels, planar
    float* d_comp_sdf = nullptr;      // [pc_in_pad][S*S]      input PCA, distance channel (init only)
    std::vector<float> in_a, in_b;    // host: x = (z + zc) * a + b
    float *d_in_a = nullptr, *d_in_b = nullptr;
    std::vector<float*> d_W, d_bias;  // [out_pad][in_pad] (K-major), [out_pad]
    float *d_out_s = nullptr, *d_out_m = nullptr;   // de-standardisation r' = r*s + m   [pc_p_pad]
    float* d_comp_out_t = nullptr;    // [S*S*C][pc_p_pad]     output PCA transposed, planar rows (c, ly, lx)
    float* d_pmean = nullptr;         // [S*S*C] planar

    // ---- geometry / tables ----------------------------------------------------------------------
    // Single-GPU handles are the world == 1 special case of a block-row shard: rows [row0,row1) of the
    // global H x W grid are gathered and placed here, rows [row1,row1+ext_rows) arrive from rank+1,
    // blocks [kb0,kb0+B) of the global plan are evaluated here.
    Plan plan;                         // GLOBAL plan
    long long n_cells = 0;             // owned cells (rows of psm_predict)
    long long n_ghost = 0, n_ghost_pix = 0;
    long long G = 0;                   // pixels placed here, (row1-row0)*W
    long long Gg = 0, G_pad = 0;       // pixels gathered here (+ local overlap rows), padded to 4
    int local_ext = 0;
    long long grid_stride = 0;         // floats per grid plane (own + halo rows)
    long long field_stride = 0;        // floats per field plane (own pixels + ghost pixels)
    int H = 0, W = 0;                  // H = own rows (row1-row0); W global
    int Hglob = 0, row0 = 0, row1 = 0, ext_rows = 0, send_rows = 0;
    int B = 0, B_pad = 0, F = 1;       // B = local blocks
    int Bg = 0, kb0 = 0;               // global block count, first global block of this rank
    int rank = 0, world = 1;
    bool have_back = false;
    ncclComm_t comm = nullptr;
    std::vector<long long> cell_send_ptr, cell_recv_ptr, pix_send_ptr, pix_recv_ptr;
    int32_t *d_cell_send_idx = nullptr, *d_pix_send_idx = nullptr;
    float *d_cell_send = nullptr, *d_pix_send = nullptr;
    double* d_means_loc = nullptr;     // world > 1: this rank's means, zero elsewhere (all-reduce source)
    // peer-memory exchange (default for world > 1; NCCL send/recv + all-reduce is the fallback, PSM_COMM=nccl)
    bool p2p = false;
    P2PArgs* d_p2p = nullptr; PeerMail* d_mail = nullptr;
    std::vector<void*> ipc_opened;
    int n_tasks_glob = 0;              // means + shift-line sums of the whole mesh
    DevShiftTerm* d_terms = nullptr; int term_start[3] = {0, 0, 0};
    int32_t *d_fv[3] = {nullptr, nullptr, nullptr}; float* d_fw[3] = {nullptr, nullptr, nullptr};
    int32_t *d_bv[3] = {nullptr, nullptr, nullptr}; float* d_bw[3] = {nullptr, nullptr, nullptr};
    uint8_t* d_gmask = nullptr; uint16_t* d_owner = nullptr;
    // single GPU: the placement folded into the grid->cell gather (back table re-indexed into the predicted blocks)
    bool fuse_place = false; bool field_stale = false; bool no_fused_offsets = false;
    // multi-GPU fused flow: ghost pixels pushed from the owners' blocks into the region behind d_blocks
    bool mgpu_fused = false; long long ghost_base = 0; int32_t* d_pix_send_blk = nullptr;
    std::vector<int32_t> host_cell_send_idx; SendRun* d_runs = nullptr; int n_runs = 0;
    uint2* d_send_words = nullptr; int2* d_send_entries = nullptr;       // send map of the fused flow (P2PFused)
    int filter_radius = 0; float* d_gauss_w = nullptr; float* d_filter_tmp = nullptr;   // optional post-filter (SMC:353-356)
    std::vector<uint8_t> host_mask;   // [Hglob][W] flow mask (sdfunct != 0), host copy
    bool plan_mask_at(int y, int x) const { return host_mask[(size_t)y * W + x] != 0; }
    std::vector<uint8_t> sdf_int;     // U_to_gradP, single GPU: int(sdfunct) per pixel, the index array of GRAD:392 (psm_integrate_gradp)
    int32_t* d_bb[3] = {nullptr, nullptr, nullptr}; uint16_t* d_bo[3] = {nullptr, nullptr, nullptr};
    int32_t *d_by0 = nullptr, *d_bx0 = nullptr;
    CoverEntry *d_rowcov = nullptr, *d_colcov = nullptr; bool fused_extract = false;   // gather writes the block operand itself
    bool keep_grid = false;           // fused path: also store the grid planes every step (PSM_KEEP_GRID=1); else psm_get_stage rebuilds them
    DevTask* d_tasks = nullptr; DevRec* d_rec = nullptr; int n_tasks = 0 /* local */, rounds = 0;
    // strip sums out of the PCA-inverse epilogue (StripRows): static per-row entry lists + the row-partial table
    int32_t *d_sr_rowptr = nullptr, *d_sr_src = nullptr, *d_sr_slot = nullptr; uint32_t* d_sr_w = nullptr; int sr_n_ent = 0;
    float* d_rowpart = nullptr; bool strip_fuse = false;
    float* d_zc = nullptr;            // [B_pad][pc_in_pad]

    // ---- per-step buffers ---------------------------------------------------------------------------
    double* d_cells = nullptr; double* d_out = nullptr; double* d_pprev = nullptr; double* d_uprev = nullptr;
    float2* d_uv = nullptr; float* d_grid = nullptr; float* d_xu = nullptr; float* d_part = nullptr; int splits = 1;
    float* d_xin = nullptr; float* d_act[2] = {nullptr, nullptr}; float* d_r = nullptr; float* d_blocks = nullptr;
    double* d_means = nullptr; double* d_dbuf[2] = {nullptr, nullptr}; int32_t* d_pbuf[2] = {nullptr, nullptr};
    double* d_offsets = nullptr; float* d_coff = nullptr; float* d_field = nullptr;
    Scalars* d_sc = nullptr; Scalars* h_sc = nullptr;   // h_sc: mapped pinned host memory, only .skip is written by the device
    int* d_host_skip = nullptr;
    TcGemm tc_proj{}, tc_inv{}; std::vector<TcGemm> tc_dense; std::vector<int> dense_splits; int tc_splits = 1;
    // projection A operand straight from the grid planes (GridA): x_array is not materialised
    bool grid_a = false; TcGemmGrid tc_proj_grid{}; float* d_part_grid = nullptr; int32_t* d_row_src = nullptr; int grid_a_rows = 0;
    ASeg* d_aseg = nullptr; int32_t* d_aseg_ptr = nullptr; int32_t* d_a_bytes = nullptr;
    unsigned long long* d_layer_trace = nullptr; int trace_steps = 0;  // PSM_TRACE_LAYERS=1: in-kernel timeline of the Dense layers
    bool dense_presplit = false; float* d_xin_lo = nullptr;            // Dense layers read pre-split operands (no converter pass)
    ProjGemm proj_cl{}; bool proj_cluster = false; float* d_proj_part = nullptr; unsigned int* d_proj_cnt = nullptr;   // one-launch projection
    float* d_dpart = nullptr;       // split-K partials of the Dense layers   // tcgen05 path (gemm_mode 0/1)
    bool dense_cluster = true;      // Dense layers: cluster split-K with on-chip reduction (one launch per layer)
    // whole Dense stack in one persistent launch (opt-in, PSM_DENSE_STACK=1: measured equal to the per-layer cluster
    // kernels on the stage and slower for the step): weights and activations pre-split into tf32 hi / lo
    bool inv_t = false; InvT tc_inv_t{}; float *d_r_hi = nullptr, *d_r_lo = nullptr;   // transposed PCA inverse (default)
    bool dense_stack = false; int stack_clusters = 0;
    std::vector<float*> d_Whi, d_Wlo; float* d_act_hi[2] = {nullptr, nullptr}; float* d_act_lo[2] = {nullptr, nullptr};
    TensorMap128* d_stack_maps = nullptr; DenseStackArgs stack_args{};
    int launches = 0;
    cudaEvent_t ev[PSM_N_TIMINGS + 1] = {};
    bool ev_valid = false, ev_created = false;
    bool last_host = false;
    // whole-step CUDA graphs, keyed by the caller's buffers (the solver passes the same ones every step,
    // FOAM/PythonComm_init.H:53)
    struct StepGraph { const void *in = nullptr, *in2 = nullptr, *in3 = nullptr; void* out = nullptr; int stride = 0; cudaGraphExec_t exec = nullptr; unsigned long long used = 0; };
    StepGraph graphs[4];             // small LRU: a solver alternates between at most a few buffer pairs
    unsigned long long graph_clock = 0;
    bool use_graphs = true;
    int eager_steps = 0;
    // host entry points: a second stream copies the caller's p array while the kernels already run (only the last kernel reads it)
    cudaStream_t copy_stream = nullptr; cudaEvent_t ev_u = nullptr, ev_p = nullptr;
    // chunked tail of the host entry points: p goes up, the grid->cell gather runs and the pressures go down chunk by chunk, so the
    // device->host copy of chunk k shares the (full-duplex) link with the host->device copy of p's chunk k+1
    static constexpr int kTailMax = 8;
    cudaStream_t d2h_stream = nullptr; cudaEvent_t ev_pc[kTailMax] = {}, ev_bc[kTailMax] = {}, ev_d2h = nullptr;
    // PSM_SPLIT_H2D=1: every array goes up as two halves on two streams, p's chunks alternate between copy_stream and
    // copy_stream2.  Measured neutral on this platform (profiles/r2h_pcie_probe.txt: 23.5 MB in 0.445 ms on one stream, 0.560 ms
    // on two; e2e 0.916 vs 0.908 ms), so one stream is the default
    cudaStream_t copy_stream2 = nullptr; cudaEvent_t ev_um = nullptr; bool split_h2d = false;
    int tail_chunks = 2;
    bool pprev_zero = false;          // d_pprev currently holds zeros (psm_predict_fields without p)
    size_t cells_capacity = 0;        // doubles in d_cells
    // cell routing (psm_route_init / psm_predict_routed): arbitrary per-rank cell sets -> block-row owners and back
    bool routed = false; long long route_n = 0;
    std::vector<long long> rt_send_ptr, rt_recv_ptr;   // [world + 1], in rows
    int32_t *d_rt_perm = nullptr, *d_rt_recv_idx = nullptr;
    double *d_rt_in = nullptr, *d_rt_send = nullptr, *d_rt_recv = nullptr, *d_rt_osend = nullptr, *d_rt_orecv = nullptr, *d_rt_out = nullptr;
};

// What the first kernel of a step reads (device pointers): the solver's packed rows, or its native field arrays.
struct StepInput {
    const double* rows = nullptr;                                   // double[n][input_cols]
    const double* U = nullptr; const double* dU = nullptr; int u_stride = 0;   // double[n][u_stride] (psm_predict_fields*)
    const double* p_dev = nullptr;                                  // fields: previous pressure read in place (NULL: h->d_pprev)
    bool wait_p = false;                                            // host fields path: the p copy runs on copy_stream
    int tail_chunks = 1;                                            // host paths: chunks of the grid->cell gather (wait_p: h->ev_pc[k] per chunk)
    double* host_out = nullptr;                                     // host paths: chunk k of d_out is copied here on d2h_stream behind its gather
};

#define PSM_FAIL(h, code, ...)                                    \
    do {                                                          \
        char _b[512];                                             \
        snprintf(_b, sizeof _b, __VA_ARGS__);                     \
        (h)->err = _b;                                            \
        return (code);                                            \
    } while (0)

#define CU(h, call)                                                                              \
    do {                                                                                         \
        cudaError_t _e = (call);                                                                 \
        if (_e != cudaSuccess) PSM_FAIL(h, PSM_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(_e)); \
    } while (0)

template <typename T>
static int dalloc(psm_handle* h, T** p, size_t n, bool zero = true) {
    void* q = nullptr;
    size_t bytes = (n ? n : 1) * sizeof(T);
    CU(h, cudaMalloc(&q, bytes));
    h->allocs.push_back(q);
    if (zero) CU(h, cudaMemsetAsync(q, 0, bytes, h->stream));
    *p = static_cast<T*>(q);
    return 0;
}
template <typename T>
static int upload(psm_handle* h, T** p, const std::vector<T>& v) {
    int rc = dalloc(h, p, v.size(), false);
    if (rc) return rc;
    CU(h, cudaMemcpy(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}
#define TRY(x) do { int _rc = (x); if (_rc) return _rc; } while (0)
#define NC(h, call)                                                                                         \
    do {                                                                                                    \
        ncclResult_t _r = (call);                                                                           \
        if (_r != ncclSuccess) PSM_FAIL(h, PSM_ERR_COMM, "%s: %s", #call, g_nccl.GetErrorString(_r));       \
    } while (0)


namespace psm {
int handle_shape(const psm_handle* h) { return h ? h->S : 0; }
int handle_variant(const psm_handle* h) { return h->cfg.variant; }
int handle_device(const psm_handle* h) { return h->cfg.device; }
double handle_delta(const psm_handle* h) { return h->cfg.delta; }
int handle_fail(psm_handle* h, int code, const char* fmt, ...) {
    char b[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(b, sizeof b, fmt, ap);
    va_end(ap);
    if (h) h->err = b; else g_create_error = b;
    return code;
}
}  // namespace psm

extern "C" int psm_api_version(void) { return PSM_API_VERSION; }

extern "C" const char* psm_last_error(const psm_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

static bool make_tail(psm_handle* h) {
    if (cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking) != cudaSuccess) return false;
    if (cudaStreamCreateWithFlags(&h->copy_stream2, cudaStreamNonBlocking) != cudaSuccess) return false;
    if (cudaEventCreateWithFlags(&h->ev_um, cudaEventDisableTiming) != cudaSuccess) return false;
    if (const char* e = getenv("PSM_SPLIT_H2D")) h->split_h2d = (e[0] == '1');
    for (int k = 0; k < psm_handle::kTailMax; ++k)
        if (cudaEventCreateWithFlags(&h->ev_pc[k], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&h->ev_bc[k], cudaEventDisableTiming) != cudaSuccess) return false;
    if (cudaEventCreateWithFlags(&h->ev_d2h, cudaEventDisableTiming) != cudaSuccess) return false;
    if (const char* e = getenv("PSM_TAIL_CHUNKS")) { const int v = atoi(e); if (v >= 1 && v <= psm_handle::kTailMax) h->tail_chunks = v; }
    return true;
}
// cells [c0, c1) of tail chunk k of nch (boundaries on multiples of 4 cells: the 128-bit table loads of the gather stay aligned)
static inline void tail_range(long long n, int nch, int k, long long* c0, long long* c1) {
    const long long per = (((n + nch - 1) / nch) + 3) & ~3ll;
    *c0 = per * k < n ? per * k : n;
    *c1 = (k == nch - 1 || per * (k + 1) > n) ? n : per * (k + 1);
}
// Host->device copy of `n` doubles as two halves on the main stream and copy_stream (two DMA streams in flight); the main stream
// then waits for the second half.  Opt-in (PSM_SPLIT_H2D=1); otherwise, and for small arrays, one copy on the main stream.
static int upload_split(psm_handle* h, double* dst, const double* src, size_t n) {
    if (!h->split_h2d || n < (size_t)(1 << 18)) {
        CU(h, cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        return PSM_OK;
    }
    const size_t half = (n / 2 + 15) & ~(size_t)15;
    CU(h, cudaMemcpyAsync(dst + half, src + half, (n - half) * sizeof(double), cudaMemcpyHostToDevice, h->copy_stream));
    CU(h, cudaMemcpyAsync(dst, src, half * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaEventRecord(h->ev_u, h->copy_stream));
    CU(h, cudaStreamWaitEvent(h->stream, h->ev_u, 0));
    return PSM_OK;
}
static inline int tail_chunks_for(const psm_handle* h) { return h->n_cells >= (1 << 18) ? h->tail_chunks : 1; }

extern "C" int psm_create(psm_handle** out, const psm_config* cfg) {
    if (!out || !cfg) { g_create_error = "psm_create: NULL argument"; return PSM_ERR_INVALID; }
    *out = nullptr;
    if (cfg->variant != PSM_DELTAU_TO_DELTAP && cfg->variant != PSM_U_TO_GRADP && cfg->variant != PSM_THESIS_U_TO_P) { g_create_error = "unknown variant"; return PSM_ERR_INVALID; }
    if (cfg->shape != 128) { g_create_error = "only shape == 128 is supported"; return PSM_ERR_INVALID; }
    if (cfg->gemm_mode < 0 || cfg->gemm_mode > 2) { g_create_error = "unknown gemm_mode"; return PSM_ERR_INVALID; }
    if (!(cfg->filter_sigma >= 0.0) || cfg->filter_sigma > 1000.0) { g_create_error = "filter_sigma must be in [0, 1000]"; return PSM_ERR_INVALID; }
    if (cfg->input_cols != 5 && !(cfg->input_cols == 7 && cfg->variant == PSM_DELTAU_TO_DELTAP)) {
        g_create_error = "input_cols must be 5, or 7 for deltaU_to_deltaP"; return PSM_ERR_INVALID;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no CUDA device (this library has no CPU fallback): ") + cudaGetErrorString(e);
        return PSM_ERR_CUDA;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { g_create_error = "device ordinal out of range"; return PSM_ERR_INVALID; }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, cfg->device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return PSM_ERR_CUDA; }
    if (prop.major != 10) {
        g_create_error = "device is not sm_100 (Blackwell B200): kernels are built for sm_100a only";
        return PSM_ERR_CUDA;
    }
    if ((e = cudaSetDevice(cfg->device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return PSM_ERR_CUDA; }
    psm_handle* h = new (std::nothrow) psm_handle();
    if (!h) { g_create_error = "out of host memory"; return PSM_ERR_INVALID; }
    h->cfg = *cfg;
    h->S = cfg->shape;
    if (const char* e = getenv("PSM_NO_GRAPHS")) h->use_graphs = !(e[0] == '1');
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        g_create_error = cudaGetErrorString(e); delete h; return PSM_ERR_CUDA;
    }
    if (cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&h->ev_u, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_p, cudaEventDisableTiming) != cudaSuccess || !make_tail(h)) {
        g_create_error = "cannot create the copy stream / events"; cudaStreamDestroy(h->stream); delete h; return PSM_ERR_CUDA;
    }
    if (cudaHostAlloc((void**)&h->h_sc, sizeof(Scalars), cudaHostAllocMapped) != cudaSuccess) { g_create_error = "cudaMallocHost failed"; cudaStreamDestroy(h->stream); delete h; return PSM_ERR_CUDA; }
    memset(h->h_sc, 0, sizeof(Scalars));
    {
        void* dp = nullptr;
        if (cudaHostGetDevicePointer(&dp, &h->h_sc->skip, 0) != cudaSuccess) { g_create_error = "cudaHostGetDevicePointer failed"; cudaFreeHost(h->h_sc); cudaStreamDestroy(h->stream); delete h; return PSM_ERR_CUDA; }
        h->d_host_skip = static_cast<int*>(dp);
    }
    *out = h;
    return PSM_OK;
}

extern "C" int psm_destroy(psm_handle* h) {
    if (!h) return PSM_OK;
    cudaSetDevice(h->cfg.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
    if (h->d2h_stream) { cudaStreamSynchronize(h->d2h_stream); cudaStreamDestroy(h->d2h_stream); }
    if (h->copy_stream2) { cudaStreamSynchronize(h->copy_stream2); cudaStreamDestroy(h->copy_stream2); }
    if (h->ev_um) cudaEventDestroy(h->ev_um);
    for (int k = 0; k < psm_handle::kTailMax; ++k) { if (h->ev_pc[k]) cudaEventDestroy(h->ev_pc[k]); if (h->ev_bc[k]) cudaEventDestroy(h->ev_bc[k]); }
    if (h->ev_d2h) cudaEventDestroy(h->ev_d2h);
    if (h->ev_u) cudaEventDestroy(h->ev_u);
    if (h->ev_p) cudaEventDestroy(h->ev_p);
    for (auto& g : h->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    for (void* p : h->ipc_opened) cudaIpcCloseMemHandle(p);
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    for (void* p : h->allocs) cudaFree(p);
    if (h->ev_created) for (auto& e : h->ev) cudaEventDestroy(e);
    if (h->h_sc) cudaFreeHost(h->h_sc);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return PSM_OK;
}

// ------------------------------------------------------------------------------------------------
extern "C" int psm_load_params(psm_handle* h, const psm_params* p) {
    if (!h) return PSM_ERR_INVALID;
    if (!p) PSM_FAIL(h, PSM_ERR_INVALID, "psm_load_params: NULL params");
    if (h->params_loaded) PSM_FAIL(h, PSM_ERR_STATE, "parameters already loaded");
    CU(h, cudaSetDevice(h->cfg.device));
    const int S = h->S, S2 = S * S;
    const int C = p->n_out_channels;
    if (C != (h->cfg.variant == PSM_U_TO_GRADP ? 2 : 1)) PSM_FAIL(h, PSM_ERR_INVALID, "n_out_channels does not match the variant");
    if (p->pc_in < 1 || p->pc_p < 1 || p->n_dense < 1 || p->n_dense > 64) PSM_FAIL(h, PSM_ERR_INVALID, "bad pc_in/pc_p/n_dense");
    if (!p->pca_in_components || !p->pca_in_mean || !p->pca_out_components || !p->pca_out_mean || !p->layer_dims ||
        !p->dense_kernels || !p->dense_biases) PSM_FAIL(h, PSM_ERR_INVALID, "NULL parameter array");
    if (p->layer_dims[0] != p->pc_in || p->layer_dims[p->n_dense] != p->pc_p) PSM_FAIL(h, PSM_ERR_INVALID, "layer_dims must start at pc_in and end at pc_p");
    if (p->standardization != PSM_STD && p->standardization != PSM_MAX_ABS && p->standardization != PSM_MIN_MAX) PSM_FAIL(h, PSM_ERR_INVALID, "unknown standardization code %d", p->standardization);
    if (p->standardization != PSM_MAX_ABS && (!p->mean_in || !p->std_in || !p->mean_out || !p->std_out)) PSM_FAIL(h, PSM_ERR_INVALID, "PSM_STD / PSM_MIN_MAX need the four per-component arrays");
    h->C = C; h->F = C; h->pc_in = p->pc_in; h->pc_p = p->pc_p; h->n_dense = p->n_dense; h->standardization = p->standardization;
    memcpy(h->maxs, p->maxs, sizeof h->maxs);
    // widths are padded to 128: the fused Dense kernel gives each of the 8 CTAs of a cluster N/8 (>= 16) columns
    h->pc_in_pad = round_up(p->pc_in, 128);
    h->pc_p_pad = round_up(p->pc_p, 128);
    h->dims.assign(p->layer_dims, p->layer_dims + p->n_dense + 1);
    h->dims_pad.resize(h->dims.size());
    for (size_t i = 0; i < h->dims.size(); ++i) {
        if (h->dims[i] < 1) PSM_FAIL(h, PSM_ERR_INVALID, "layer width < 1");
        h->dims_pad[i] = round_up(h->dims[i], 128);
    }
    const int Kin = S2 * 3, Kout = S2 * C;

    // input PCA: split the reference's channel-last columns k = p*3 + c into the two velocity planes
    // (dynamic, K = 2*S*S) and the distance plane (static per mesh -> folded into zc at init).
    {
        std::vector<float> cu((size_t)h->pc_in_pad * 2 * S2, 0.f), cs((size_t)h->pc_in_pad * S2, 0.f);
        std::vector<double> mconst(p->pc_in, 0.0);
        for (int n = 0; n < p->pc_in; ++n) {
            const double* row = p->pca_in_components + (size_t)n * Kin;
            double acc = 0.0;
            for (int q = 0; q < S2; ++q) {
                cu[((size_t)n * 2 + 0) * S2 + q] = (float)row[q * 3 + 0];
                cu[((size_t)n * 2 + 1) * S2 + q] = (float)row[q * 3 + 1];
                cs[(size_t)n * S2 + q] = (float)row[q * 3 + 2];
            }
            for (int k = 0; k < Kin; ++k) acc += p->pca_in_mean[k] * row[k];      // mean_ . components_^T (PMP:349)
            mconst[n] = acc;
        }
        TRY(upload(h, &h->d_comp_u, cu));
        TRY(upload(h, &h->d_comp_sdf, cs));
        h->in_a.assign(h->pc_in_pad, 0.f);
        h->in_b.assign(h->pc_in_pad, 0.f);
        for (int n = 0; n < p->pc_in; ++n) {
            double a, m;
            if (p->standardization == PSM_STD) { a = 1.0 / p->std_in[n]; m = p->mean_in[n]; }      // SMC:512
            else if (p->standardization == PSM_MIN_MAX) { a = 1.0 / (p->std_in[n] - p->mean_in[n]); m = p->mean_in[n]; }   // SMC:520 (min in mean_in, max in std_in)
            else { a = 1.0 / p->max_abs_input_PCA; m = 0.0; }                                      // SMC:523
            h->in_a[n] = (float)a;
            h->in_b[n] = (float)(-(mconst[n] + m) * a);
        }
        TRY(upload(h, &h->d_in_a, h->in_a));
        TRY(upload(h, &h->d_in_b, h->in_b));
    }
    // Dense stack: Keras kernels [in][out] -> K-major [out_pad][in_pad], zero padded
    for (int l = 0; l < p->n_dense; ++l) {
        const int in = h->dims[l], out = h->dims[l + 1], in_p = h->dims_pad[l], out_p = h->dims_pad[l + 1];
        if (!p->dense_kernels[l] || !p->dense_biases[l]) PSM_FAIL(h, PSM_ERR_INVALID, "NULL dense layer");
        std::vector<float> w((size_t)out_p * in_p, 0.f), b(out_p, 0.f);
        for (int i = 0; i < in; ++i)
            for (int o = 0; o < out; ++o) w[(size_t)o * in_p + i] = p->dense_kernels[l][(size_t)i * out + o];
        for (int o = 0; o < out; ++o) b[o] = p->dense_biases[l][o];
        float *dw = nullptr, *db = nullptr;
        TRY(upload(h, &dw, w));
        TRY(upload(h, &db, b));
        h->d_W.push_back(dw);
        h->d_bias.push_back(db);
        {   // 3xTF32 operands of the fused Dense stack: hi = the TF32 part, lo = the exact remainder
            std::vector<float> whi(w.size()), wlo(w.size());
            for (size_t i = 0; i < w.size(); ++i) {
                uint32_t u; memcpy(&u, &w[i], 4); u &= 0xFFFFE000u;
                memcpy(&whi[i], &u, 4); wlo[i] = w[i] - whi[i];
            }
            float *dh = nullptr, *dl = nullptr;
            TRY(upload(h, &dh, whi));
            TRY(upload(h, &dl, wlo));
            h->d_Whi.push_back(dh);
            h->d_Wlo.push_back(dl);
        }
    }
    {   // de-standardisation (SMC:533 / 537)
        std::vector<float> s(h->pc_p_pad, 0.f), m(h->pc_p_pad, 0.f);
        for (int n = 0; n < p->pc_p; ++n) {
            if (p->standardization == PSM_STD) { s[n] = (float)p->std_out[n]; m[n] = (float)p->mean_out[n]; }
            else if (p->standardization == PSM_MIN_MAX) { s[n] = (float)(p->std_out[n] - p->mean_out[n]); m[n] = (float)p->mean_out[n]; }   // SMC:536
            else { s[n] = (float)p->max_abs_output_PCA; m[n] = 0.f; }
        }
        TRY(upload(h, &h->d_out_s, s));
        TRY(upload(h, &h->d_out_m, m));
    }
    {   // output PCA transposed to K-major rows in planar pixel order q' = c*S2 + q  (reference k = q*C + c)
        std::vector<float> ct((size_t)Kout * h->pc_p_pad, 0.f), pm(Kout, 0.f);
        for (int n = 0; n < p->pc_p; ++n) {
            const double* row = p->pca_out_components + (size_t)n * Kout;
            for (int q = 0; q < S2; ++q)
                for (int c = 0; c < C; ++c) ct[((size_t)c * S2 + q) * h->pc_p_pad + n] = (float)row[q * C + c];
        }
        for (int q = 0; q < S2; ++q)
            for (int c = 0; c < C; ++c) pm[(size_t)c * S2 + q] = (float)p->pca_out_mean[q * C + c];
        TRY(upload(h, &h->d_comp_out_t, ct));
        TRY(upload(h, &h->d_pmean, pm));
    }
    h->params_loaded = true;
    return PSM_OK;
}

// ------------------------------------------------------------------------------------------------
// Everything init needs about this rank's share, tables already folded and in local numbering.
namespace {
struct LocalInit {
    int rank = 0, world = 1;
    int H = 0, W = 0;                          // GLOBAL grid
    int row0 = 0, row1 = 0, ext_rows = 0, send_rows = 0, blk_row0 = 0, blk_row1 = 0, local_ext = 0;
    const uint8_t* mask_global = nullptr;      // [H][W]
    long long n_owned = 0, n_ghost = 0, n_ghost_pix = 0;
    std::vector<int32_t> fv[3]; std::vector<float> fw[3];      // [(row1-row0+local_ext)*W]
    const double* sdf_rows = nullptr;          // [row1-row0+local_ext+ext_rows][W]
    bool have_back = false;
    std::vector<int32_t> bv[3]; std::vector<float> bw[3];      // [n_owned]
    std::vector<long long> cell_send_ptr, cell_recv_ptr, pix_send_ptr, pix_recv_ptr;
    std::vector<int32_t> cell_send_idx, pix_send_idx;
    std::vector<long long> ghost_pix;          // global pixel id per ghost slot (empty: unknown -> legacy multi-GPU flow)
};
}  // namespace

// Box plan of the grid-plane A operand (GridA): blocks whose origins are an arithmetic progression of `st` pixels along x (same
// y0) are fetched gx at a time, the blocks left over (the clamped column, SMC:467-472 / GRAD:486-494) gy at a time along y, the
// rest one by one; boxes are packed into 128-row tiles.  Returns false when the layout does not pay (no run of >= 4 blocks).
namespace {
struct GridAPlan { std::vector<ASeg> segs; std::vector<int32_t> seg_ptr, a_bytes, row_src; int tiles = 0, gx = 1, gy = 1; };
static int best_box(int n) { for (int g = std::min(n, 16); g >= 2; --g) if (n % g == 0) return g; return 1; }
static bool build_grid_a(int B, const std::vector<int32_t>& by0, const std::vector<int32_t>& bx0, int st, GridAPlan& P) {
    struct Item { int map, x, y, len; std::vector<int> blk; };
    std::vector<Item> items;
    std::vector<int> order(B);
    for (int b = 0; b < B; ++b) order[b] = b;
    // chains along x
    std::sort(order.begin(), order.end(), [&](int a, int b) { return by0[a] != by0[b] ? by0[a] < by0[b] : bx0[a] < bx0[b]; });
    std::vector<std::vector<int>> chains;
    for (int i = 0; i < B; ++i) {
        const int b = order[i];
        if (!chains.empty()) { const int p = chains.back().back(); if (by0[p] == by0[b] && bx0[b] - bx0[p] == st) { chains.back().push_back(b); continue; } }
        chains.push_back({b});
    }
    size_t nreg = 0;
    for (auto& c : chains) nreg = std::max(nreg, c.size());
    if (nreg < 4) return false;
    P.gx = best_box((int)nreg);
    if (P.gx < 4) return false;
    std::vector<int> left;
    for (auto& c : chains) {
        size_t i = 0;
        for (; i + P.gx <= c.size(); i += P.gx) items.push_back(Item{0, bx0[c[i]], by0[c[i]], P.gx, std::vector<int>(c.begin() + i, c.begin() + i + P.gx)});
        for (; i < c.size(); ++i) left.push_back(c[i]);
    }
    // chains along y among the blocks left over
    std::sort(left.begin(), left.end(), [&](int a, int b) { return bx0[a] != bx0[b] ? bx0[a] < bx0[b] : by0[a] < by0[b]; });
    std::vector<std::vector<int>> cols;
    for (int b : left) {
        if (!cols.empty()) { const int p = cols.back().back(); if (bx0[p] == bx0[b] && by0[b] - by0[p] == st) { cols.back().push_back(b); continue; } }
        cols.push_back({b});
    }
    size_t ncol = 0;
    for (auto& c : cols) ncol = std::max(ncol, c.size());
    P.gy = ncol >= 2 ? best_box((int)ncol) : 1;
    std::vector<int> singles;
    for (auto& c : cols) {
        size_t i = 0;
        if (P.gy >= 2)
            for (; i + P.gy <= c.size(); i += P.gy) items.push_back(Item{1, bx0[c[i]], by0[c[i]], P.gy, std::vector<int>(c.begin() + i, c.begin() + i + P.gy)});
        for (; i < c.size(); ++i) singles.push_back(c[i]);
    }
    for (int b : singles) items.push_back(Item{2, bx0[b], by0[b], 1, {b}});
    // first fit into 128-row tiles (boxes arrive largest first, singles last: they fill the gaps)
    std::vector<int> used;
    std::vector<std::vector<ASeg>> per_tile;
    P.row_src.assign(B, -1);
    for (const Item& it : items) {
        int t = -1;
        for (size_t k = 0; k < used.size(); ++k) if (used[k] + it.len <= 128) { t = (int)k; break; }
        if (t < 0) { used.push_back(0); per_tile.emplace_back(); t = (int)used.size() - 1; }
        per_tile[t].push_back(ASeg{it.map, used[t], it.x, it.y});
        for (int i = 0; i < it.len; ++i) P.row_src[it.blk[i]] = t * 128 + used[t] + i;
        used[t] += it.len;
    }
    P.tiles = (int)used.size();
    P.seg_ptr.assign(1, 0);
    for (int t = 0; t < P.tiles; ++t) {
        for (const ASeg& sg : per_tile[t]) P.segs.push_back(sg);
        P.seg_ptr.push_back((int32_t)P.segs.size());
        P.a_bytes.push_back(used[t] * 128);
    }
    // self-check: every block is fetched exactly once, from its own origin
    std::vector<int> seen(B, 0);
    for (int t = 0; t < P.tiles; ++t)
        for (int q = P.seg_ptr[t]; q < P.seg_ptr[t + 1]; ++q) {
            const ASeg& sg = P.segs[q];
            const int len = sg.map == 0 ? P.gx : (sg.map == 1 ? P.gy : 1);
            for (int i = 0; i < len; ++i) {
                const int row = t * 128 + sg.row + i;
                int b = -1;
                for (int k = 0; k < B; ++k) if (P.row_src[k] == row) { b = k; break; }
                if (b < 0) return false;
                const int x = sg.x + (sg.map == 0 ? i * st : 0), y = sg.y + (sg.map == 1 ? i * st : 0);
                if (bx0[b] != x || by0[b] != y) return false;
                ++seen[b];
            }
        }
    for (int b = 0; b < B; ++b) if (seen[b] != 1) return false;
    return true;
}
}  // namespace

// Send map of the fused multi-GPU flow (P2PFused::send_words / send_entries): see psm_kernels.cuh.  send_idx[send_ptr[p] ..
// send_ptr[p+1]) are the owned cells peer p needs, entry e of that list lands in slot e - send_ptr[p] of p's ghost region.
static void build_send_map(long long n_cells, int world, const long long* send_ptr, const int32_t* send_idx,
                           std::vector<uint2>& words, std::vector<int2>& entries) {
    struct Ent { int32_t cell, peer; long long dst; };
    std::vector<Ent> ent;
    for (int p = 0; p < world; ++p)
        for (long long e = send_ptr[p]; e < send_ptr[p + 1]; ++e) ent.push_back(Ent{send_idx[e], p, e - send_ptr[p]});
    std::stable_sort(ent.begin(), ent.end(), [](const Ent& a, const Ent& b) { return a.cell < b.cell; });
    const size_t nw = (size_t)(n_cells + 31) / 32 + 1;
    words.assign(nw, make_uint2(0u, 0u));
    // one main entry per marked cell (index = rank of the cell among the marked ones: bitmap prefix + popcount); a cell
    // that goes to several peers chains its further entries behind the main ones (next index in the bits above 9)
    size_t n_marked = 0;
    for (size_t k = 0; k < ent.size(); ++k) if (k == 0 || ent[k].cell != ent[k - 1].cell) ++n_marked;
    entries.assign(ent.size() ? ent.size() : 1, make_int2(0, 0));
    size_t main_i = 0, over_i = n_marked;
    for (size_t k = 0; k < ent.size();) {
        size_t k1 = k;
        while (k1 < ent.size() && ent[k1].cell == ent[k].cell) ++k1;
        words[(size_t)ent[k].cell >> 5].x |= 1u << (ent[k].cell & 31);
        size_t at = main_i++;
        for (size_t q = k; q < k1; ++q) {
            const bool more = q + 1 < k1;
            const size_t nxt = more ? over_i++ : 0;
            entries[at] = make_int2(ent[q].peer | (more ? 0x100 : 0) | (int)(nxt << 9), (int)ent[q].dst);
            at = nxt;
        }
        k = k1;
    }
    unsigned int run = 0;
    for (size_t w = 0; w < nw; ++w) { words[w].y = run; run += (unsigned int)__builtin_popcount(words[w].x); }
}

// Map every peer's exchange buffers (cudaIpc over NVLink) and build the push tables.  Collective.  Falls back
// to the NCCL exchange when any rank cannot map a peer, when PSM_COMM=nccl, or in the grid-row halo mode.
static int setup_p2p(psm_handle* h) {
    struct PeerInfo {
        cudaIpcMemHandle_t uv, field, means, mail, blocks;
        long long n_owned, G, field_stride, ghost_base;
        int fused;
        long long cell_recv_ptr[kMaxPeers + 1], pix_recv_ptr[kMaxPeers + 1];
        int ok;
    };
    const int Wd = h->world, me = h->rank;
    const char* mode = getenv("PSM_COMM");
    int want = (Wd <= kMaxPeers) && h->ext_rows == 0 && h->send_rows == 0 && !(mode && std::string(mode) == "nccl");
    TRY(dalloc(h, &h->d_mail, 1));
    PeerInfo mine{};
    mine.ok = want;
    if (want) {
        if (cudaIpcGetMemHandle(&mine.uv, h->d_uv) != cudaSuccess || cudaIpcGetMemHandle(&mine.field, h->d_field) != cudaSuccess ||
            cudaIpcGetMemHandle(&mine.means, h->d_means) != cudaSuccess || cudaIpcGetMemHandle(&mine.mail, h->d_mail) != cudaSuccess ||
            cudaIpcGetMemHandle(&mine.blocks, h->d_blocks) != cudaSuccess) {
            mine.ok = 0; cudaGetLastError();
        }
    }
    mine.n_owned = h->n_cells; mine.G = h->G; mine.field_stride = h->field_stride; mine.ghost_base = h->ghost_base;
    // the fused flow needs the from-blocks back table, the strip sums out of the PCA-inverse epilogue and offsets that fit one CTA
    mine.fused = (h->fuse_place && h->strip_fuse && h->inv_t && h->have_back && (long long)h->Bg * h->F <= 4096 && !h->no_fused_offsets) ? 1 : 0;
    for (int p = 0; p <= Wd && p <= kMaxPeers; ++p) { mine.cell_recv_ptr[p] = h->cell_recv_ptr[p]; mine.pix_recv_ptr[p] = h->pix_recv_ptr[p]; }
    // all-gather the descriptors (bytes) through NCCL
    PeerInfo* d_all = nullptr;
    CU(h, cudaMalloc(&d_all, sizeof(PeerInfo) * Wd));
    CU(h, cudaMemcpyAsync(d_all + me, &mine, sizeof mine, cudaMemcpyHostToDevice, h->stream));
    NC(h, g_nccl.AllGather(d_all + me, d_all, sizeof(PeerInfo), ncclChar, h->comm, h->stream));
    std::vector<PeerInfo> all(Wd);
    CU(h, cudaMemcpyAsync(all.data(), d_all, sizeof(PeerInfo) * Wd, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    bool ok = true;
    for (int p = 0; p < Wd; ++p) ok = ok && all[p].ok;
    P2PArgs pa{};
    pa.rank = me; pa.world = Wd; pa.sc = h->d_sc;
    if (ok) {
        for (int p = 0; p < Wd && ok; ++p) {
            void *uv = h->d_uv, *field = h->d_field, *means = h->d_means, *mail = h->d_mail, *blocks = h->d_blocks;
            if (p != me) {
                auto open = [&](void** out, cudaIpcMemHandle_t hd) {
                    if (cudaIpcOpenMemHandle(out, hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return false; }
                    h->ipc_opened.push_back(*out);
                    return true;
                };
                ok = open(&uv, all[p].uv) && open(&field, all[p].field) && open(&means, all[p].means) && open(&mail, all[p].mail) &&
                     open(&blocks, all[p].blocks);
                if (!ok) break;
            }
            pa.mail[p] = static_cast<PeerMail*>(mail);
            pa.means[p] = static_cast<double*>(means);
            pa.uv_ghost[p] = static_cast<float2*>(uv) + all[p].n_owned + all[p].cell_recv_ptr[me];
            pa.field_ghost[p] = static_cast<float*>(field) + all[p].G + all[p].pix_recv_ptr[me];
            pa.field_stride[p] = all[p].field_stride;
            pa.blk_ghost[p] = static_cast<float*>(blocks) + all[p].ghost_base;
            pa.pix_slot0[p] = all[p].pix_recv_ptr[me];
        }
    }
    bool fused = true;
    for (int p = 0; p < Wd; ++p) fused = fused && all[p].fused;
    // every rank must take the same path: agree through one more tiny all-gather
    int* d_flag = nullptr;
    CU(h, cudaMalloc(&d_flag, sizeof(int) * Wd));
    int okv = ok ? 1 : 0;
    CU(h, cudaMemcpyAsync(d_flag + me, &okv, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    NC(h, g_nccl.AllGather(d_flag + me, d_flag, sizeof(int), ncclChar, h->comm, h->stream));
    std::vector<int> flags(Wd);
    CU(h, cudaMemcpyAsync(flags.data(), d_flag, sizeof(int) * Wd, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    cudaFree(d_all); cudaFree(d_flag);
    for (int p = 0; p < Wd; ++p) ok = ok && flags[p];
    h->p2p = ok;
    h->mgpu_fused = ok && fused;
    if (h->mgpu_fused) {
        // the cell send list as runs of consecutive cells (see SendRun): pushed by the prep CTAs themselves when there are few
        std::vector<SendRun> runs;
        for (int p = 0; p < Wd; ++p)
            for (long long e = h->cell_send_ptr[p]; e < h->cell_send_ptr[p + 1]; ++e) {
                const int32_t c = h->host_cell_send_idx[e];
                if (!runs.empty() && runs.back().peer == p && runs.back().end == c) { ++runs.back().end; continue; }
                runs.push_back(SendRun{c, c + 1, p, 0, e - h->cell_send_ptr[p]});
            }
        h->n_runs = (runs.size() <= (size_t)kMaxRuns && !env_on("PSM_NO_SEND_RUNS")) ? (int)runs.size() : 0;
        if (runs.empty()) runs.push_back(SendRun{0, 0, 0, 0, 0});
        TRY(upload(h, &h->d_runs, runs));
        // The send map: works for any cell numbering and any list length (a solver's cells are not numbered row by row, and even on
        // the synthetic lattice a 4000-wide boundary is hundreds of runs: with the run list the LAST prep CTA then moved 130 k ghost
        // cells alone, 180 us of the 470 us step at 4 GPUs on c4).  PSM_SEND_RUNS=1 keeps the run list.
        if (!env_on("PSM_SEND_RUNS") && h->cell_send_ptr[Wd] < (1ll << 22)) {     // entry indices are packed into 22 bits
            std::vector<uint2> words; std::vector<int2> entries;
            build_send_map(h->n_cells, Wd, h->cell_send_ptr.data(), h->host_cell_send_idx.data(), words, entries);
            TRY(upload(h, &h->d_send_words, words));
            TRY(upload(h, &h->d_send_entries, entries));
            h->n_runs = 0;
        }
    }
    if (!h->mgpu_fused && h->fuse_place) h->fuse_place = false;      // legacy flow: the assembled field + ghost pixels of the field
    if (!ok) return PSM_OK;                       // NCCL exchange
    for (int p = 0; p <= Wd; ++p) { pa.cell_send_ptr[p] = h->cell_send_ptr[p]; pa.pix_send_ptr[p] = h->pix_send_ptr[p]; }
    for (int p = Wd + 1; p <= kMaxPeers; ++p) { pa.cell_send_ptr[p] = pa.cell_send_ptr[Wd]; pa.pix_send_ptr[p] = pa.pix_send_ptr[Wd]; }
    pa.pix_recv_mask = 0;
    for (int p = 0; p < Wd; ++p) if (h->pix_recv_ptr[p + 1] > h->pix_recv_ptr[p]) pa.pix_recv_mask |= 1u << p;
    std::vector<P2PArgs> v(1, pa);
    TRY(upload(h, &h->d_p2p, v));
    h->eager_steps = 0;                           // nothing lazy left on the step's path: capture from the first step
    return PSM_OK;
}

static int init_local(psm_handle* h, LocalInit& L) {
    const int W = L.W, S = h->S, S2 = S * S;
    h->Hglob = L.H; h->W = W; h->row0 = L.row0; h->row1 = L.row1; h->ext_rows = L.ext_rows; h->send_rows = L.send_rows;
    h->H = L.row1 - L.row0;
    h->rank = L.rank; h->world = L.world;
    h->n_cells = L.n_owned; h->n_ghost = L.n_ghost; h->n_ghost_pix = L.n_ghost_pix;
    const long long N = L.n_owned, Nall = L.n_owned + L.n_ghost;
    const long long G = (long long)h->H * W;                        // pixels placed here
    const long long Gg = (long long)(h->H + L.local_ext) * W;       // pixels gathered here
    const int Hl = h->H + L.local_ext + L.ext_rows;                 // rows the local blocks are cut from
    h->local_ext = L.local_ext;
    h->G = G; h->Gg = Gg; h->G_pad = round_up_ll(Gg, 4);
    h->grid_stride = round_up_ll(std::max(h->G_pad, (long long)Hl * W), 4);
    h->field_stride = round_up_ll(G + L.n_ghost_pix, 4);
    h->have_back = L.have_back;

    // ---- global plan, then this rank's slice ----------------------------------------------------------
    int rc = compile_plan(h->cfg.variant, L.H, W, S, h->cfg.overlap, L.mask_global, h->plan);
    if (rc) PSM_FAIL(h, rc, "%s", h->plan.error.c_str());
    const Plan& P = h->plan;
    const int ncolb = P.ncolb;
    if (L.blk_row1 < 0) L.blk_row1 = P.n_y + 2;
    if (L.blk_row0 < 0 || L.blk_row1 > P.n_y + 2 || L.blk_row0 >= L.blk_row1) PSM_FAIL(h, PSM_ERR_INVALID, "bad block-row range [%d,%d)", L.blk_row0, L.blk_row1);
    h->Bg = P.B; h->kb0 = L.blk_row0 * ncolb; h->B = (L.blk_row1 - L.blk_row0) * ncolb;
    h->B_pad = round_up(h->B, 128); h->F = P.F;
    const int kb0 = h->kb0, kb1 = h->kb0 + h->B;
    h->rounds = 0;
    while ((1 << h->rounds) < P.max_depth + 1) ++h->rounds;
    for (int k = kb0; k < kb1; ++k)
        if (P.y0[k] < L.row0 || P.y0[k] + S > L.row1 + L.local_ext + L.ext_rows)
            PSM_FAIL(h, PSM_ERR_GEOMETRY, "block %d (rows %d..%d) is outside this rank's rows [%d,%d)+%d", k, P.y0[k], P.y0[k] + S, L.row0, L.row1, L.local_ext + L.ext_rows);

    for (int j = 0; j < 3; ++j) {
        L.fv[j].resize(h->G_pad, 0); L.fw[j].resize(h->G_pad, 0.f);
        for (long long q = 0; q < Gg; ++q)
            if (L.fv[j][q] < 0 || L.fv[j][q] >= Nall) PSM_FAIL(h, PSM_ERR_INVALID, "forward table: cell id out of range at pixel %lld", q);
        TRY(upload(h, &h->d_fv[j], L.fv[j])); TRY(upload(h, &h->d_fw[j], L.fw[j]));
    }
    if (L.have_back)
        for (int j = 0; j < 3; ++j) {
            for (long long c = 0; c < N; ++c)
                if (L.bv[j][c] >= G + L.n_ghost_pix || (L.bv[j][c] < 0 && j > 0)) PSM_FAIL(h, PSM_ERR_INVALID, "back table: pixel id out of range at cell %lld", c);
            TRY(upload(h, &h->d_bv[j], L.bv[j])); TRY(upload(h, &h->d_bw[j], L.bw[j]));
        }

    // ---- plan to device -----------------------------------------------------------------------------
    {
        std::vector<uint8_t> mask(L.mask_global, L.mask_global + (size_t)L.H * W);
        TRY(upload(h, &h->d_gmask, mask));
        h->host_mask = mask;
        std::vector<uint16_t> ow(G);
        for (long long q = 0; q < G; ++q) {
            const int o = P.owner[(size_t)L.row0 * W + q];
            if (o < kb0 || o >= kb1) PSM_FAIL(h, PSM_ERR_GEOMETRY, "pixel row %lld is placed by block %d, outside this rank's blocks [%d,%d)", L.row0 + q / W, o, kb0, kb1);
            ow[q] = (uint16_t)(o - kb0);
        }
        TRY(upload(h, &h->d_owner, ow));
        if (h->cfg.filter_sigma > 0.0) {
            // scipy.ndimage.gaussian_filter1d: radius = int(truncate * sigma + 0.5), truncate = 4; weights normalised in FP64
            if (L.world > 1) PSM_FAIL(h, PSM_ERR_INVALID, "the Gaussian post-filter is not available on a sharded handle");
            const double sd = h->cfg.filter_sigma;
            const int lw = (int)(4.0 * sd + 0.5);
            std::vector<double> wd(2 * lw + 1);
            double sum = 0.0;
            for (int i = -lw; i <= lw; ++i) { wd[i + lw] = std::exp(-0.5 / (sd * sd) * (double)i * (double)i); sum += wd[i + lw]; }
            std::vector<float> wf(2 * lw + 1);
            for (int i = 0; i <= 2 * lw; ++i) wf[i] = (float)(wd[i] / sum);
            h->filter_radius = lw;
            TRY(upload(h, &h->d_gauss_w, wf));
            TRY(dalloc(h, &h->d_filter_tmp, (size_t)G));
        }
        // single GPU, or a shard that knows the global pixel of every ghost slot (fused multi-GPU flow)
        const bool ghosts_known = L.n_ghost_pix == 0 || (long long)L.ghost_pix.size() == L.n_ghost_pix;
        h->fuse_place = (L.world == 1 || (ghosts_known && L.ext_rows == 0 && L.send_rows == 0 && !env_on("PSM_MGPU_LEGACY"))) && L.have_back &&
                        !env_on("PSM_NO_FUSED_PLACE") && h->filter_radius == 0 && P.B < 65536;
        h->no_fused_offsets = env_on("PSM_NO_FUSED_OFFSETS");
        h->ghost_base = (long long)h->B_pad * h->C * S2;
        if (h->fuse_place) {
            // pixel -> (last-writer block, offset inside the blocks array) is static: re-index the back table.  Ghost pixels live
            // behind the local blocks as [chunk][C][S*S]; the owner id is GLOBAL (it indexes the global offset array).
            for (long long q : L.ghost_pix) if (q < 0 || q >= (long long)L.H * W) PSM_FAIL(h, PSM_ERR_INVALID, "ghost_pix out of range");
            for (int j = 0; j < 3; ++j) {
                std::vector<int32_t> bb(N); std::vector<uint16_t> bo(N);
                for (long long c = 0; c < N; ++c) {
                    const long long q = L.bv[j][c];
                    if (q < 0) { bb[c] = -1; bo[c] = 0; continue; }          // keep p_prev marker (vertex 0 only)
                    if (q >= G) {                                            // ghost pixel: pushed by its owner
                        const long long sl = q - G;
                        bb[c] = (int32_t)(h->ghost_base + ((sl / S2) * h->C) * S2 + sl % S2);
                        bo[c] = (uint16_t)P.owner[(size_t)L.ghost_pix[sl]];
                        continue;
                    }
                    const int o = ow[q];
                    const int y = (int)(q / W), x = (int)(q - (long long)y * W);
                    bb[c] = (int32_t)(((long long)o * h->C) * S2 + (long long)(y - (P.y0[kb0 + o] - L.row0)) * S + (x - P.x0[kb0 + o]));
                    bo[c] = (uint16_t)(kb0 + o);
                }
                TRY(upload(h, &h->d_bb[j], bb)); TRY(upload(h, &h->d_bo[j], bo));
            }
            if (L.world > 1) {      // where MY pixels that are ghost pixels elsewhere sit in my blocks (channel 0)
                std::vector<int32_t> sb(L.pix_send_idx.size());
                for (size_t e = 0; e < sb.size(); ++e) {
                    const long long q = L.pix_send_idx[e];
                    if (q < 0 || q >= G) PSM_FAIL(h, PSM_ERR_INVALID, "pix_send_idx out of range");
                    const int o = ow[q];
                    const int y = (int)(q / W), x = (int)(q - (long long)y * W);
                    sb[e] = (int32_t)(((long long)o * h->C) * S2 + (long long)(y - (P.y0[kb0 + o] - L.row0)) * S + (x - P.x0[kb0 + o]));
                }
                if (sb.empty()) sb.push_back(0);
                TRY(upload(h, &h->d_pix_send_blk, sb));
            }
        }
        std::vector<int32_t> by0(h->B_pad, 0), bx0(h->B_pad, 0);
        for (int k = 0; k < h->B; ++k) { by0[k] = P.y0[kb0 + k] - L.row0; bx0[k] = P.x0[kb0 + k]; }
        TRY(upload(h, &h->d_by0, by0));
        TRY(upload(h, &h->d_bx0, bx0));
        {   // which block rows / block columns cover a pixel row / a 4-pixel column group (fused gather + extraction)
            const int nrows_g = h->H + L.local_ext, nbr = L.blk_row1 - L.blk_row0;
            bool ok = (W % 4 == 0) && L.ext_rows == 0 && !env_on("PSM_NO_FUSED_EXTRACT") && (long long)Bp2_guard(h->B) * 2 * S2 < (1ll << 31);
            for (int j = 0; j < ncolb && ok; ++j) ok = (bx0[j] % 4 == 0);
            std::vector<CoverEntry> rcv(nrows_g), ccv(W / 4 + 1);
            for (int y = 0; y < nrows_g && ok; ++y) {
                CoverEntry ce{}; 
                for (int r = 0; r < nbr; ++r) {
                    const int y0 = by0[r * ncolb];
                    if (y >= y0 && y < y0 + S) { if (ce.n >= 7) { ok = false; break; } ce.off[ce.n++] = (int32_t)((long long)r * ncolb * 2 * S2 + (long long)(y - y0) * S); }
                }
                rcv[y] = ce;
            }
            for (int xg = 0; xg < W / 4 && ok; ++xg) {
                CoverEntry ce{};
                for (int j = 0; j < ncolb; ++j)
                    if (xg * 4 >= bx0[j] && xg * 4 < bx0[j] + S) { if (ce.n >= 7) { ok = false; break; } ce.off[ce.n++] = (int32_t)((long long)j * 2 * S2 + (xg * 4 - bx0[j])); }
                ccv[xg] = ce;
            }
            h->fused_extract = ok;
            h->keep_grid = env_on("PSM_KEEP_GRID") || L.send_rows > 0;     // grid-row halo: rank-1 reads this rank's first rows
            if (ok) { TRY(upload(h, &h->d_rowcov, rcv)); TRY(upload(h, &h->d_colcov, ccv)); }
        }
        // tasks: masked means first, then the shift-line sums; a task is evaluated by the rank holding `src`
        const int n_means = (int)P.tasks.size();
        int n_lines = 0;
        for (int f = 0; f < P.F; ++f) n_lines += (int)P.lines[f].size();
        h->n_tasks_glob = n_means + n_lines;
        std::vector<DevTask> tk;
        for (int i = 0; i < n_means; ++i) {
            const Task& s = P.tasks[i];
            if (s.src < kb0 || s.src >= kb1) continue;
            if (s.msk >= 0) tk.push_back(DevTask{s.src - kb0, 0, s.ch, s.y0, s.y1, s.x0, s.x1, s.count, P.y0[s.msk], P.x0[s.msk], i, 0});
            else tk.push_back(DevTask{s.src - kb0, 2, s.ch, s.y0, s.y1, s.x0, s.x1, s.count, 0, 0, i, 0});      // plain mean (PMP:437)
        }
        std::vector<DevShiftTerm> terms;
        int slot = n_means;
        for (int f = 0; f < P.F; ++f) {
            h->term_start[f] = (int)terms.size();
            for (const LineTask& s : P.lines[f]) {
                terms.push_back(DevShiftTerm{slot, s.src, s.coef, s.n});
                if (s.src >= kb0 && s.src < kb1)
                    tk.push_back(DevTask{s.src - kb0, 1, s.ch, s.y0, s.y1, s.x0, s.x1, s.n, 0, 0, slot, 0});
                ++slot;
            }
            h->term_start[f + 1] = (int)terms.size();
        }
        h->n_tasks = (int)tk.size();
        {   // row partials of every task: slot = part_base + (row - y0); one entry per (task, block row it covers), grouped by the
            // CTA of pca_inverse_t_kernel that holds that pixel row (r = channel * S + local row), sorted by source block
            int n_slots = 0;
            for (DevTask& t : tk) { t.part_base = n_slots; n_slots += t.y1 - t.y0; }
            struct Ent { int32_t src, slot; uint32_t w[4]; };
            std::vector<std::vector<Ent>> rows((size_t)h->C * S);
            for (const DevTask& t : tk)
                for (int ly = t.y0; ly < t.y1; ++ly) {
                    Ent e{t.src, t.part_base + (ly - t.y0), {0u, 0u, 0u, 0u}};
                    const uint8_t* mrow = (t.kind == 0) ? L.mask_global + (size_t)(t.my0 + ly) * W + t.mx0 : nullptr;
                    for (int lx = t.x0; lx < t.x1; ++lx)
                        if (!mrow || mrow[lx]) e.w[lx >> 5] |= 1u << (lx & 31);
                    rows[(size_t)t.ch * S + ly].push_back(e);
                }
            const int n_c32 = (h->B + 31) / 32;
            std::vector<int32_t> rp((size_t)h->C * S * (n_c32 + 1), 0), es, el;
            size_t n_ent = 0;
            for (auto& r : rows) n_ent += r.size();
            std::vector<uint32_t> ew(4 * (n_ent ? n_ent : 1), 0u);
            es.reserve(n_ent); el.reserve(n_ent);
            for (size_t r = 0; r < rows.size(); ++r) {
                std::stable_sort(rows[r].begin(), rows[r].end(), [](const Ent& a, const Ent& b) { return a.src < b.src; });
                int32_t* rpr = rp.data() + r * (n_c32 + 1);
                int chunk = 0;
                rpr[0] = (int32_t)es.size();
                for (const Ent& e : rows[r]) {
                    while (chunk < e.src / 32) rpr[++chunk] = (int32_t)es.size();
                    for (int q = 0; q < 4; ++q) ew[(size_t)q * n_ent + es.size()] = e.w[q];
                    es.push_back(e.src); el.push_back(e.slot);
                }
                while (chunk < n_c32) rpr[++chunk] = (int32_t)es.size();
            }
            h->sr_n_ent = (int)n_ent;
            if (es.empty()) { es.push_back(0); el.push_back(0); }
            TRY(upload(h, &h->d_sr_rowptr, rp)); TRY(upload(h, &h->d_sr_src, es)); TRY(upload(h, &h->d_sr_slot, el)); TRY(upload(h, &h->d_sr_w, ew));
            TRY(dalloc(h, &h->d_rowpart, (size_t)(n_slots ? n_slots : 1) * 4));
        }
        TRY(upload(h, &h->d_tasks, tk));
        TRY(upload(h, &h->d_terms, terms));
        std::vector<DevRec> rc2(P.rec.size());
        for (size_t i = 0; i < rc2.size(); ++i) rc2[i] = DevRec{P.rec[i].ta, P.rec[i].tb, P.rec[i].parent, P.rec[i].is_nan};
        TRY(upload(h, &h->d_rec, rc2));
    }
    // ---- exchange lists (world > 1) -------------------------------------------------------------------
    if (L.world > 1) {
        h->cell_send_ptr = L.cell_send_ptr; h->cell_recv_ptr = L.cell_recv_ptr;
        h->pix_send_ptr = L.pix_send_ptr; h->pix_recv_ptr = L.pix_recv_ptr;
        for (int32_t v : L.cell_send_idx) if (v < 0 || v >= N) PSM_FAIL(h, PSM_ERR_INVALID, "cell_send_idx out of range");
        for (int32_t v : L.pix_send_idx) if (v < 0 || v >= G) PSM_FAIL(h, PSM_ERR_INVALID, "pix_send_idx out of range");
        if (h->cell_recv_ptr[L.world] != L.n_ghost || h->pix_recv_ptr[L.world] != L.n_ghost_pix) PSM_FAIL(h, PSM_ERR_INVALID, "recv lists do not match the ghost counts");
        TRY(upload(h, &h->d_cell_send_idx, L.cell_send_idx));
        h->host_cell_send_idx = L.cell_send_idx;
        TRY(upload(h, &h->d_pix_send_idx, L.pix_send_idx));
        TRY(dalloc(h, &h->d_cell_send, (size_t)L.cell_send_idx.size() * 2));
        TRY(dalloc(h, &h->d_pix_send, (size_t)L.pix_send_idx.size() * h->F));
        TRY(dalloc(h, &h->d_means_loc, (size_t)h->n_tasks_glob));
        h->eager_steps = 2;            // NCCL sets its connections up lazily: the first steps run uncaptured
    }
    // ---- per-step buffers -------------------------------------------------------------------------------
    const int ncol = h->cfg.input_cols;
    const int Bp = h->B_pad;
    int maxw = 0;
    for (int d : h->dims_pad) maxw = d > maxw ? d : maxw;
    h->cells_capacity = (size_t)N * (ncol > 6 ? ncol : 6);          // rows, or U + dU as double[n][3] each (psm_predict_fields)
    TRY(dalloc(h, &h->d_cells, h->cells_capacity));
    TRY(dalloc(h, &h->d_out, (size_t)N * h->F));
    TRY(dalloc(h, &h->d_pprev, (size_t)N));
    if (h->cfg.variant == PSM_DELTAU_TO_DELTAP) TRY(dalloc(h, &h->d_uprev, (size_t)N * 2));     // resident U(t-1): 5-column rows, fields without dU
    TRY(dalloc(h, &h->d_uv, (size_t)Nall));
    TRY(dalloc(h, &h->d_grid, (size_t)h->grid_stride * 2));
    TRY(dalloc(h, &h->d_xu, (size_t)Bp * 2 * S2));
    {
        const int tiles = (Bp / 64) * (h->pc_in_pad / 64);
        int sp = (4 * 148 + tiles - 1) / tiles;
        if (sp > 64) sp = 64;
        if (sp < 1) sp = 1;
        h->splits = sp;
        // tcgen05 path: one 128-row tile per CTA, split-K so that about one wave of CTAs streams the operands
        const int kb_total = 2 * S2 / 32;
        const int tc_tiles = (Bp / 128) * (h->pc_in_pad / tc_gemm_bn(h->pc_in_pad));
        // at most ONE resident wave (1 CTA per SM): 7 row tiles x 22 splits = 154 CTAs ran as two waves at c3 (ncu: 76 us, tensor
        // pipe 50 %), so the split count is rounded down, not up
        int ts = 148 / tc_tiles;
        if (ts > 128) ts = 128;
        if (ts < 1) ts = 1;
        const int per = (kb_total + ts - 1) / ts;
        h->tc_splits = (kb_total + per - 1) / per;
    }
    TRY(dalloc(h, &h->d_part, (size_t)(h->splits > h->tc_splits ? h->splits : h->tc_splits) * Bp * h->pc_in_pad));
    TRY(dalloc(h, &h->d_xin, (size_t)Bp * h->pc_in_pad));
    TRY(dalloc(h, &h->d_act[0], (size_t)Bp * maxw));
    TRY(dalloc(h, &h->d_act[1], (size_t)Bp * maxw));
    TRY(dalloc(h, &h->d_r, (size_t)Bp * h->pc_p_pad));
    TRY(dalloc(h, &h->d_blocks, (size_t)Bp * h->C * S2 + (size_t)((L.n_ghost_pix + S2 - 1) / S2) * h->C * S2));   // + ghost pixels [chunk][C][S*S]
    TRY(dalloc(h, &h->d_means, (size_t)h->n_tasks_glob));
    for (int i = 0; i < 2; ++i) { TRY(dalloc(h, &h->d_dbuf[i], (size_t)h->F * h->Bg)); TRY(dalloc(h, &h->d_pbuf[i], (size_t)h->F * h->Bg)); }
    TRY(dalloc(h, &h->d_offsets, (size_t)h->F * h->Bg));
    TRY(dalloc(h, &h->d_coff, (size_t)h->F * h->Bg));
    TRY(dalloc(h, &h->d_field, (size_t)h->F * h->field_stride));
    TRY(dalloc(h, &h->d_sc, 1));
    TRY(dalloc(h, &h->d_zc, (size_t)Bp * h->pc_in_pad));

    if (h->cfg.variant == PSM_U_TO_GRADP && L.world == 1) {
        h->sdf_int.resize((size_t)L.H * W);
        for (size_t q = 0; q < h->sdf_int.size(); ++q) { const double v = L.sdf_rows[q]; h->sdf_int[q] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : (int)v)); }
    }
    // ---- static distance-channel contribution per block: zc[b][n] = sum_p sdf_n[b,p] * comp[n][p*3+2] ----
    // (grid[...,2] = sdfunct / max_abs_dist, SMC:434,443 -- constant for the mesh)
    {
        const size_t nl = (size_t)Hl * W;
        std::vector<float> sdfn(nl, 0.f);
        const double dist_scale = (h->cfg.variant == PSM_THESIS_U_TO_P) ? 1.0 : h->maxs[2];     // PMP:292 feeds the distance unscaled
        for (size_t q = 0; q < nl; ++q) sdfn[q] = (float)(L.sdf_rows[q] / dist_scale);
        float* d_sdfn = nullptr; float* d_sdfb = nullptr;
        CU(h, cudaMalloc(&d_sdfn, nl * sizeof(float)));
        CU(h, cudaMalloc(&d_sdfb, (size_t)Bp * S2 * sizeof(float)));
        CU(h, cudaMemsetAsync(d_sdfb, 0, (size_t)Bp * S2 * sizeof(float), h->stream));
        CU(h, cudaMemcpyAsync(d_sdfn, sdfn.data(), nl * sizeof(float), cudaMemcpyHostToDevice, h->stream));
        ExtractArgs ea{d_sdfn, d_sdfn, h->d_by0, h->d_bx0, d_sdfb, h->B, W, S, 1};
        launch_extract(ea, h->stream);
        GemmArgs ga{};
        ga.A = d_sdfb; ga.B = h->d_comp_sdf; ga.C = h->d_zc; ga.M = Bp; ga.N = h->pc_in_pad; ga.K = S2;
        ga.lda = S2; ga.ldb = S2; ga.ldc = h->pc_in_pad; ga.splits = 1; ga.epi = EPI_PLAIN;
        launch_sgemm(ga, h->stream);
        CU(h, cudaStreamSynchronize(h->stream));
        CU(h, cudaGetLastError());
        cudaFree(d_sdfn); cudaFree(d_sdfb);
    }
    if (h->cfg.gemm_mode != PSM_GEMM_FP32_SIMT) {
        // TMA tensor maps + argument blocks of the tcgen05 GEMMs (all pointers are fixed for the handle's life)
        if (tc_gemm_prepare() != 0) PSM_FAIL(h, PSM_ERR_CUDA, "cannot opt in to %d B of shared memory for the tcgen05 GEMM", 197888);
        // 3xTF32: 1 = converters store the masked hi tile, 2 = the raw FP32 tile is the hi operand (the tensor core drops
        // the low 13 mantissa bits itself -- verified bit-exact by tests/test_gpu_gemm.py::test_tf32_operand_truncation)
        const int three = (h->cfg.gemm_mode == PSM_GEMM_TC_3XTF32) ? (env_on("PSM_TF32_MASK_HI") ? 1 : 2) : 0;
        auto mk = [&](TcGemm& g, const float* A, int a_rows, const float* Bm, int b_rows, int K, float* Cp, int ldc, int splits,
                      int epi, const float* v0, const float* v1, const float* v2, int bn) -> int {
            if (make_kmajor_map(&g.mapA, A, a_rows, K, K, 128) != 0 || make_kmajor_map(&g.mapB, Bm, b_rows, K, K, bn) != 0)
                PSM_FAIL(h, PSM_ERR_CUDA, "cuTensorMapEncodeTiled failed");
            g.args = TcGemmArgs{Cp, a_rows, b_rows, K, ldc, splits, epi, three, v0, v1, v2, h->d_sc, nullptr, nullptr, env_on("PSM_NO_PREFETCH") ? 0 : 1};
            g.bn = bn;
            return 0;
        };
        TRY(mk(h->tc_proj, h->d_xu, Bp, h->d_comp_u, h->pc_in_pad, 2 * S2, h->d_part, h->pc_in_pad, h->tc_splits, EPI_PARTIAL,
               nullptr, nullptr, nullptr, tc_gemm_bn(h->pc_in_pad)));
        // ---- projection A operand straight from the grid planes (TMA boxes over overlapping windows) ----
        // Default on a single-GPU handle (PSM_NO_GRID_A=1: the gather writes x_array as before).  Measured (same box, back to back):
        // c2 96.0 vs 96.8 us, c5 232.8 vs 236.8 us, c3 264.2 vs 282.1 us per step (gather 15.6 vs 29.3 us at c3).
        if (h->fused_extract && L.world == 1 && h->pc_in_pad == 128 && S % 32 == 0 && (S - h->cfg.overlap) % 4 == 0 && !env_on("PSM_NO_GRID_A") &&
            tc_gemm_grid_prepare() == 0) {
            GridAPlan gp;
            std::vector<int32_t> hy(h->B), hx(h->B);
            CU(h, cudaMemcpy(hy.data(), h->d_by0, sizeof(int32_t) * h->B, cudaMemcpyDeviceToHost));
            CU(h, cudaMemcpy(hx.data(), h->d_bx0, sizeof(int32_t) * h->B, cudaMemcpyDeviceToHost));
            const int st = S - h->cfg.overlap;
            TcGemmGrid& tg = h->tc_proj_grid;
            if (build_grid_a(h->B, hy, hx, st, gp) &&
                make_grid_maps(&tg, h->d_grid, W, (int)(h->grid_stride / W), h->grid_stride, st, gp.gx, gp.gy) == 0 &&
                make_kmajor_map(&tg.mapB, h->d_comp_u, h->pc_in_pad, 2 * S2, 2 * S2, 128) == 0) {
                const int kb_total = 2 * S2 / 32;
                int ts = 148 / gp.tiles;                           // one resident wave (see tc_splits above)
                ts = std::max(1, std::min(ts, 128));
                const int per = (kb_total + ts - 1) / ts;
                const int splits = (kb_total + per - 1) / per;
                h->grid_a_rows = gp.tiles * 128;
                TRY(dalloc(h, &h->d_part_grid, (size_t)splits * h->grid_a_rows * 128));
                std::vector<int32_t> rs(Bp, -1);
                for (int b = 0; b < h->B; ++b) rs[b] = gp.row_src[b];
                TRY(upload(h, &h->d_row_src, rs));
                TRY(upload(h, &h->d_aseg, gp.segs)); TRY(upload(h, &h->d_aseg_ptr, gp.seg_ptr)); TRY(upload(h, &h->d_a_bytes, gp.a_bytes));
                tg.args = TcGemmArgs{h->d_part_grid, h->grid_a_rows, 128, 2 * S2, 128, splits, EPI_PARTIAL, three, nullptr, nullptr, nullptr, h->d_sc,
                                     nullptr, nullptr, env_on("PSM_NO_PREFETCH") ? 0 : 1};
                tg.ga = GridA{h->d_aseg, h->d_aseg_ptr, h->d_a_bytes, S};
                tg.tiles = gp.tiles;
                h->grid_a = true;
            }
        }
        // ---- projection + standardisation in one launch (cluster fold, last-arrival fold per row slab) ----
        // Measured on c2 (one B200, profiles/README.md round 2): 58.9 us for this kernel against 23.0 us for tc_gemm_kernel + the
        // reduce launch -- distributed shared memory moves ~17-21 B/clk per SM, so pushing a 64 KB partial per CTA costs more than
        // leaving it in L2.  Opt-in (PSM_PROJ_CLUSTER=1); parity-tested like the default path.
        if (h->pc_in_pad == 128 && env_on("PSM_PROJ_CLUSTER")) {
            const int tiles = Bp / 128, kb_total = 2 * S2 / 32;
            int best_ks = 0, best_ncl = 0;
            for (int ks = 8; ks >= 1; ks >>= 1) {                 // most CTAs of one resident wave; on a tie the larger cluster
                int max_cl = 0;
                if (proj_cluster_prepare(tiles, ks, &max_cl) != 0) continue;
                int ncl = std::min(max_cl, 148 / ks) / tiles;     // clusters per tile
                ncl = std::min(ncl, kb_total / ks);
                if (ncl < 1) continue;
                if (ncl * ks > best_ncl * best_ks) { best_ks = ks; best_ncl = ncl; }
            }
            if (const char* e = getenv("PSM_PROJ_KS")) {           // A/B: force the cluster size
                const int ks = atoi(e); int max_cl = 0;
                if ((ks == 1 || ks == 2 || ks == 4 || ks == 8) && proj_cluster_prepare(tiles, ks, &max_cl) == 0) {
                    const int ncl = std::min(std::min(max_cl, 148 / ks) / tiles, kb_total / ks);
                    if (ncl >= 1) { best_ks = ks; best_ncl = ncl; }
                }
            }
            if (best_ks > 0) {
                const int per = (kb_total + best_ks * best_ncl - 1) / (best_ks * best_ncl);
                int splits = (kb_total + per - 1) / per;
                splits = (splits + best_ks - 1) / best_ks * best_ks;      // whole clusters; surplus splits are empty (zero partials)
                TRY(dalloc(h, &h->d_proj_part, (size_t)(splits / best_ks) * Bp * 128));
                TRY(dalloc(h, &h->d_proj_cnt, (size_t)tiles * 8));
                ProjGemm& pg = h->proj_cl;
                if (make_kmajor_map(&pg.mapA, h->d_xu, Bp, 2 * S2, 2 * S2, 128) != 0 || make_kmajor_map(&pg.mapB, h->d_comp_u, h->pc_in_pad, 2 * S2, 2 * S2, 128) != 0)
                    PSM_FAIL(h, PSM_ERR_CUDA, "cuTensorMapEncodeTiled failed (projection)");
                pg.args = ProjArgs{Bp, 128, 2 * S2, splits, best_ks, three, env_on("PSM_NO_PREFETCH") ? 0 : 1, h->d_proj_part, h->d_proj_cnt,
                                   h->d_zc, h->d_in_a, h->d_in_b, h->d_xin};
                h->proj_cluster = true;
            }
        }
        // Dense layers: the batch is only B blocks (one or a few 128-row tiles), so every layer is cut into
        // 64-column tiles x split-K partials (each CTA streams <= ~100 KB), folded by the reduce kernel.
        h->tc_dense.resize(h->n_dense);
        h->dense_splits.assign(h->n_dense, 1);
        h->dense_cluster = !env_on("PSM_NO_DENSE_CLUSTER");
        if (h->dense_cluster && dense_cluster_prepare() != 0) PSM_FAIL(h, PSM_ERR_CUDA, "cannot opt in to the shared memory of the Dense cluster kernel");
        TRY(dalloc(h, &h->d_dpart, (size_t)8 * Bp * maxw));
        // pre-split operands: every producer of a Dense input also stores x - tf32(x), the Dense kernels' lo halves are static, so
        // a layer stages four tiles by TMA and issues its MMAs straight away (no in-kernel converter pass on the critical path)
        h->dense_presplit = h->dense_cluster && three == 2 && !env_on("PSM_NO_DENSE_PRESPLIT");
        if (h->dense_presplit) {
            TRY(dalloc(h, &h->d_xin_lo, (size_t)Bp * h->pc_in_pad));
            for (int i = 0; i < 2; ++i) if (!h->d_act_lo[i]) TRY(dalloc(h, &h->d_act_lo[i], (size_t)Bp * maxw));
            h->proj_cl.args.x_lo = h->d_xin_lo;
        }
        const float* in = h->d_xin;
        const float* in_lo = h->d_xin_lo;
        for (int l = 0; l < h->n_dense; ++l) {
            const bool last = (l == h->n_dense - 1);
            float* outp = last ? h->d_r : h->d_act[l & 1];
            const int kb = h->dims_pad[l] / 32;
            const int tiles = (Bp / 128) * (h->dims_pad[l + 1] / 64);
            if (h->dense_cluster) {
                // K-slices per tile = cluster size: as many as keep the grid within ~2 waves, at most one k-block each
                int ks = 8;
                // one resident wave of CTAs: at c5 / c3 (4 / 7 row tiles) the Dense stack takes 34.8 / 40.2 us against 57.3 / 55.4 us
                // with twice the K-slices over two waves (PSM_DENSE_TWO_WAVES=1); c2 (one tile) is the same either way
                const int waves = env_on("PSM_DENSE_TWO_WAVES") ? 2 : 1;
                while (ks > 1 && (ks > kb || tiles * ks > waves * 148)) ks >>= 1;
                h->dense_splits[l] = ks;
                TRY(mk(h->tc_dense[l], in, Bp, h->d_W[l], h->dims_pad[l + 1], h->dims_pad[l], outp, h->dims_pad[l + 1], ks,
                       last ? EPI_BIAS_AFFINE : EPI_BIAS_RELU, h->d_bias[l], h->d_out_s, h->d_out_m, 64));
                if (h->dense_presplit) {
                    TcGemm& g = h->tc_dense[l];
                    if (make_kmajor_map(&g.mapAlo, in_lo, Bp, h->dims_pad[l], h->dims_pad[l], 128) != 0 ||
                        make_kmajor_map(&g.mapBlo, h->d_Wlo[l], h->dims_pad[l + 1], h->dims_pad[l], h->dims_pad[l], 64) != 0)
                        PSM_FAIL(h, PSM_ERR_CUDA, "cuTensorMapEncodeTiled failed (Dense lo operands)");
                    g.args.presplit = 1;
                    if (!last) g.args.C_lo = h->d_act_lo[l & 1];
                    in_lo = h->d_act_lo[l & 1];
                }
                in = outp;
                continue;
            }
            int sp = (tiles >= 96) ? 1 : (kb >= 16 ? 4 : (kb >= 4 ? 2 : 1));
            const int per = (kb + sp - 1) / sp;
            sp = (kb + per - 1) / per;
            h->dense_splits[l] = sp;
            TRY(mk(h->tc_dense[l], in, Bp, h->d_W[l], h->dims_pad[l + 1], h->dims_pad[l], sp > 1 ? h->d_dpart : outp,
                   h->dims_pad[l + 1], sp, sp > 1 ? EPI_PARTIAL : (last ? EPI_BIAS_AFFINE : EPI_BIAS_RELU), h->d_bias[l],
                   h->d_out_s, h->d_out_m, 64));
            in = outp;
        }
        TRY(mk(h->tc_inv, h->d_r, Bp, h->d_comp_out_t, S2 * h->C, h->pc_p_pad, h->d_blocks, S2 * h->C, 1, EPI_PCA_INV,
               h->d_pmean, nullptr, nullptr, 128));
        if (h->dense_cluster && env_on("PSM_TRACE_LAYERS")) {
            TRY(dalloc(h, &h->d_layer_trace, (size_t)h->n_dense * 64 * 8));
            for (int l = 0; l < h->n_dense; ++l) h->tc_dense[l].args.trace = h->d_layer_trace + (size_t)l * 64 * 8;
        }
        // ---- transposed PCA inverse: needs the last Dense layer to deliver r pre-split (cluster or stack kernel) ----
        h->inv_t = !env_on("PSM_NO_INV_T") && h->dense_cluster && h->pc_p_pad <= pca_inverse_t_max_k() && (S2 * h->C) % 128 == 0 &&
                   pca_inverse_t_prepare() == 0;
        if (h->inv_t) {
            TRY(dalloc(h, &h->d_r_hi, (size_t)Bp * h->pc_p_pad));
            TRY(dalloc(h, &h->d_r_lo, (size_t)Bp * h->pc_p_pad));
            h->tc_dense[h->n_dense - 1].args.C_hi = h->d_r_hi;
            h->tc_dense[h->n_dense - 1].args.C_lo = h->d_r_lo;
            InvT& it = h->tc_inv_t;
            if (make_kmajor_map(&it.mapA, h->d_comp_out_t, S2 * h->C, h->pc_p_pad, h->pc_p_pad, 128) != 0 ||
                make_kmajor_map(&it.mapBhi, h->d_r_hi, Bp, h->pc_p_pad, h->pc_p_pad, 128) != 0 ||
                make_kmajor_map(&it.mapBlo, h->d_r_lo, Bp, h->pc_p_pad, h->pc_p_pad, 128) != 0)
                PSM_FAIL(h, PSM_ERR_CUDA, "cuTensorMapEncodeTiled failed (PCA inverse)");
            it.args = InvTArgs{h->d_blocks, (long long)S2 * h->C, h->B, Bp, h->pc_p_pad, S2 * h->C, three, h->d_pmean, h->d_sc, StripRows{}};
            // The masked strip sums can come out of this kernel's epilogue (the blocks are not read again); needs the CTA <-> pixel-row
            // match S == 128 and one resident wave of CTAs (with two output channels -- 256 CTAs -- the longer epilogue sits on the
            // critical path twice: 117 vs 61 us at configs[2]).  Measured on one GPU (same box, c2 / c5): 99.1 / 267.4 us per step with
            // the fused epilogue against 96.6 / 263.1 us with the separate task_means_kernel, whose 263 CTAs start under the previous
            // kernel (PDL) -- so a single-GPU handle keeps the separate kernel.  A sharded handle takes the epilogue sums: the fold
            // kernel that follows carries the whole exchange, 136 vs 151-161 us per step at two GPUs.
            const bool want = (L.world > 1 || env_on("PSM_STRIP_FUSE")) && !env_on("PSM_NO_STRIP_FUSE");
            h->strip_fuse = want && S == 128 && h->n_tasks > 0 && ((S2 * h->C) / 128 <= 148 || env_on("PSM_FORCE_STRIP_FUSE"));
            if (h->strip_fuse) it.args.strips = StripRows{h->d_sr_rowptr, h->d_sr_src, h->d_sr_slot, h->d_sr_w, h->sr_n_ent, h->d_rowpart};
        }
        // ---- the whole Dense stack as one persistent launch -------------------------------------------------
        int max_cl = 0;
        h->dense_stack = env_on("PSM_DENSE_STACK") && h->n_dense <= kMaxDense && dense_stack_prepare(&max_cl) == 0;
        if (h->dense_stack) {
            for (int i = 0; i < 2; ++i) { TRY(dalloc(h, &h->d_act_hi[i], (size_t)Bp * maxw)); TRY(dalloc(h, &h->d_act_lo[i], (size_t)Bp * maxw)); }
            std::vector<TensorMap128> maps((size_t)4 * h->n_dense);
            DenseStackArgs& sa = h->stack_args;
            sa = DenseStackArgs{};
            int max_tiles = 1;
            for (int l = 0; l < h->n_dense; ++l) {
                const bool last = (l == h->n_dense - 1);
                const int K = h->dims_pad[l], N = h->dims_pad[l + 1];
                if (make_kmajor_map(&maps[4 * l + 0], h->d_act_hi[l & 1], Bp, K, K, 128) != 0 ||
                    make_kmajor_map(&maps[4 * l + 1], h->d_act_lo[l & 1], Bp, K, K, 128) != 0 ||
                    make_kmajor_map(&maps[4 * l + 2], h->d_Whi[l], N, K, K, 64) != 0 ||
                    make_kmajor_map(&maps[4 * l + 3], h->d_Wlo[l], N, K, K, 64) != 0)
                    PSM_FAIL(h, PSM_ERR_CUDA, "cuTensorMapEncodeTiled failed (Dense stack)");
                sa.L[l] = DenseLayerDesc{K, N, last ? EPI_BIAS_AFFINE : EPI_BIAS_RELU, h->d_bias[l], h->d_out_s, h->d_out_m,
                                         last ? h->d_r : nullptr, last ? h->d_r_hi : h->d_act_hi[(l + 1) & 1],
                                         last ? h->d_r_lo : h->d_act_lo[(l + 1) & 1]};
                max_tiles = std::max(max_tiles, (Bp / 128) * (N / 64));
            }
            TRY(upload(h, &h->d_stack_maps, maps));
            sa.maps = h->d_stack_maps; sa.n_layers = h->n_dense; sa.M = Bp; sa.three_pass = three;
            sa.barrier = &h->d_sc->dense_barrier; sa.error = &h->d_sc->comm_error;
            h->stack_clusters = std::min(std::min(max_cl, 16), max_tiles);
        }
    }
    if (L.world > 1) TRY(setup_p2p(h));
    if (h->cfg.enable_timings) TRY(psm_set_timings(h, 1));
    CU(h, cudaStreamSynchronize(h->stream));
    h->initialised = true;
    return PSM_OK;
}

extern "C" int psm_init_with_tables(psm_handle* h, const psm_tables* t) {
    if (!h) return PSM_ERR_INVALID;
    if (!t) PSM_FAIL(h, PSM_ERR_INVALID, "psm_init_with_tables: NULL tables");
    if (!h->params_loaded) PSM_FAIL(h, PSM_ERR_STATE, "psm_load_params must be called before psm_init_with_tables");
    if (h->initialised) PSM_FAIL(h, PSM_ERR_STATE, "handle already initialised");
    if (h->comm) PSM_FAIL(h, PSM_ERR_STATE, "a communicator is attached: use psm_init_sharded");
    if (!t->vert || !t->weights || !t->indices || !t->sdfunct || t->n_cells < 3 || t->grid_h < 1 || t->grid_w < 1)
        PSM_FAIL(h, PSM_ERR_INVALID, "bad tables");
    CU(h, cudaSetDevice(h->cfg.device));
    const int H = t->grid_h, W = t->grid_w;
    const long long G = (long long)H * W, N = t->n_cells;
    LocalInit L;
    L.H = H; L.W = W; L.row0 = 0; L.row1 = H; L.n_owned = N;
    // static flow mask: x_array[...,2] != 0 (SMC:224) <=> sdfunct != 0
    std::vector<uint8_t> mask(G);
    for (long long q = 0; q < G; ++q) mask[q] = (t->sdfunct[q] != 0.0) ? 1 : 0;
    L.mask_global = mask.data();
    L.sdf_rows = t->sdfunct;
    L.blk_row0 = 0; L.blk_row1 = -1;      // all block rows
    // ---- forward table with the validity fold (SMC:161-178,432-438; UTL:89) ------------------------
    // grid[...][tuple(indices.T)] = interp : for duplicate targets the LAST source point wins, so
    // pixel (0,0) receives the value of the last invalid point; untouched pixels stay 0.
    {
        std::vector<long long> src(G, -1);
        for (long long m = 0; m < G; ++m) {
            const long long ii = t->indices[2 * m], jj = t->indices[2 * m + 1];
            if (ii < 0 || ii >= H || jj < 0 || jj >= W) PSM_FAIL(h, PSM_ERR_INVALID, "indices out of range at %lld", m);
            src[ii * W + jj] = m;
        }
        for (int j = 0; j < 3; ++j) { L.fv[j].assign(G, 0); L.fw[j].assign(G, 0.f); }
        for (long long q = 0; q < G; ++q) {
            const long long m = src[q];
            if (m < 0) continue;
            const double* wm = t->weights + 3 * m;
            // negative weight -> NaN (interpolate_fill, UTL:89) -> 0 (SMC:438); the thesis module interpolates without
            // the fill (PMP:64-65,280-281): outside points keep their extrapolated value
            const bool neg = h->cfg.variant != PSM_THESIS_U_TO_P && ((wm[0] < 0) || (wm[1] < 0) || (wm[2] < 0));
            for (int j = 0; j < 3; ++j) {
                const int32_t vi = t->vert[3 * m + j];
                if (vi < 0 || vi >= N) PSM_FAIL(h, PSM_ERR_INVALID, "vert out of range at %lld", m);
                L.fv[j][q] = vi;
                L.fw[j][q] = neg ? 0.f : (float)wm[j];
            }
        }
    }
    // ---- back table (PMP:481-496): vertex ids hop through `indices` (incl. the (0,0) quirk) ---------
    L.have_back = (t->vert_back && t->weights_back);
    if (L.have_back) {
        for (int j = 0; j < 3; ++j) { L.bv[j].assign(N, 0); L.bw[j].assign(N, 0.f); }
        const bool near_wall = h->cfg.near_wall_sdf > 0.0;
        for (long long c = 0; c < N; ++c) {
            const double* wc = t->weights_back + 3 * c;
            bool keep_prev = (wc[0] < 0) || (wc[1] < 0) || (wc[2] < 0);         // interpolate_fill -> NaN -> p_prev (PMP:496)
            double sdf_mesh = 0.0;
            for (int j = 0; j < 3; ++j) {
                const long long g = t->vert_back[3 * c + j];
                if (g < 0 || g >= G) PSM_FAIL(h, PSM_ERR_INVALID, "vert_back out of range at %lld", c);
                sdf_mesh += t->sdfunct[g] * wc[j];                               // PMP:492 (flat take, no indices hop)
                L.bv[j][c] = (int32_t)(t->indices[2 * g] * W + t->indices[2 * g + 1]);
                L.bw[j][c] = (float)wc[j];
            }
            if (near_wall && !keep_prev && sdf_mesh < h->cfg.near_wall_sdf) keep_prev = true;   // PMP:494
            if (keep_prev) L.bv[0][c] = -1;
        }
    }
    return init_local(h, L);
}

// Multi-GPU: this rank's pre-folded share (psm_b200/shard.py).
extern "C" int psm_init_sharded(psm_handle* h, const psm_shard* s) {
    if (!h) return PSM_ERR_INVALID;
    if (!s) PSM_FAIL(h, PSM_ERR_INVALID, "psm_init_sharded: NULL shard");
    if (!h->params_loaded) PSM_FAIL(h, PSM_ERR_STATE, "psm_load_params must be called before psm_init_sharded");
    if (h->initialised) PSM_FAIL(h, PSM_ERR_STATE, "handle already initialised");
    if (s->world < 1 || s->rank < 0 || s->rank >= s->world) PSM_FAIL(h, PSM_ERR_INVALID, "bad rank/world");
    if (s->world > 1 && (!h->comm || h->world != s->world || h->rank != s->rank)) PSM_FAIL(h, PSM_ERR_STATE, "psm_comm_init(rank %d of %d) must precede psm_init_sharded", s->rank, s->world);
    if (!s->vert || !s->weights || !s->sdfunct || !s->mask_global || s->n_owned < 1 || s->row0 < 0 || s->row1 <= s->row0 || s->row1 > s->grid_h ||
        s->ext_rows < 0 || s->send_rows < 0 || s->local_ext_rows < 0 || s->row1 + s->ext_rows + s->local_ext_rows > s->grid_h)
        PSM_FAIL(h, PSM_ERR_INVALID, "bad shard");
    if (s->world > 1 && (!s->cell_send_ptr || !s->cell_recv_ptr || !s->pix_send_ptr || !s->pix_recv_ptr)) PSM_FAIL(h, PSM_ERR_INVALID, "missing exchange lists");
    if ((s->rank == s->world - 1 && s->ext_rows) || (s->rank == 0 && s->send_rows)) PSM_FAIL(h, PSM_ERR_INVALID, "halo rows at the ends of the rank chain");
    if (h->cfg.variant == PSM_THESIS_U_TO_P) PSM_FAIL(h, PSM_ERR_INVALID, "the thesis variant is not available on a sharded handle");
    CU(h, cudaSetDevice(h->cfg.device));
    LocalInit L;
    L.rank = s->rank; L.world = s->world; L.H = s->grid_h; L.W = s->grid_w; L.row0 = s->row0; L.row1 = s->row1;
    L.ext_rows = s->ext_rows; L.send_rows = s->send_rows; L.blk_row0 = s->blk_row0; L.blk_row1 = s->blk_row1;
    L.local_ext = s->local_ext_rows;
    L.mask_global = s->mask_global; L.n_owned = s->n_owned; L.n_ghost = s->n_ghost; L.n_ghost_pix = s->n_ghost_pix;
    L.sdf_rows = s->sdfunct;
    const long long G = (long long)(s->row1 - s->row0 + s->local_ext_rows) * s->grid_w;
    for (int j = 0; j < 3; ++j) {
        L.fv[j].resize(G); L.fw[j].resize(G);
        for (long long q = 0; q < G; ++q) { L.fv[j][q] = s->vert[3 * q + j]; L.fw[j][q] = (float)s->weights[3 * q + j]; }
    }
    L.have_back = (s->vert_back && s->weights_back);
    if (L.have_back)
        for (int j = 0; j < 3; ++j) {
            L.bv[j].resize(s->n_owned); L.bw[j].resize(s->n_owned);
            for (long long c = 0; c < s->n_owned; ++c) { L.bv[j][c] = s->vert_back[3 * c + j]; L.bw[j][c] = (float)s->weights_back[3 * c + j]; }
        }
    if (s->world > 1) {
        const int Wd = s->world;
        L.cell_send_ptr.assign(s->cell_send_ptr, s->cell_send_ptr + Wd + 1);
        L.cell_recv_ptr.assign(s->cell_recv_ptr, s->cell_recv_ptr + Wd + 1);
        L.pix_send_ptr.assign(s->pix_send_ptr, s->pix_send_ptr + Wd + 1);
        L.pix_recv_ptr.assign(s->pix_recv_ptr, s->pix_recv_ptr + Wd + 1);
        if (L.cell_send_ptr[Wd] > 0 && !s->cell_send_idx) PSM_FAIL(h, PSM_ERR_INVALID, "cell_send_idx is NULL");
        if (L.pix_send_ptr[Wd] > 0 && !s->pix_send_idx) PSM_FAIL(h, PSM_ERR_INVALID, "pix_send_idx is NULL");
        L.cell_send_idx.assign(s->cell_send_idx, s->cell_send_idx + L.cell_send_ptr[Wd]);
        L.pix_send_idx.assign(s->pix_send_idx, s->pix_send_idx + L.pix_send_ptr[Wd]);
        if (s->ghost_pix && s->n_ghost_pix > 0) L.ghost_pix.assign(s->ghost_pix, s->ghost_pix + s->n_ghost_pix);
    }
    return init_local(h, L);
}

// ---- communicator --------------------------------------------------------------------------------------
extern "C" int psm_comm_get_unique_id(void* id_out) {
    if (!id_out) return PSM_ERR_INVALID;
    static_assert(sizeof(ncclUniqueId) <= PSM_UNIQUE_ID_BYTES, "unique id size");
    if (!g_nccl.load()) { g_create_error = g_nccl.error; return PSM_ERR_COMM; }
    ncclUniqueId id;
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) { g_create_error = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r); return PSM_ERR_COMM; }
    memset(id_out, 0, PSM_UNIQUE_ID_BYTES);
    memcpy(id_out, &id, sizeof id);
    return PSM_OK;
}

extern "C" int psm_comm_init(psm_handle* h, const void* unique_id, int32_t rank, int32_t world) {
    if (!h) return PSM_ERR_INVALID;
    if (!unique_id || world < 1 || rank < 0 || rank >= world) PSM_FAIL(h, PSM_ERR_INVALID, "psm_comm_init: bad arguments");
    if (h->initialised || h->comm) PSM_FAIL(h, PSM_ERR_STATE, "psm_comm_init must be called once, before psm_init_sharded");
    if (!g_nccl.load()) PSM_FAIL(h, PSM_ERR_COMM, "%s", g_nccl.error.c_str());
    CU(h, cudaSetDevice(h->cfg.device));
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof id);
    ncclResult_t r = g_nccl.CommInitRank(&h->comm, world, id, rank);
    if (r != ncclSuccess) { h->comm = nullptr; PSM_FAIL(h, PSM_ERR_COMM, "ncclCommInitRank: %s", g_nccl.GetErrorString(r)); }
    h->rank = rank; h->world = world;
    return PSM_OK;
}

// ------------------------------------------------------------------------------------------------
// One step on the device, input already in d_cells, output to d_out.
// Static sparse exchange: packed send buffer -> the ghost region of every peer (grouped send/recv).
static int sparse_exchange(psm_handle* h, const float* sendbuf, const std::vector<long long>& sp, float* recvbuf,
                           const std::vector<long long>& rp, int width) {
    for (int p = 0; p < h->world; ++p) {
        if (p == h->rank) continue;
        const long long ns = sp[p + 1] - sp[p], nr = rp[p + 1] - rp[p];
        if (ns > 0) NC(h, g_nccl.Send(sendbuf + sp[p] * width, (size_t)ns * width, ncclFloat, p, h->comm, h->stream));
        if (nr > 0) NC(h, g_nccl.Recv(recvbuf + rp[p] * width, (size_t)nr * width, ncclFloat, p, h->comm, h->stream));
    }
    return PSM_OK;
}

static int run_step(psm_handle* h, const StepInput& in, double* d_out) {
    cudaStream_t s = h->stream;
    const int S = h->S, S2 = S * S, Bp = h->B_pad;
    const bool deltas = h->cfg.variant == PSM_DELTAU_TO_DELTAP;
    const bool multi = h->world > 1;
    const bool fields = in.U != nullptr;
    const int mode = !deltas ? 0 : (fields ? (in.dU ? 1 : 2) : (h->cfg.input_cols == 7 ? 1 : 2));
    int nl = 0, te = 1;
    auto tick = [&]() { if (h->ev_valid) cudaEventRecord(h->ev[te], s); ++te; };

    const bool dim = h->cfg.variant != PSM_U_TO_GRADP;        // blocks re-dimensionalised by max_abs_p * U^2 (SMC:551, PMP:490); GRAD: none
    ScalarArgs sa{h->d_sc, h->maxs[0], h->maxs[1], dim ? h->maxs[3] : 1.0, dim ? 1 : 0, h->cfg.skip_threshold, mode, 0};
    const bool p2p = multi && h->p2p;
    const P2PArgs* d_p2p = p2p ? h->d_p2p : nullptr;
    // fused multi-GPU flow: the exchanges ride in the tails of prep and of the fold kernel (no push launches, no assembled field)
    const bool mfused = p2p && h->mgpu_fused && d_out != nullptr;
    P2PFused fx{};
    if (mfused) fx = P2PFused{h->d_p2p, h->d_cell_send_idx, h->d_runs, h->n_runs, h->d_blocks, h->d_pix_send_blk, h->C, S2, h->pix_send_ptr[h->world],
                             h->d_send_words, h->d_send_entries};
    if (fields) {
        PrepFieldsArgs pf{in.U, in.dU, in.u_stride, h->n_cells, mode, h->d_uv, h->d_uprev, h->d_sc, fx};
        launch_prep_fields(pf, s); ++nl;
    } else {
        PrepArgs pa{in.rows, h->n_cells, h->cfg.input_cols, mode, h->d_uv, h->d_pprev, h->d_uprev, h->d_sc, fx};
        launch_prep(pa, s); ++nl;
        h->pprev_zero = false;
    }
    if (mfused) {
        // exchange 1 happened in the tail of prep
    } else if (p2p) {
        // exchange 1 over peer memory: maxima + ghost cells are pushed, the gather kernel waits on its mailbox
        launch_p2p_push_cells(h->d_p2p, h->d_uv, h->d_cell_send_idx, h->cell_send_ptr[h->world], s); ++nl;
    } else if (multi) {
        // exchange 1: max|U|^2, max|dU|^2 over all ranks (non-negative doubles order like their bit patterns)
        //             + the ghost cells other ranks' forward tables reference
        const long long nsend = h->cell_send_ptr[h->world];
        if (nsend > 0) { PackArgs pk{reinterpret_cast<const float*>(h->d_uv), h->d_cell_send_idx, h->d_cell_send, nsend, 2, 0}; launch_pack(pk, s); ++nl; }
        NC(h, g_nccl.GroupStart());
        NC(h, g_nccl.AllReduce(&h->d_sc->umax2_bits, &h->d_sc->umax2_bits, 2, ncclUint64, ncclMax, h->comm, s));
        TRY(sparse_exchange(h, h->d_cell_send, h->cell_send_ptr, reinterpret_cast<float*>(h->d_uv + h->n_cells), h->cell_recv_ptr, 2));
        NC(h, g_nccl.GroupEnd());
    }
    tick();   // prep
    float* grid0 = h->d_grid; float* grid1 = h->d_grid + h->grid_stride;
    GatherArgs ga{h->d_fv[0], h->d_fv[1], h->d_fv[2], h->d_fw[0], h->d_fw[1], h->d_fw[2], h->d_uv,
                  grid0, grid1, h->G_pad / 4, sa, 0, d_p2p};
    if (h->fused_extract) {
        GatherExtractArgs ge{ga, h->d_rowcov, h->d_colcov, h->d_xu, h->W / 4, S, (h->keep_grid || h->grid_a) ? 1 : 0, h->grid_a ? 1 : 0};
        launch_gather_extract(ge, s); ++nl;
    } else {
        launch_gather(ga, s); ++nl;
    }
    if (multi && (h->ext_rows || h->send_rows)) {
        // exchange 2: the overlap strip -- the first rows of rank+1 complete this rank's last block row
        NC(h, g_nccl.GroupStart());
        if (h->send_rows) {
            NC(h, g_nccl.Send(grid0, (size_t)h->send_rows * h->W, ncclFloat, h->rank - 1, h->comm, s));
            NC(h, g_nccl.Send(grid1, (size_t)h->send_rows * h->W, ncclFloat, h->rank - 1, h->comm, s));
        }
        if (h->ext_rows) {
            NC(h, g_nccl.Recv(grid0 + h->Gg, (size_t)h->ext_rows * h->W, ncclFloat, h->rank + 1, h->comm, s));
            NC(h, g_nccl.Recv(grid1 + h->Gg, (size_t)h->ext_rows * h->W, ncclFloat, h->rank + 1, h->comm, s));
        }
        NC(h, g_nccl.GroupEnd());
    }
    tick();   // gather
    if (!h->fused_extract) {
        ExtractArgs ea{grid0, grid1, h->d_by0, h->d_bx0, h->d_xu, h->B, h->W, S, 2};
        launch_extract(ea, s); ++nl;
    }
    tick();   // extract
    const bool tc = h->cfg.gemm_mode != PSM_GEMM_FP32_SIMT;
    {
        const bool one_launch = tc && h->proj_cluster && !h->dense_stack && !h->grid_a;
        const bool from_grid = tc && h->grid_a && !h->dense_stack;
        if (from_grid) launch_tc_gemm_grid(h->tc_proj_grid, s);
        else if (one_launch) {
            if (launch_proj_cluster(h->proj_cl, s) != 0) PSM_FAIL(h, PSM_ERR_CUDA, "projection launch: %s", cudaGetErrorString(cudaGetLastError()));
        } else if (tc) launch_tc_gemm(h->tc_proj, s);
        else {
            GemmArgs g{};
            g.A = h->d_xu; g.B = h->d_comp_u; g.C = h->d_part; g.M = Bp; g.N = h->pc_in_pad; g.K = 2 * S2;
            g.lda = 2 * S2; g.ldb = 2 * S2; g.ldc = h->pc_in_pad; g.splits = h->splits; g.epi = EPI_PARTIAL;
            launch_sgemm(g, s);
        }
        ++nl;
        ReduceArgs r{h->d_part, tc ? h->tc_splits : h->splits, Bp, h->pc_in_pad, h->d_zc, h->d_in_a, h->d_in_b, h->d_xin,
                     RED_STANDARDISE, nullptr, nullptr, nullptr};
        if (tc && h->dense_stack) { r.x_hi = h->d_act_hi[0]; r.x_lo = h->d_act_lo[0]; }
        else if (tc && h->dense_presplit) r.x_lo = h->d_xin_lo;
        if (from_grid) { r.part = h->d_part_grid; r.splits = h->tc_proj_grid.args.splits; r.row_src = h->d_row_src; r.part_rows = h->grid_a_rows; }
        if (!one_launch) { launch_reduce_standardise(r, s); ++nl; }
    }
    tick();   // pca_project
    if (tc && h->dense_stack) {
        if (launch_dense_stack(h->stack_args, h->stack_clusters, s) != 0) PSM_FAIL(h, PSM_ERR_CUDA, "Dense stack launch: %s", cudaGetErrorString(cudaGetLastError()));
        ++nl;
    } else if (tc && h->dense_cluster) {
        for (int l = 0; l < h->n_dense; ++l) {
            if (launch_dense_cluster(h->tc_dense[l], s) != 0) PSM_FAIL(h, PSM_ERR_CUDA, "Dense cluster launch: %s", cudaGetErrorString(cudaGetLastError()));
            ++nl;
        }
    } else {
        const float* in = h->d_xin;
        for (int l = 0; l < h->n_dense; ++l) {
            const bool last = (l == h->n_dense - 1);
            GemmArgs g{};
            g.A = in; g.B = h->d_W[l]; g.C = last ? h->d_r : h->d_act[l & 1];
            g.M = Bp; g.N = h->dims_pad[l + 1]; g.K = h->dims_pad[l];
            g.lda = g.K; g.ldb = g.K; g.ldc = g.N; g.splits = 1;
            g.epi = last ? EPI_BIAS_AFFINE : EPI_BIAS_RELU;
            g.v0 = h->d_bias[l]; g.v1 = h->d_out_s; g.v2 = h->d_out_m;
            if (tc) {
                launch_tc_gemm(h->tc_dense[l], s);
                if (h->dense_splits[l] > 1) {
                    ReduceArgs r{h->d_dpart, h->dense_splits[l], Bp, g.N, nullptr, h->d_out_s, h->d_out_m, g.C,
                                 last ? RED_BIAS_AFFINE : RED_BIAS_RELU, h->d_bias[l]};
                    launch_reduce_standardise(r, s); ++nl;
                }
            } else launch_sgemm(g, s);
            ++nl;
            in = g.C;
        }
    }
    tick();   // mlp
    {
        GemmArgs g{};
        g.A = h->d_r; g.B = h->d_comp_out_t; g.C = h->d_blocks; g.M = Bp; g.N = S2 * h->C; g.K = h->pc_p_pad;
        g.lda = g.K; g.ldb = g.K; g.ldc = g.N; g.splits = 1; g.epi = EPI_PCA_INV; g.v0 = h->d_pmean; g.sc = h->d_sc;
        if (tc && h->inv_t) launch_pca_inverse_t(h->tc_inv_t, s);
        else if (tc) launch_tc_gemm(h->tc_inv, s);
        else launch_sgemm(g, s);
        ++nl;
    }
    tick();   // pca_inverse
    OffsetsArgs oa{};
    oa.rec = h->d_rec; oa.B = h->Bg; oa.F = h->F; oa.rounds = h->rounds; oa.ref_bc = h->cfg.ref_bc; oa.means = h->d_means;
    oa.dbuf0 = h->d_dbuf[0]; oa.dbuf1 = h->d_dbuf[1]; oa.pbuf0 = h->d_pbuf[0]; oa.pbuf1 = h->d_pbuf[1];
    oa.offsets = h->d_offsets; oa.coff = h->d_coff; oa.terms = h->d_terms;
    for (int f = 0; f < 3; ++f) oa.term_start[f] = h->term_start[f];
    for (int f = 0; f < 2; ++f) oa.shift_len[f] = h->plan.shift_len[f];
    oa.sc = h->d_sc;
    oa.host_skip = h->d_host_skip;
    oa.p2p = d_p2p;
    const bool use_fold = tc && h->inv_t && h->strip_fuse;      // row partials left by the PCA-inverse epilogue
    const bool fuse_offsets = (!multi || mfused) && h->n_tasks > 0 && h->Bg * h->F <= (use_fold ? 4096 : 1024) && !h->no_fused_offsets;
    MeansArgs ma{h->d_tasks, h->n_tasks, h->d_blocks, h->d_gmask, h->C, S, h->W, (multi && !p2p) ? h->d_means_loc : h->d_means};
    if (mfused && !(use_fold && fuse_offsets)) PSM_FAIL(h, PSM_ERR_STATE, "fused multi-GPU flow without the fold kernel");
    if (use_fold) launch_fold(ma, h->d_rowpart, fuse_offsets ? &oa : nullptr, fx, s);
    else launch_means(ma, fuse_offsets ? &oa : nullptr, s);
    ++nl;
    if (mfused) {
        // exchange 2 (strip means) and 3 (ghost pixels, straight from the blocks) happened inside the fold kernel
    } else if (p2p && !fuse_offsets) { launch_p2p_push_means(h->d_p2p, h->d_tasks, h->n_tasks, h->world, h->d_means, s); ++nl; }   // exchange 3 over peer memory
    else if (multi)   // exchange 3: every rank contributes its own slots (zero elsewhere) -> identical means everywhere
        NC(h, g_nccl.AllReduce(h->d_means_loc, h->d_means, (size_t)h->n_tasks_glob, ncclDouble, ncclSum, h->comm, s));
    tick();   // strip_means
    if (!fuse_offsets) { launch_offsets(oa, s); ++nl; }
    tick();   // offsets
    PlaceArgs pl{h->d_blocks, h->d_owner, h->d_by0, h->d_bx0, h->d_coff, h->d_field, h->Bg, h->kb0, h->C, h->F, S, h->H, h->W,
                 h->field_stride};
    const bool fuse_place = h->fuse_place && d_out != nullptr && (!multi || mfused);
    if (!fuse_place) { launch_place(pl, s); ++nl; }
    if (h->filter_radius > 0)
        for (int f = 0; f < h->F; ++f) {      // axis 0, then axis 1, like scipy.ndimage.gaussian_filter
            float* fld = h->d_field + (size_t)f * h->field_stride;
            launch_gauss(GaussArgs{fld, h->d_filter_tmp, h->H, h->W, h->d_gauss_w, h->filter_radius, 0}, s);
            launch_gauss(GaussArgs{h->d_filter_tmp, fld, h->H, h->W, h->d_gauss_w, h->filter_radius, 1}, s);
            nl += 2;
        }
    tick();   // place
    if (h->have_back && d_out) {
        if (mfused) {
            // ghost pixels already sit behind the local blocks (pushed by their owners from the fold kernel)
        } else if (p2p) {
            launch_p2p_push_pix(h->d_p2p, h->d_field, h->d_pix_send_idx, h->F, h->field_stride, h->pix_send_ptr[h->world], s); ++nl;   // exchange 4 over peer memory
        } else if (multi) {
            // exchange 4: the field pixels other ranks' grid->cell tables reference (rows next to a rank
            //             boundary, and pixel (0,0) for the raster quirk of PMP:481)
            const long long nsend = h->pix_send_ptr[h->world];
            if (nsend > 0) { PackArgs pk{h->d_field, h->d_pix_send_idx, h->d_pix_send, nsend, h->F, h->field_stride}; launch_pack(pk, s); ++nl; }
            NC(h, g_nccl.GroupStart());
            for (int f = 0; f < h->F; ++f)
                TRY(sparse_exchange(h, h->d_pix_send + (size_t)f * nsend, h->pix_send_ptr, h->d_field + (size_t)f * h->field_stride + h->G,
                                    h->pix_recv_ptr, 1));
            NC(h, g_nccl.GroupEnd());
        }
        BackArgs ba{h->d_bv[0], h->d_bv[1], h->d_bv[2], h->d_bw[0], h->d_bw[1], h->d_bw[2], h->d_field, h->n_cells,
                    h->field_stride, (fields && in.p_dev) ? in.p_dev : h->d_pprev, d_out, h->F, h->cfg.additive, h->d_sc, d_p2p, nullptr, nullptr, nullptr, nullptr, 0, 0};
        if (fuse_place) {
            ba.v0 = h->d_bb[0]; ba.v1 = h->d_bb[1]; ba.v2 = h->d_bb[2]; ba.o0 = h->d_bo[0]; ba.o1 = h->d_bo[1]; ba.o2 = h->d_bo[2];
            ba.field = h->d_blocks; ba.plane = S2; ba.coff = h->d_coff; ba.n_blocks = h->Bg; ba.block_plane = S2;
        }
        const int nch = in.tail_chunks > 1 ? in.tail_chunks : 1;
        if (nch == 1) {
            if (in.wait_p) CU(h, cudaStreamWaitEvent(s, h->ev_p, 0));      // the caller's p array has landed in d_pprev
            launch_back(ba, s); ++nl;
            if (in.host_out) CU(h, cudaMemcpyAsync(in.host_out, d_out, (size_t)h->n_cells * h->F * sizeof(double), cudaMemcpyDeviceToHost, s));
        } else {
            // chunked tail (host entry points): gather chunk k as soon as chunk k of p has landed, copy it out on the second copy
            // stream while chunk k+1 of p is still on its way up -- the two directions of the link work at the same time
            for (int k = 0; k < nch; ++k) {
                long long c0, c1; tail_range(h->n_cells, nch, k, &c0, &c1);
                if (c1 <= c0) continue;
                if (in.wait_p) CU(h, cudaStreamWaitEvent(s, h->ev_pc[k], 0));
                BackArgs bk = ba;
                bk.v0 += c0; bk.v1 += c0; bk.v2 += c0; bk.w0 += c0; bk.w1 += c0; bk.w2 += c0;
                if (bk.o0) { bk.o0 += c0; bk.o1 += c0; bk.o2 += c0; }
                bk.p_prev += c0; bk.out += c0 * h->F; bk.n = c1 - c0;
                launch_back(bk, s); ++nl;
                if (in.host_out) {
                    CU(h, cudaEventRecord(h->ev_bc[k], s));
                    CU(h, cudaStreamWaitEvent(h->d2h_stream, h->ev_bc[k], 0));
                    CU(h, cudaMemcpyAsync(in.host_out + c0 * h->F, d_out + c0 * h->F, (size_t)(c1 - c0) * h->F * sizeof(double), cudaMemcpyDeviceToHost, h->d2h_stream));
                }
            }
            if (in.host_out) { CU(h, cudaEventRecord(h->ev_d2h, h->d2h_stream)); CU(h, cudaStreamWaitEvent(s, h->ev_d2h, 0)); }
        }
    }
    tick();   // back_gather
    h->launches = nl;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) PSM_FAIL(h, PSM_ERR_CUDA, "kernel launch: %s", cudaGetErrorString(e));
    return PSM_OK;
}

// Device entry points: one step as a captured CUDA graph, keyed by the caller's device buffers (a solver passes the same
// ones every step); per-stage events, PSM_NO_GRAPHS=1 and the first multi-GPU steps launch eagerly.
static int submit_device(psm_handle* h, const StepInput& in, double* out) {
    h->last_host = false;
    h->field_stale = h->fuse_place && out != nullptr;    // the placement is folded into the grid->cell gather
    if (h->ev_valid) cudaEventRecord(h->ev[0], h->stream);
    if (!h->use_graphs || h->ev_valid || h->eager_steps > 0) {   // per-stage events: eager launches (event nodes of a graph carry no usable timestamps)
        if (h->eager_steps > 0) --h->eager_steps;
        TRY(run_step(h, in, out));
    } else {
        const void* k1 = in.U ? (const void*)in.U : (const void*)in.rows;
        psm_handle::StepGraph* g = nullptr;
        for (auto& c : h->graphs)
            if (c.exec && c.in == k1 && c.in2 == in.dU && c.in3 == in.p_dev && c.out == out && c.stride == in.u_stride) g = &c;
        if (!g) {
            g = &h->graphs[0];
            for (auto& c : h->graphs) if (!c.exec) { g = &c; break; } else if (c.used < g->used) g = &c;
            if (g->exec) { cudaGraphExecDestroy(g->exec); g->exec = nullptr; }
            cudaGraph_t graph = nullptr;
            CU(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
            int rc = run_step(h, in, out);
            cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
            if (rc == PSM_OK && e == cudaSuccess) e = cudaGraphInstantiate(&g->exec, graph, 0);
            if (graph) cudaGraphDestroy(graph);
            if (rc != PSM_OK || e != cudaSuccess) {
                g->exec = nullptr;
                cudaGetLastError();
                h->use_graphs = false;
                TRY(run_step(h, in, out));
                if (h->ev_valid) cudaEventRecord(h->ev[PSM_N_TIMINGS], h->stream);
                return PSM_OK;
            }
            g->in = k1; g->in2 = in.dU; g->in3 = in.p_dev; g->out = out; g->stride = in.u_stride;
        }
        g->used = ++h->graph_clock;
        CU(h, cudaGraphLaunch(g->exec, h->stream));
    }
    if (h->ev_valid) cudaEventRecord(h->ev[PSM_N_TIMINGS], h->stream);
    return PSM_OK;
}

// PSM_TRACE_LAYERS=1: after the 30th step, the globaltimer stamps of the Dense layers (min / max over the CTAs of a layer, in
// ns since the first layer's first CTA entered) go to stderr once
static void dump_layer_trace(psm_handle* h) {
    std::vector<unsigned long long> tr((size_t)h->n_dense * 64 * 8);
    if (cudaMemcpy(tr.data(), h->d_layer_trace, tr.size() * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return;
    unsigned long long t0 = ~0ull;
    for (size_t i = 0; i < 64; ++i) if (tr[i * 8]) t0 = std::min(t0, tr[i * 8]);
    static const char* nm[7] = {"entry", "wait_done", "operands", "accum", "pushed", "cluster_bar", "stored"};
    for (int l = 0; l < h->n_dense; ++l) {
        fprintf(stderr, "dense layer %d:", l);
        for (int k = 0; k < 7; ++k) {
            unsigned long long lo = ~0ull, hi = 0;
            for (int c = 0; c < 64; ++c) {
                const unsigned long long v = tr[((size_t)l * 64 + c) * 8 + k];
                if (!v) continue;
                lo = std::min(lo, v); hi = std::max(hi, v);
            }
            if (hi) fprintf(stderr, "  %s %lld..%lld", nm[k], (long long)(lo - t0), (long long)(hi - t0));
        }
        fprintf(stderr, "\n");
    }
}

static int finish(psm_handle* h) {
    CU(h, cudaStreamSynchronize(h->stream));
    if (h->d_layer_trace && ++h->trace_steps == 30) dump_layer_trace(h);
    const int v = h->h_sc->skip;        // skip | comm_error << 8, written by offsets_kernel through mapped memory
    if (v >> 8) PSM_FAIL(h, PSM_ERR_COMM, "a device-side wait timed out (peer-memory exchange: ranks out of step? / Dense-stack grid barrier)");
    return (v & 1) ? PSM_SKIPPED : PSM_OK;
}

// Host entry points launch EAGERLY: the CPU enqueues the eleven kernels while the DMA engine is still copying the inputs
// (hundreds of microseconds), so nothing is gained by a graph -- and nothing depends on the caller re-using its buffers
// (a Python caller may hand in a fresh temporary every step).
extern "C" int psm_predict(psm_handle* h, const double* cells, int64_t n_cells, double* p_out) {
    if (!h) return PSM_ERR_INVALID;
    if (!h->initialised) PSM_FAIL(h, PSM_ERR_STATE, "psm_predict before psm_init_with_tables");
    if (!cells || !p_out) PSM_FAIL(h, PSM_ERR_INVALID, "NULL buffer");
    if (n_cells != h->n_cells) PSM_FAIL(h, PSM_ERR_INVALID, "n_cells %lld does not match the initialised mesh (%lld)", (long long)n_cells, h->n_cells);
    if (!h->have_back) PSM_FAIL(h, PSM_ERR_STATE, "no grid->cell tables were given: use psm_predict_device(..., NULL, ...) + psm_get_stage(FIELD)");
    CU(h, cudaSetDevice(h->cfg.device));
    h->last_host = true;
    h->field_stale = h->fuse_place;
    if (h->ev_valid) cudaEventRecord(h->ev[0], h->stream);
    TRY(upload_split(h, h->d_cells, cells, (size_t)h->n_cells * h->cfg.input_cols));
    if (h->ev_valid) cudaEventRecord(h->ev[11], h->stream);
    StepInput in; in.rows = h->d_cells;
    in.host_out = p_out; in.tail_chunks = tail_chunks_for(h);       // the pressures go down chunk by chunk behind their gather
    if (h->eager_steps > 0) --h->eager_steps;
    TRY(run_step(h, in, h->d_out));
    if (h->ev_valid) cudaEventRecord(h->ev[PSM_N_TIMINGS], h->stream);
    return finish(h);
}

static int check_fields(psm_handle* h, const void* U, int32_t u_stride, const void* dU, int64_t n_cells, const char* who) {
    if (!h->initialised) PSM_FAIL(h, PSM_ERR_STATE, "%s before init", who);
    if (!U) PSM_FAIL(h, PSM_ERR_INVALID, "%s: NULL U", who);
    if (u_stride != 2 && u_stride != 3) PSM_FAIL(h, PSM_ERR_INVALID, "%s: u_stride must be 3 (OpenFOAM vector) or 2", who);
    if (n_cells != h->n_cells) PSM_FAIL(h, PSM_ERR_INVALID, "n_cells %lld does not match the initialised mesh (%lld)", (long long)n_cells, h->n_cells);
    if (dU && h->cfg.variant != PSM_DELTAU_TO_DELTAP) PSM_FAIL(h, PSM_ERR_INVALID, "%s: dU is only meaningful for deltaU_to_deltaP", who);
    return PSM_OK;
}

extern "C" int psm_predict_fields(psm_handle* h, const double* U, int32_t u_stride, const double* dU, const double* p, int64_t n_cells,
                                  double* out) {
    if (!h) return PSM_ERR_INVALID;
    TRY(check_fields(h, U, u_stride, dU, n_cells, "psm_predict_fields"));
    if (!out) PSM_FAIL(h, PSM_ERR_INVALID, "NULL output buffer");
    if (!h->have_back) PSM_FAIL(h, PSM_ERR_STATE, "no grid->cell tables were given");
    CU(h, cudaSetDevice(h->cfg.device));
    const size_t nu = (size_t)h->n_cells * u_stride;
    h->last_host = true;
    h->field_stale = h->fuse_place;
    if (h->ev_valid) cudaEventRecord(h->ev[0], h->stream);
    TRY(upload_split(h, h->d_cells, U, nu));
    if (dU) TRY(upload_split(h, h->d_cells + nu, dU, nu));
    if (h->ev_valid) cudaEventRecord(h->ev[11], h->stream);
    StepInput in; in.U = h->d_cells; in.dU = dU ? h->d_cells + nu : nullptr; in.u_stride = u_stride;
    if (p) {
        // p is read by the LAST kernel only: its copy starts when U has landed (so the two do not share the link) and runs on
        // the copy stream while the kernels execute
        CU(h, cudaEventRecord(h->ev_um, h->stream));                // U (both halves) has landed
        CU(h, cudaStreamWaitEvent(h->copy_stream, h->ev_um, 0));
        CU(h, cudaStreamWaitEvent(h->copy_stream2, h->ev_um, 0));
        const int nch = tail_chunks_for(h);
        if (nch == 1) {
            CU(h, cudaMemcpyAsync(h->d_pprev, p, (size_t)h->n_cells * sizeof(double), cudaMemcpyHostToDevice, h->copy_stream));
            CU(h, cudaEventRecord(h->ev_p, h->copy_stream));
        } else {
            for (int k = 0; k < nch; ++k) {                          // chunks alternate between the two copy streams
                cudaStream_t cs = (h->split_h2d && (k & 1)) ? h->copy_stream2 : h->copy_stream;
                long long c0, c1; tail_range(h->n_cells, nch, k, &c0, &c1);
                if (c1 > c0) CU(h, cudaMemcpyAsync(h->d_pprev + c0, p + c0, (size_t)(c1 - c0) * sizeof(double), cudaMemcpyHostToDevice, cs));
                CU(h, cudaEventRecord(h->ev_pc[k], cs));
            }
        }
        in.wait_p = true;
        h->pprev_zero = false;
    } else if (!h->pprev_zero) {
        CU(h, cudaMemsetAsync(h->d_pprev, 0, (size_t)h->n_cells * sizeof(double), h->stream));
        h->pprev_zero = true;
    }
    in.host_out = out; in.tail_chunks = tail_chunks_for(h);
    if (h->eager_steps > 0) --h->eager_steps;
    TRY(run_step(h, in, h->d_out));
    if (h->ev_valid) cudaEventRecord(h->ev[PSM_N_TIMINGS], h->stream);
    return finish(h);
}

extern "C" int psm_predict_fields_device(psm_handle* h, const double* d_U, int32_t u_stride, const double* d_dU, const double* d_p,
                                         int64_t n_cells, double* d_out, int32_t sync) {
    if (!h) return PSM_ERR_INVALID;
    TRY(check_fields(h, d_U, u_stride, d_dU, n_cells, "psm_predict_fields_device"));
    if (d_out && !h->have_back) PSM_FAIL(h, PSM_ERR_STATE, "no grid->cell tables were given");
    CU(h, cudaSetDevice(h->cfg.device));
    StepInput in; in.U = d_U; in.dU = d_dU; in.u_stride = u_stride; in.p_dev = d_p;
    if (!d_p && !h->pprev_zero) {
        CU(h, cudaMemsetAsync(h->d_pprev, 0, (size_t)h->n_cells * sizeof(double), h->stream));
        h->pprev_zero = true;
    }
    TRY(submit_device(h, in, d_out));
    if (sync) return finish(h);
    return PSM_OK;
}

extern "C" int psm_predict_device(psm_handle* h, const double* d_cells, int64_t n_cells, double* d_p_out, int32_t sync) {
    if (!h) return PSM_ERR_INVALID;
    if (!h->initialised) PSM_FAIL(h, PSM_ERR_STATE, "psm_predict_device before psm_init_with_tables");
    if (!d_cells) PSM_FAIL(h, PSM_ERR_INVALID, "NULL buffer");
    if (n_cells != h->n_cells) PSM_FAIL(h, PSM_ERR_INVALID, "n_cells does not match the initialised mesh");
    if (d_p_out && !h->have_back) PSM_FAIL(h, PSM_ERR_STATE, "no grid->cell tables were given");
    CU(h, cudaSetDevice(h->cfg.device));
    StepInput in; in.rows = d_cells;
    h->pprev_zero = false;                  // prep writes column 4 into d_pprev (also on every replay of the captured step)
    TRY(submit_device(h, in, d_p_out));
    if (sync) return finish(h);
    return PSM_OK;
}

// ---- cell routing -----------------------------------------------------------------------------------------------
// The reference accepts any domain decomposition (scotch, system/decomposeParDict): every MPI rank hands its cells to rank 0,
// which computes alone and scatters the pressures back (PMP:179-185, 258, 501-511).  Here every rank keeps a GPU and the block
// rows of the plan decide who evaluates what, so a rank's cells are routed: rows go to the rank whose block rows contain the
// cell (grouped ncclSend / ncclRecv over NVLink), the pressures come back the same way.  Static lists, built once.
extern "C" int psm_route_init(psm_handle* h, const psm_route* r) {
    if (!h) return PSM_ERR_INVALID;
    if (!h->initialised) PSM_FAIL(h, PSM_ERR_STATE, "psm_route_init before psm_init_sharded");
    if (!r || r->n_local < 0 || (r->n_local > 0 && (!r->dest_rank || !r->dest_index))) PSM_FAIL(h, PSM_ERR_INVALID, "psm_route_init: bad route");
    if (h->routed) PSM_FAIL(h, PSM_ERR_STATE, "a route is already installed");
    if (h->world > 1 && !h->comm) PSM_FAIL(h, PSM_ERR_STATE, "psm_route_init needs the communicator of psm_comm_init");
    CU(h, cudaSetDevice(h->cfg.device));
    const int Wd = h->world, me = h->rank;
    const long long nl = r->n_local, no = h->n_cells;
    // send side: local cells grouped by destination rank (stable: the order inside a group is the caller's order)
    std::vector<long long> cnt(Wd, 0);
    for (long long i = 0; i < nl; ++i) {
        if (r->dest_rank[i] < 0 || r->dest_rank[i] >= Wd) PSM_FAIL(h, PSM_ERR_INVALID, "dest_rank[%lld] = %d is not a rank", i, r->dest_rank[i]);
        ++cnt[r->dest_rank[i]];
    }
    h->rt_send_ptr.assign(Wd + 1, 0);
    for (int p = 0; p < Wd; ++p) h->rt_send_ptr[p + 1] = h->rt_send_ptr[p] + cnt[p];
    std::vector<int32_t> perm(nl > 0 ? nl : 1), sidx(nl > 0 ? nl : 1);
    {
        std::vector<long long> fill(h->rt_send_ptr.begin(), h->rt_send_ptr.end() - 1);
        for (long long i = 0; i < nl; ++i) { const long long s = fill[r->dest_rank[i]]++; perm[s] = (int32_t)i; sidx[s] = r->dest_index[i]; }
    }
    // counts of every (source, destination) pair
    std::vector<long long> all((size_t)Wd * Wd, 0);
    for (int p = 0; p < Wd; ++p) all[(size_t)me * Wd + p] = cnt[p];
    if (Wd > 1) {
        long long* d_all = nullptr;
        CU(h, cudaMalloc(&d_all, sizeof(long long) * Wd * Wd));
        CU(h, cudaMemcpyAsync(d_all + (size_t)me * Wd, cnt.data(), sizeof(long long) * Wd, cudaMemcpyHostToDevice, h->stream));
        NC(h, g_nccl.AllGather(d_all + (size_t)me * Wd, d_all, sizeof(long long) * Wd, ncclChar, h->comm, h->stream));
        CU(h, cudaMemcpyAsync(all.data(), d_all, sizeof(long long) * Wd * Wd, cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
        cudaFree(d_all);
    }
    h->rt_recv_ptr.assign(Wd + 1, 0);
    for (int p = 0; p < Wd; ++p) h->rt_recv_ptr[p + 1] = h->rt_recv_ptr[p] + all[(size_t)p * Wd + me];
    if (h->rt_recv_ptr[Wd] != no)
        PSM_FAIL(h, PSM_ERR_INVALID, "the ranks route %lld cells to rank %d, its block rows own %lld", h->rt_recv_ptr[Wd], me, no);
    // the receiver learns where every incoming row goes: exchange the dest_index lists along the same routes
    TRY(upload(h, &h->d_rt_perm, perm));
    int32_t* d_sidx = nullptr;
    CU(h, cudaMalloc(&d_sidx, sizeof(int32_t) * (nl > 0 ? nl : 1)));
    CU(h, cudaMemcpyAsync(d_sidx, sidx.data(), sizeof(int32_t) * nl, cudaMemcpyHostToDevice, h->stream));
    TRY(dalloc(h, &h->d_rt_recv_idx, (size_t)(no > 0 ? no : 1)));
    if (Wd > 1) {
        NC(h, g_nccl.GroupStart());
        for (int p = 0; p < Wd; ++p) {
            if (p == me) continue;
            const long long ns = cnt[p], nr = h->rt_recv_ptr[p + 1] - h->rt_recv_ptr[p];
            if (ns > 0) NC(h, g_nccl.Send(d_sidx + h->rt_send_ptr[p], (size_t)ns, ncclInt32, p, h->comm, h->stream));
            if (nr > 0) NC(h, g_nccl.Recv(h->d_rt_recv_idx + h->rt_recv_ptr[p], (size_t)nr, ncclInt32, p, h->comm, h->stream));
        }
        NC(h, g_nccl.GroupEnd());
    }
    CU(h, cudaMemcpyAsync(h->d_rt_recv_idx + h->rt_recv_ptr[me], d_sidx + h->rt_send_ptr[me], sizeof(int32_t) * cnt[me], cudaMemcpyDeviceToDevice, h->stream));
    std::vector<int32_t> ridx(no > 0 ? no : 1);
    CU(h, cudaMemcpyAsync(ridx.data(), h->d_rt_recv_idx, sizeof(int32_t) * no, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    cudaFree(d_sidx);
    {   // every owned cell exactly once
        std::vector<uint8_t> seen(no > 0 ? no : 1, 0);
        for (long long k = 0; k < no; ++k) {
            if (ridx[k] < 0 || ridx[k] >= no || seen[ridx[k]]) PSM_FAIL(h, PSM_ERR_INVALID, "routed rows do not cover the owned cells of rank %d exactly once (row %lld -> %d)", me, k, ridx[k]);
            seen[ridx[k]] = 1;
        }
    }
    const size_t nlz = (size_t)(nl > 0 ? nl : 1), noz = (size_t)(no > 0 ? no : 1);
    TRY(dalloc(h, &h->d_rt_in, nlz * 7));                  // U (stride <= 3) + dU (stride <= 3) + p
    TRY(dalloc(h, &h->d_rt_send, nlz * 5)); TRY(dalloc(h, &h->d_rt_recv, noz * 5));
    TRY(dalloc(h, &h->d_rt_osend, noz * h->F)); TRY(dalloc(h, &h->d_rt_orecv, nlz * h->F)); TRY(dalloc(h, &h->d_rt_out, nlz * h->F));
    CU(h, cudaStreamSynchronize(h->stream));
    h->route_n = nl; h->routed = true;
    return PSM_OK;
}

static int route_exchange(psm_handle* h, const double* send, const std::vector<long long>& sp, double* recv, const std::vector<long long>& rp, int k) {
    const int Wd = h->world, me = h->rank;
    if (Wd > 1) {
        NC(h, g_nccl.GroupStart());
        for (int p = 0; p < Wd; ++p) {
            if (p == me) continue;
            const long long ns = sp[p + 1] - sp[p], nr = rp[p + 1] - rp[p];
            if (ns > 0) NC(h, g_nccl.Send(send + sp[p] * k, (size_t)ns * k, ncclDouble, p, h->comm, h->stream));
            if (nr > 0) NC(h, g_nccl.Recv(recv + rp[p] * k, (size_t)nr * k, ncclDouble, p, h->comm, h->stream));
        }
        NC(h, g_nccl.GroupEnd());
    }
    const long long ns = sp[me + 1] - sp[me];
    if (ns > 0) CU(h, cudaMemcpyAsync(recv + rp[me] * k, send + sp[me] * k, sizeof(double) * ns * k, cudaMemcpyDeviceToDevice, h->stream));
    return PSM_OK;
}

extern "C" int psm_predict_routed(psm_handle* h, const double* U, int32_t u_stride, const double* dU, const double* p, int64_t n_local,
                                  double* out) {
    if (!h) return PSM_ERR_INVALID;
    if (!h->initialised || !h->routed) PSM_FAIL(h, PSM_ERR_STATE, "psm_predict_routed before psm_route_init");
    if (n_local != h->route_n) PSM_FAIL(h, PSM_ERR_INVALID, "n_local %lld does not match the installed route (%lld)", (long long)n_local, h->route_n);
    if (n_local > 0 && (!U || !out)) PSM_FAIL(h, PSM_ERR_INVALID, "NULL buffer");
    if (u_stride != 2 && u_stride != 3) PSM_FAIL(h, PSM_ERR_INVALID, "u_stride must be 3 (OpenFOAM vector) or 2");
    if (dU && h->cfg.variant != PSM_DELTAU_TO_DELTAP) PSM_FAIL(h, PSM_ERR_INVALID, "dU is only meaningful for deltaU_to_deltaP");
    if (!h->have_back) PSM_FAIL(h, PSM_ERR_STATE, "no grid->cell tables were given");
    CU(h, cudaSetDevice(h->cfg.device));
    cudaStream_t s = h->stream;
    const long long nl = n_local, no = h->n_cells;
    const size_t nu = (size_t)nl * u_stride;
    double* dUl = h->d_rt_in; double* ddUl = h->d_rt_in + nu; double* dpl = h->d_rt_in + 2 * nu;
    if (nl > 0) {
        CU(h, cudaMemcpyAsync(dUl, U, nu * sizeof(double), cudaMemcpyHostToDevice, s));
        if (dU) CU(h, cudaMemcpyAsync(ddUl, dU, nu * sizeof(double), cudaMemcpyHostToDevice, s));
        if (p) CU(h, cudaMemcpyAsync(dpl, p, (size_t)nl * sizeof(double), cudaMemcpyHostToDevice, s));
    }
    const int k = 3 + (dU ? 2 : 0);
    launch_route_pack(RoutePackArgs{dUl, dU ? ddUl : nullptr, p ? dpl : nullptr, u_stride, h->d_rt_perm, nl, k, h->d_rt_send}, s);
    TRY(route_exchange(h, h->d_rt_send, h->rt_send_ptr, h->d_rt_recv, h->rt_recv_ptr, k));
    // the routed rows become the handle's native-field input: U as double[n_owned][2], dU behind it, p in the p_prev buffer
    double* U2 = h->d_cells; double* dU2 = h->d_cells + 2 * (size_t)no;
    launch_route_scatter(RouteScatterArgs{h->d_rt_recv, h->d_rt_recv_idx, no, k, dU ? 1 : 0, U2, dU2, h->d_pprev}, s);
    h->pprev_zero = false;
    h->last_host = true; h->field_stale = h->fuse_place;
    StepInput in; in.U = U2; in.dU = dU ? dU2 : nullptr; in.u_stride = 2;
    if (h->eager_steps > 0) --h->eager_steps;
    TRY(run_step(h, in, h->d_out));
    launch_route_back(RouteBackArgs{h->d_out, h->d_rt_recv_idx, no, h->F, h->d_rt_osend, 1}, s);
    TRY(route_exchange(h, h->d_rt_osend, h->rt_recv_ptr, h->d_rt_orecv, h->rt_send_ptr, h->F));
    launch_route_back(RouteBackArgs{h->d_rt_orecv, h->d_rt_perm, nl, h->F, h->d_rt_out, 0}, s);
    if (nl > 0) CU(h, cudaMemcpyAsync(out, h->d_rt_out, (size_t)nl * h->F * sizeof(double), cudaMemcpyDeviceToHost, s));
    return finish(h);
}

extern "C" int psm_synchronize(psm_handle* h) {
    if (!h) return PSM_ERR_INVALID;
    if (!h->initialised) PSM_FAIL(h, PSM_ERR_STATE, "not initialised");
    CU(h, cudaSetDevice(h->cfg.device));
    return finish(h);
}

extern "C" int psm_get_stream(const psm_handle* h, void** stream) {
    if (!h || !stream) return PSM_ERR_INVALID;
    *stream = (void*)h->stream;
    return PSM_OK;
}

extern "C" int psm_register_host_buffer(void* ptr, int64_t bytes) {
    if (!ptr || bytes <= 0) return PSM_ERR_INVALID;
    return cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault) == cudaSuccess ? PSM_OK : PSM_ERR_CUDA;
}
extern "C" int psm_unregister_host_buffer(void* ptr) {
    if (!ptr) return PSM_ERR_INVALID;
    return cudaHostUnregister(ptr) == cudaSuccess ? PSM_OK : PSM_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------------
extern "C" int psm_get_geometry(const psm_handle* h, psm_geometry* g) {
    if (!h || !g) return PSM_ERR_INVALID;
    if (!h->initialised) return PSM_ERR_STATE;
    const Plan& P = h->plan;
    g->grid_h = P.H; g->grid_w = P.W; g->shape = P.S; g->overlap = P.ov; g->n_x = P.n_x; g->n_y = P.n_y;
    g->p_i = P.p_i; g->p_j = P.p_j; g->n_blocks = P.B; g->n_fields = P.F; g->n_cells = h->n_cells;
    g->n_tasks = (int32_t)P.tasks.size(); g->peer_memory_exchange = h->p2p ? 1 : 0;
    g->row0 = h->row0; g->row1 = h->row1; g->ext_rows = h->ext_rows + h->local_ext; g->first_block = h->kb0; g->n_local_blocks = h->B;
    g->world = h->world; g->n_ghost_cells = h->n_ghost; g->n_ghost_pix = h->n_ghost_pix;
    return PSM_OK;
}

extern "C" int psm_get_plan(const psm_handle* h, int32_t* origins, int32_t* indices_list) {
    if (!h) return PSM_ERR_INVALID;
    if (!h->initialised) return PSM_ERR_STATE;
    for (int k = 0; k < h->plan.B; ++k) {
        if (origins) { origins[2 * k] = h->plan.y0[k]; origins[2 * k + 1] = h->plan.x0[k]; }
        if (indices_list) { indices_list[2 * k] = h->plan.idx_i[k]; indices_list[2 * k + 1] = h->plan.idx_j[k]; }
    }
    return PSM_OK;
}

extern "C" int psm_get_owner_map(const psm_handle* h, int32_t* owner) {
    if (!h || !owner) return PSM_ERR_INVALID;
    if (!h->initialised) return PSM_ERR_STATE;
    memcpy(owner, h->plan.owner.data(), h->plan.owner.size() * sizeof(int32_t));
    return PSM_OK;
}

extern "C" int psm_get_forward_table(const psm_handle* hc, int32_t* vert, float* weights) {
    psm_handle* h = const_cast<psm_handle*>(hc);
    if (!h) return PSM_ERR_INVALID;
    if (!h->initialised) return PSM_ERR_STATE;
    CU(h, cudaSetDevice(h->cfg.device));
    std::vector<int32_t> v(h->G); std::vector<float> w(h->G);
    for (int j = 0; j < 3; ++j) {
        CU(h, cudaMemcpy(v.data(), h->d_fv[j], h->G * sizeof(int32_t), cudaMemcpyDeviceToHost));
        CU(h, cudaMemcpy(w.data(), h->d_fw[j], h->G * sizeof(float), cudaMemcpyDeviceToHost));
        for (long long q = 0; q < h->G; ++q) {
            if (vert) vert[3 * q + j] = v[q];
            if (weights) weights[3 * q + j] = w[q];
        }
    }
    return PSM_OK;
}

extern "C" int psm_get_stage(psm_handle* h, int32_t stage, void* out, int64_t n_bytes) {
    if (!h || !out) return PSM_ERR_INVALID;
    if (!h->initialised) PSM_FAIL(h, PSM_ERR_STATE, "not initialised");
    CU(h, cudaSetDevice(h->cfg.device));
    CU(h, cudaStreamSynchronize(h->stream));
    const int S2 = h->S * h->S;
    auto need = [&](int64_t want) -> int {
        if (n_bytes != want) PSM_FAIL(h, PSM_ERR_INVALID, "stage %d needs %lld bytes, got %lld", stage, (long long)want, (long long)n_bytes);
        return 0;
    };
    switch (stage) {
        case PSM_STAGE_GRID: {
            const int64_t gl = (int64_t)(h->H + h->local_ext + h->ext_rows) * h->W;
            TRY(need(2 * gl * 4));
            if (h->fused_extract && !h->keep_grid && !h->grid_a) {
                // the fused gather writes only the block operand: rebuild the planes of the LAST step from the
                // still-resident cell field and the scales that step published
                ScalarArgs sa{h->d_sc, h->maxs[0], h->maxs[1], 1.0, 0, 0.0, 0, 1};
                GatherArgs ga{h->d_fv[0], h->d_fv[1], h->d_fv[2], h->d_fw[0], h->d_fw[1], h->d_fw[2], h->d_uv,
                              h->d_grid, h->d_grid + h->grid_stride, h->G_pad / 4, sa, 1, nullptr};
                launch_gather(ga, h->stream);
                CU(h, cudaStreamSynchronize(h->stream));
            }
            CU(h, cudaMemcpy(out, h->d_grid, gl * 4, cudaMemcpyDeviceToHost));
            CU(h, cudaMemcpy((char*)out + gl * 4, h->d_grid + h->grid_stride, gl * 4, cudaMemcpyDeviceToHost));
            return PSM_OK;
        }
        case PSM_STAGE_XINPUT:
            TRY(need((int64_t)h->B * h->pc_in * 4));
            CU(h, cudaMemcpy2D(out, (size_t)h->pc_in * 4, h->d_xin, (size_t)h->pc_in_pad * 4, (size_t)h->pc_in * 4, h->B, cudaMemcpyDeviceToHost));
            return PSM_OK;
        case PSM_STAGE_MLPOUT:
            TRY(need((int64_t)h->B * h->pc_p * 4));
            CU(h, cudaMemcpy2D(out, (size_t)h->pc_p * 4, h->d_r, (size_t)h->pc_p_pad * 4, (size_t)h->pc_p * 4, h->B, cudaMemcpyDeviceToHost));
            return PSM_OK;
        case PSM_STAGE_BLOCKS:
            TRY(need((int64_t)h->B * h->C * S2 * 4));
            CU(h, cudaMemcpy(out, h->d_blocks, (size_t)h->B * h->C * S2 * 4, cudaMemcpyDeviceToHost));
            return PSM_OK;
        case PSM_STAGE_OFFSETS:
            TRY(need((int64_t)h->F * h->Bg * 8));
            CU(h, cudaMemcpy(out, h->d_offsets, (size_t)h->F * h->Bg * 8, cudaMemcpyDeviceToHost));
            return PSM_OK;
        case PSM_STAGE_FIELD:
            TRY(need((int64_t)h->F * h->G * 4));
            if (h->field_stale) {
                // the step folded the placement into the grid->cell gather: materialise the field of the LAST step now
                PlaceArgs pl{h->d_blocks, h->d_owner, h->d_by0, h->d_bx0, h->d_coff, h->d_field, h->Bg, h->kb0, h->C, h->F, h->S, h->H, h->W,
                             h->field_stride};
                launch_place(pl, h->stream);
                CU(h, cudaStreamSynchronize(h->stream));
                h->field_stale = false;
            }
            CU(h, cudaMemcpy2D(out, (size_t)h->G * 4, h->d_field, (size_t)h->field_stride * 4, (size_t)h->G * 4, h->F, cudaMemcpyDeviceToHost));
            return PSM_OK;
        case PSM_STAGE_SCALARS: {
            TRY(need(4 * 8));
            Scalars sc;
            CU(h, cudaMemcpy(&sc, h->d_sc, sizeof sc, cudaMemcpyDeviceToHost));
            double* o = (double*)out;
            o[0] = sc.U_max_norm; o[1] = sc.dU_max_norm; o[2] = sc.shift[0]; o[3] = sc.shift[1];
            return PSM_OK;
        }
        case PSM_STAGE_XU:
            TRY(need((int64_t)h->B * 2 * S2 * 4));
            if (h->grid_a) {     // the projection read the grid planes itself: build the block operand of the last step on demand
                ExtractArgs ea{h->d_grid, h->d_grid + h->grid_stride, h->d_by0, h->d_bx0, h->d_xu, h->B, h->W, h->S, 2};
                launch_extract(ea, h->stream);
                CU(h, cudaStreamSynchronize(h->stream));
            }
            CU(h, cudaMemcpy(out, h->d_xu, (size_t)h->B * 2 * S2 * 4, cudaMemcpyDeviceToHost));
            return PSM_OK;
        case PSM_STAGE_MEANS:
            TRY(need((int64_t)h->plan.tasks.size() * 8));
            CU(h, cudaMemcpy(out, h->d_means, h->plan.tasks.size() * 8, cudaMemcpyDeviceToHost));
            return PSM_OK;
        default:
            PSM_FAIL(h, PSM_ERR_INVALID, "unknown stage %d", stage);
    }
}

// U_to_gradP: the pressure field recovered from the assembled gradient fields of the LAST step (GRAD:371-416, 585-628).
extern "C" int psm_integrate_gradp(psm_handle* h, const psm_integrate_geometry* g, double* p_field) {
    if (!h) return PSM_ERR_INVALID;
    if (!h->initialised) PSM_FAIL(h, PSM_ERR_STATE, "not initialised");
    if (!g || !p_field) PSM_FAIL(h, PSM_ERR_INVALID, "psm_integrate_gradp: NULL argument");
    if (h->cfg.variant != PSM_U_TO_GRADP || h->world != 1) PSM_FAIL(h, PSM_ERR_STATE, "psm_integrate_gradp needs a single-GPU U_to_gradP handle");
    const int H = h->Hglob, W = h->W, cy = g->center_row;
    if (cy < 1 || cy >= H) PSM_FAIL(h, PSM_ERR_GEOMETRY, "center_row %d is outside the %d grid rows (the reference hard-codes 200, GRAD:592)", cy, H);
    CU(h, cudaSetDevice(h->cfg.device));
    CU(h, cudaStreamSynchronize(h->stream));
    // centre column: middle of the zero run (obstacle) of the distance field in the centre row, GRAD:591
    const double sx = (g->max_x - g->min_x) / (double)(W - 1);
    double lo = 0.0, hi = 0.0; bool any = false;
    for (int x = 0; x < W; ++x)
        if (!h->plan_mask_at(cy, x)) { const double xv = (x == W - 1) ? g->max_x : g->min_x + sx * x; if (!any) { lo = hi = xv; any = true; } else { if (xv < lo) lo = xv; if (xv > hi) hi = xv; } }
    if (!any) PSM_FAIL(h, PSM_ERR_GEOMETRY, "row %d of the distance field has no zero (the obstacle must cross the centre row, GRAD:591)", cy);
    const int cx = (int)(((hi + lo) / 2 - g->x0_min) / h->cfg.delta);
    if (cx < 1 || cx >= W) PSM_FAIL(h, PSM_ERR_GEOMETRY, "centre column %d outside the grid", cx);
    // the static fix-up list of every block-local row (GRAD:389-395 with nn = int(sdfunct[i_local, :]))
    const int nloc = cy > H - cy ? cy : H - cy;
    const int wmin = cx < W - cx + 1 ? cx : W - cx + 1;
    std::vector<IntegrateFix> fix(nloc);
    for (int i = 0; i < nloc; ++i) {
        IntegrateFix f{};
        const uint8_t* nn = h->sdf_int.data() + (size_t)i * W;
        int last[256]; for (int v = 0; v < 256; ++v) last[v] = -1;
        for (int k = 0; k < W; ++k) last[nn[k]] = k;
        for (int v = 0; v < 256; ++v) {
            if (last[v] < 0) continue;
            if (v >= wmin) PSM_FAIL(h, PSM_ERR_GEOMETRY, "distance %d m indexes past a quadrant of width %d (the reference raises IndexError, GRAD:393)", v, wmin);
            if (f.n >= 4) PSM_FAIL(h, PSM_ERR_GEOMETRY, "more than 4 distinct integer distances in one grid row");
            f.pos[f.n] = v; f.prev[f.n] = last[v] > 0 ? nn[last[v] - 1] : -1; ++f.n;
        }
        fix[i] = f;
    }
    if (h->field_stale) {
        PlaceArgs pl{h->d_blocks, h->d_owner, h->d_by0, h->d_bx0, h->d_coff, h->d_field, h->Bg, h->kb0, h->C, h->F, h->S, h->H, h->W, h->field_stride};
        launch_place(pl, h->stream);
        h->field_stale = false;
    }
    IntegrateFix* d_fix = nullptr; double *d_sdpx = nullptr, *d_anchor = nullptr, *d_corr = nullptr, *d_out = nullptr; int* d_status = nullptr;
    const size_t G = (size_t)H * W;
    int rc = PSM_OK;
    if (cudaMalloc(&d_fix, fix.size() * sizeof(IntegrateFix)) != cudaSuccess || cudaMalloc(&d_sdpx, 2 * G * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&d_anchor, 4 * (size_t)H * sizeof(double)) != cudaSuccess || cudaMalloc(&d_corr, 2 * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&d_out, G * sizeof(double)) != cudaSuccess || cudaMalloc(&d_status, sizeof(int)) != cudaSuccess) rc = PSM_ERR_CUDA;
    int status = 0;
    if (rc == PSM_OK) {
        cudaMemcpyAsync(d_fix, fix.data(), fix.size() * sizeof(IntegrateFix), cudaMemcpyHostToDevice, h->stream);
        cudaMemsetAsync(d_status, 0, sizeof(int), h->stream);
        cudaMemsetAsync(d_anchor, 0, 4 * (size_t)H * sizeof(double), h->stream);
        // np.diff(np.linspace(a, b, n))[0] = (1 * step + a) - a, rounding included
        const double sy = (g->max_y - g->min_y) / (double)(H - 1);
        volatile double x1 = 1.0 * sx + g->min_x, y1 = 1.0 * sy + g->min_y;
        const double dx = x1 - g->min_x, dy = y1 - g->min_y;
        IntegrateArgs ia{h->d_field, h->d_field + h->field_stride, h->d_gmask, d_fix, H, W, cx, cy, dx, dy,
                         d_sdpx, d_anchor, d_corr, d_status, d_out};
        launch_integrate(ia, h->stream);
        cudaMemcpyAsync(p_field, d_out, G * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
        cudaMemcpyAsync(&status, d_status, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
        if (cudaStreamSynchronize(h->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) rc = PSM_ERR_CUDA;
    }
    cudaFree(d_fix); cudaFree(d_sdpx); cudaFree(d_anchor); cudaFree(d_corr); cudaFree(d_out); cudaFree(d_status);
    if (rc != PSM_OK) PSM_FAIL(h, rc, "psm_integrate_gradp: CUDA failure");
    if (status) PSM_FAIL(h, PSM_ERR_GEOMETRY, "the two stitch columns have different numbers of flow pixels (numpy raises at GRAD:606)");
    return PSM_OK;
}

extern "C" int psm_get_timings(psm_handle* h, float* ms, int32_t n) {
    if (!h || !ms) return PSM_ERR_INVALID;
    if (!h->ev_valid) PSM_FAIL(h, PSM_ERR_STATE, "timings were not enabled in psm_config");
    CU(h, cudaSetDevice(h->cfg.device));
    CU(h, cudaStreamSynchronize(h->stream));
    // events: ev[0] start, ev[11] after the H2D copy (host entry point only), ev[1..10] after
    // prep .. back_gather, ev[12] end (after the D2H copy)
    for (int i = 0; i < n && i < PSM_N_TIMINGS; ++i) ms[i] = 0.f;
    float t = 0.f;
    if (h->last_host && cudaEventElapsedTime(&t, h->ev[0], h->ev[11]) == cudaSuccess) ms[0] = t;
    if (n > 1 && cudaEventElapsedTime(&t, h->ev[h->last_host ? 11 : 0], h->ev[1]) == cudaSuccess) ms[1] = t;
    for (int i = 2; i <= 10 && i < n; ++i)
        if (cudaEventElapsedTime(&t, h->ev[i - 1], h->ev[i]) == cudaSuccess) ms[i] = t;
    if (n > 11 && cudaEventElapsedTime(&t, h->ev[10], h->ev[PSM_N_TIMINGS]) == cudaSuccess) ms[11] = t;
    return PSM_OK;
}

extern "C" int psm_set_timings(psm_handle* h, int32_t on) {
    if (!h) return PSM_ERR_INVALID;
    CU(h, cudaSetDevice(h->cfg.device));
    if (on && !h->ev_created) {
        for (auto& e : h->ev) CU(h, cudaEventCreate(&e));
        h->ev_created = true;
    }
    h->ev_valid = on != 0;
    return PSM_OK;
}

extern "C" int psm_get_launch_count(const psm_handle* h) { return h ? h->launches : 0; }

// ---- host-only test entries (no GPU): the box plan of the grid-plane A operand and the ghost-cell send map -------------------
extern "C" int psm_grid_operand_plan(int32_t n_blocks, const int32_t* by0, const int32_t* bx0, int32_t stride, int32_t* tiles, int32_t* gx,
                                     int32_t* gy, int32_t* n_segs, int32_t* row_src, int32_t* segs, int32_t max_segs) {
    if (n_blocks < 1 || !by0 || !bx0 || stride < 1 || !tiles || !gx || !gy || !n_segs) return PSM_ERR_INVALID;
    GridAPlan P;
    std::vector<int32_t> y(by0, by0 + n_blocks), x(bx0, bx0 + n_blocks);
    if (!build_grid_a(n_blocks, y, x, stride, P)) { *tiles = 0; *gx = *gy = 0; *n_segs = 0; return PSM_OK; }    // layout does not pay: 0 tiles
    *tiles = P.tiles; *gx = P.gx; *gy = P.gy; *n_segs = (int32_t)P.segs.size();
    if (row_src) for (int b = 0; b < n_blocks; ++b) row_src[b] = P.row_src[b];
    if (segs) {
        if ((int32_t)P.segs.size() > max_segs) return PSM_ERR_INVALID;
        size_t t = 0;
        for (size_t q = 0; q < P.segs.size(); ++q) {
            while ((int32_t)q >= P.seg_ptr[t + 1]) ++t;
            segs[5 * q + 0] = (int32_t)t; segs[5 * q + 1] = P.segs[q].map; segs[5 * q + 2] = P.segs[q].row;
            segs[5 * q + 3] = P.segs[q].x; segs[5 * q + 4] = P.segs[q].y;
        }
    }
    return PSM_OK;
}

extern "C" int psm_send_map_build(int64_t n_cells, int32_t world, const int64_t* send_ptr, const int32_t* send_idx, uint32_t* words,
                                  int32_t* entries, int64_t* n_entries) {
    if (n_cells < 0 || world < 1 || world > kMaxPeers || !send_ptr || !n_entries) return PSM_ERR_INVALID;
    std::vector<long long> sp(send_ptr, send_ptr + world + 1);
    if (sp[world] >= (1ll << 22)) return PSM_ERR_INVALID;         // entry indices are packed into 22 bits (the handle then keeps the run list)
    for (long long e = 0; e < sp[world]; ++e) if (!send_idx || send_idx[e] < 0 || send_idx[e] >= n_cells) return PSM_ERR_INVALID;
    std::vector<uint2> w; std::vector<int2> en;
    build_send_map(n_cells, world, sp.data(), send_idx, w, en);
    *n_entries = sp[world];
    if (words) for (size_t i = 0; i < w.size(); ++i) { words[2 * i] = w[i].x; words[2 * i + 1] = w[i].y; }
    if (entries) for (long long i = 0; i < sp[world]; ++i) { entries[2 * i] = en[i].x; entries[2 * i + 1] = en[i].y; }
    return PSM_OK;
}

extern "C" int psm_get_wait_ns(psm_handle* h, uint64_t ns[3], uint32_t count[3], int32_t reset) {
    if (!h || !ns || !count) return PSM_ERR_INVALID;
    if (!h->initialised) PSM_FAIL(h, PSM_ERR_STATE, "not initialised");
    CU(h, cudaSetDevice(h->cfg.device));
    CU(h, cudaStreamSynchronize(h->stream));
    Scalars sc;
    CU(h, cudaMemcpy(&sc, h->d_sc, sizeof(Scalars), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 3; ++i) { ns[i] = sc.wait_ns[i]; count[i] = sc.wait_n[i]; }
    if (reset) {
        CU(h, cudaMemset(reinterpret_cast<char*>(h->d_sc) + offsetof(Scalars, wait_ns), 0, sizeof(sc.wait_ns)));
        CU(h, cudaMemset(reinterpret_cast<char*>(h->d_sc) + offsetof(Scalars, wait_n), 0, sizeof(sc.wait_n)));
    }
    return PSM_OK;
}

// ------------------------------------------------------------------------------------------------
extern "C" int psm_debug_gemm(int32_t device, int32_t mode, int32_t M, int32_t N, int32_t K, const float* A, const float* B,
                              float* C, int32_t splits) {
    if (!A || !B || !C || M % 128 || N % 64 || K % 32 || splits < 1 || mode < 0 || mode > 2) return PSM_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return PSM_ERR_CUDA;
    const int kb_total = K / 32;
    const int per = (kb_total + splits - 1) / splits;
    if ((kb_total + per - 1) / per != splits) return PSM_ERR_INVALID;     // every split must own >= 1 k-block
    float *dA = nullptr, *dB = nullptr, *dC = nullptr;
    int rc = PSM_OK;
    const size_t nA = (size_t)M * K, nB = (size_t)N * K, nC = (size_t)splits * M * N;
    if (cudaMalloc(&dA, nA * 4) != cudaSuccess || cudaMalloc(&dB, nB * 4) != cudaSuccess || cudaMalloc(&dC, nC * 4) != cudaSuccess) rc = PSM_ERR_CUDA;
    if (rc == PSM_OK) {
        cudaMemcpy(dA, A, nA * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, B, nB * 4, cudaMemcpyHostToDevice);
        cudaMemset(dC, 0xFF, nC * 4);                                     // NaN pattern: unwritten outputs are caught
        if (mode == PSM_GEMM_FP32_SIMT) {
            GemmArgs g{};
            g.A = dA; g.B = dB; g.C = dC; g.M = M; g.N = N; g.K = K; g.lda = K; g.ldb = K; g.ldc = N; g.splits = splits;
            g.epi = EPI_PARTIAL;
            launch_sgemm(g, 0);
        } else {
            TcGemm t{};
            if (tc_gemm_prepare() != 0 || make_kmajor_map(&t.mapA, dA, M, K, K, 128) != 0 ||
                make_kmajor_map(&t.mapB, dB, N, K, K, tc_gemm_bn(N)) != 0) rc = PSM_ERR_CUDA;
            else {
                t.args = TcGemmArgs{dC, M, N, K, N, splits, EPI_PARTIAL, mode == PSM_GEMM_TC_3XTF32 ? (env_on("PSM_TF32_MASK_HI") ? 1 : 2) : 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 1};
                t.bn = tc_gemm_bn(N);
                launch_tc_gemm(t, 0);
            }
        }
        if (rc == PSM_OK && (cudaDeviceSynchronize() != cudaSuccess || cudaGetLastError() != cudaSuccess)) rc = PSM_ERR_CUDA;
        if (rc == PSM_OK) cudaMemcpy(C, dC, nC * 4, cudaMemcpyDeviceToHost);
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
    return rc;
}

// Unit-test entry of the fused Dense stack: out[M][dims[n]] = Dense(linear)(relu(...relu(x W0 + b0)...)), Keras layout
// kernels[l][in][out].  M is padded to 128 and the widths to 128 internally, like the handle does.
extern "C" int psm_debug_dense_stack(int32_t device, int32_t mode, int32_t M, int32_t n_layers, const int32_t* dims,
                                     const float* const* kernels, const float* const* biases, const float* x, float* out,
                                     int32_t clusters) {
    if (!dims || !kernels || !biases || !x || !out || M < 1 || n_layers < 1 || n_layers > kMaxDense || mode < 0 || mode > 1) return PSM_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return PSM_ERR_CUDA;
    int max_cl = 0;
    if (tc_gemm_prepare() != 0 || dense_stack_prepare(&max_cl) != 0) return PSM_ERR_CUDA;
    const int Mp = round_up(M, 128);
    std::vector<int> dp(n_layers + 1);
    int maxw = 0;
    for (int i = 0; i <= n_layers; ++i) { if (dims[i] < 1) return PSM_ERR_INVALID; dp[i] = round_up(dims[i], 128); maxw = std::max(maxw, dp[i]); }
    std::vector<void*> dev;
    auto dmal = [&](size_t n) -> float* { void* q = nullptr; if (cudaMalloc(&q, n * 4) != cudaSuccess) return nullptr; cudaMemset(q, 0, n * 4); dev.push_back(q); return (float*)q; };
    auto split = [](const std::vector<float>& v, std::vector<float>& hi, std::vector<float>& lo) {
        hi.resize(v.size()); lo.resize(v.size());
        for (size_t i = 0; i < v.size(); ++i) { uint32_t u; memcpy(&u, &v[i], 4); u &= 0xFFFFE000u; memcpy(&hi[i], &u, 4); lo[i] = v[i] - hi[i]; }
    };
    int rc = PSM_OK;
    float* act_hi[2] = {dmal((size_t)Mp * maxw), dmal((size_t)Mp * maxw)};
    float* act_lo[2] = {dmal((size_t)Mp * maxw), dmal((size_t)Mp * maxw)};
    float* d_out = dmal((size_t)Mp * dp[n_layers]);
    float* d_one = dmal(maxw); float* d_zero = dmal(maxw);
    unsigned int* d_bar = (unsigned int*)dmal(2);
    TensorMap128* d_maps = (TensorMap128*)dmal((size_t)4 * n_layers * sizeof(TensorMap128) / 4);
    if (!act_hi[0] || !act_hi[1] || !act_lo[0] || !act_lo[1] || !d_out || !d_one || !d_zero || !d_bar || !d_maps) rc = PSM_ERR_CUDA;
    DenseStackArgs sa{};
    std::vector<TensorMap128> maps((size_t)4 * n_layers);
    int max_tiles = 1;
    if (rc == PSM_OK) {
        std::vector<float> ones(maxw, 1.f), xp((size_t)Mp * dp[0], 0.f), hi, lo;
        cudaMemcpy(d_one, ones.data(), maxw * 4, cudaMemcpyHostToDevice);
        for (int m = 0; m < M; ++m) for (int k = 0; k < dims[0]; ++k) xp[(size_t)m * dp[0] + k] = x[(size_t)m * dims[0] + k];
        split(xp, hi, lo);
        cudaMemcpy(act_hi[0], hi.data(), hi.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(act_lo[0], lo.data(), lo.size() * 4, cudaMemcpyHostToDevice);
        for (int l = 0; l < n_layers && rc == PSM_OK; ++l) {
            const int K = dp[l], N = dp[l + 1];
            std::vector<float> w((size_t)N * K, 0.f), b(N, 0.f);
            for (int i = 0; i < dims[l]; ++i) for (int o = 0; o < dims[l + 1]; ++o) w[(size_t)o * K + i] = kernels[l][(size_t)i * dims[l + 1] + o];
            for (int o = 0; o < dims[l + 1]; ++o) b[o] = biases[l][o];
            split(w, hi, lo);
            float* dh = dmal(w.size()); float* dl = dmal(w.size()); float* db = dmal(N);
            if (!dh || !dl || !db) { rc = PSM_ERR_CUDA; break; }
            cudaMemcpy(dh, hi.data(), hi.size() * 4, cudaMemcpyHostToDevice);
            cudaMemcpy(dl, lo.data(), lo.size() * 4, cudaMemcpyHostToDevice);
            cudaMemcpy(db, b.data(), N * 4, cudaMemcpyHostToDevice);
            const bool last = (l == n_layers - 1);
            if (make_kmajor_map(&maps[4 * l + 0], act_hi[l & 1], Mp, K, K, 128) != 0 || make_kmajor_map(&maps[4 * l + 1], act_lo[l & 1], Mp, K, K, 128) != 0 ||
                make_kmajor_map(&maps[4 * l + 2], dh, N, K, K, 64) != 0 || make_kmajor_map(&maps[4 * l + 3], dl, N, K, K, 64) != 0) { rc = PSM_ERR_CUDA; break; }
            sa.L[l] = DenseLayerDesc{K, N, last ? EPI_BIAS_AFFINE : EPI_BIAS_RELU, db, d_one, d_zero, last ? d_out : nullptr,
                                     last ? nullptr : act_hi[(l + 1) & 1], last ? nullptr : act_lo[(l + 1) & 1]};
            max_tiles = std::max(max_tiles, (Mp / 128) * (N / 64));
        }
    }
    if (rc == PSM_OK) {
        cudaMemcpy(d_maps, maps.data(), maps.size() * sizeof(TensorMap128), cudaMemcpyHostToDevice);
        sa.maps = d_maps; sa.n_layers = n_layers; sa.M = Mp; sa.three_pass = (mode == PSM_GEMM_TC_3XTF32) ? 1 : 0;
        sa.barrier = d_bar; sa.error = (int*)(d_bar + 1);
        int cl = std::min(std::min(max_cl, 16), max_tiles);
        if (clusters > 0) cl = std::min(cl, clusters);
        const bool trace = getenv("PSM_TRACE_DENSE") != nullptr;
        if (trace) sa.trace = (unsigned long long*)dmal((size_t)cl * 8 * 64 * 2);
        for (int rep = 0; rep < (trace ? 3 : 1) && rc == PSM_OK; ++rep) {
            cudaMemset(d_bar, 0, 8);
            if (launch_dense_stack(sa, cl, 0) != 0 || cudaDeviceSynchronize() != cudaSuccess || cudaGetLastError() != cudaSuccess) rc = PSM_ERR_CUDA;
        }
        if (trace && rc == PSM_OK) {
            std::vector<unsigned long long> tr((size_t)cl * 8 * 64);
            cudaMemcpy(tr.data(), sa.trace, tr.size() * 8, cudaMemcpyDeviceToHost);
            unsigned long long t0 = ~0ull;
            for (int b = 0; b < cl * 8; ++b) if (tr[(size_t)b * 64] && tr[(size_t)b * 64] < t0) t0 = tr[(size_t)b * 64];
            for (int b : {0, 1, 7, 8, cl * 8 - 1}) {
                fprintf(stderr, "dense_stack trace CTA %3d (ns since first CTA start):", b);
                for (int i = 0; i < 64 && tr[(size_t)b * 64 + i]; ++i) fprintf(stderr, " %llu", tr[(size_t)b * 64 + i] - t0);
                fprintf(stderr, "\n");
            }
        }
    }
    if (rc == PSM_OK) {
        unsigned int st[2];
        cudaMemcpy(st, d_bar, 8, cudaMemcpyDeviceToHost);
        if (st[1]) rc = PSM_ERR_COMM;
        std::vector<float> o((size_t)Mp * dp[n_layers]);
        cudaMemcpy(o.data(), d_out, o.size() * 4, cudaMemcpyDeviceToHost);
        for (int m = 0; m < M; ++m) for (int k = 0; k < dims[n_layers]; ++k) out[(size_t)m * dims[n_layers] + k] = o[(size_t)m * dp[n_layers] + k];
    }
    for (void* q : dev) cudaFree(q);
    return rc;
}
