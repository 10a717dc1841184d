// Hot-path kernels for sm_100a.  Reference lines each stage replaces are cited per kernel
// (aliases as in include/psm_b200.h).  All HBM-bound kernels use 128-bit accesses over SoA
// tables and grids sized in multiples of the SM count.
#include "psm_kernels.cuh"

#include <math_constants.h>

#include <cstdint>
#include <cstdlib>

namespace psm {

static constexpr int kSMs = 148;
static constexpr int kOffsetsSmemMax = 8192;      // F*B scalars whose recurrence runs in shared memory (192 KB)

bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("PSM_NO_PDL"); return !(e && e[0] == '1'); }();
    return on;
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------
// Peer-memory exchange primitives (see psm_kernels.cuh).
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Thread 0 of the CTA waits until every peer in `mask` has raised flag[phase] to the current step, then the
// CTA proceeds.  Bounded: after ~2 s the step is marked failed (PSM_ERR_COMM) instead of hanging the GPU.
__device__ __forceinline__ void p2p_wait(const P2PArgs* P, int phase, unsigned int mask, int sample = -1) {
    const bool sampled = sample < 0 ? blockIdx.x == 0 : sample != 0;
    if (threadIdx.x == 0) {
        const unsigned int step = P->sc->step;
        const PeerMail* mine = P->mail[P->rank];
        const long long t0 = clock64();
        unsigned long long g0 = 0;
        if (sampled) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g0));
        for (int p = 0; p < P->world; ++p) {
            if (!((mask >> p) & 1u) || p == P->rank) continue;
            while ((int)(ld_acquire_sys(&mine->flag[phase][p]) - step) < 0) {
                if (clock64() - t0 > 4000000000ll) { P->sc->comm_error = 1; break; }
                __nanosleep(64);
            }
        }
        if (sampled) {               // one sample per kernel: the wait histogram
            unsigned long long g1;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g1));
            atomicAdd(&P->sc->wait_ns[phase], g1 - g0);
            atomicAdd(&P->sc->wait_n[phase], 1u);
        }
    }
    __syncthreads();
}

// A push kernel runs on several CTAs; the LAST one to finish its stores (system-scope fence, then a
// device-scope counter) raises the flags.  Returns true in thread 0..world-1 of that last CTA.
__device__ __forceinline__ bool p2p_last_block(Scalars* sc, int phase) {
    __shared__ bool s_last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(&sc->push_done[phase], 1u);
        s_last = (prev == gridDim.x - 1);
        if (s_last) sc->push_done[phase] = 0u;
    }
    __syncthreads();
    if (s_last) __threadfence();
    return s_last;
}

// Exchange 1 (after prep): my two running maxima to everyone, my cells that are ghost cells elsewhere.
__global__ void __launch_bounds__(256) p2p_push_cells_kernel(const P2PArgs* pa, const float2* uv, const int32_t* send_idx) {
    pdl_enter();
    const P2PArgs& P = *pa;
    const unsigned int step = P.sc->step + 1u;            // the last CTA publishes the increment
    const long long nsend = P.cell_send_ptr[P.world];
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nsend; e += (long long)gridDim.x * blockDim.x) {
        int p = 0;
        while (e >= P.cell_send_ptr[p + 1]) ++p;
        P.uv_ghost[p][e - P.cell_send_ptr[p]] = uv[send_idx[e]];
    }
    if (!p2p_last_block(P.sc, 0)) return;
    if (threadIdx.x == 0) P.sc->step = step;
    if ((int)threadIdx.x < P.world) {
        const int p = threadIdx.x;
        PeerMail* m = P.mail[p];
        m->maxima[P.rank][0] = P.sc->umax2_bits;
        m->maxima[P.rank][1] = P.sc->dumax2_bits;
        __threadfence_system();
        st_release_sys(&m->flag[0][P.rank], step);
    }
}
void launch_p2p_push_cells(const P2PArgs* d_pa, const float2* uv, const int32_t* send_idx, long long nsend, cudaStream_t s) {
    long long want = (nsend + 255) / 256;
    launch_k(p2p_push_cells_kernel, dim3((int)(want < 1 ? 1 : (want > 64 ? 64 : want))), dim3(256), 0, s, d_pa, uv, send_idx);
}

// Exchange 2 (after the strip means): my slots of the global mean array to everyone.
__global__ void __launch_bounds__(256) p2p_push_means_kernel(const P2PArgs* pa, const DevTask* tasks, int n_tasks, const double* means) {
    pdl_enter();
    const P2PArgs& P = *pa;
    const unsigned int step = P.sc->step;
    const int total = n_tasks * P.world;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int p = e / n_tasks, i = e - p * n_tasks;
        if (p == P.rank) continue;
        const int slot = tasks[i].out;
        P.means[p][slot] = means[slot];
    }
    if (!p2p_last_block(P.sc, 1)) return;
    if ((int)threadIdx.x < P.world && (int)threadIdx.x != P.rank) st_release_sys(&P.mail[threadIdx.x]->flag[1][P.rank], step);
}
void launch_p2p_push_means(const P2PArgs* d_pa, const DevTask* tasks, int n_tasks, int world, const double* means, cudaStream_t s) {
    long long want = ((long long)n_tasks * world + 255) / 256;
    launch_k(p2p_push_means_kernel, dim3((int)(want < 1 ? 1 : (want > 32 ? 32 : want))), dim3(256), 0, s, d_pa, tasks, n_tasks, means);
}

// Exchange 3 (after the placement): my field pixels that other ranks' grid->cell tables reference.
__global__ void __launch_bounds__(256) p2p_push_pix_kernel(const P2PArgs* pa, const float* field, const int32_t* send_idx, int F, long long my_stride) {
    pdl_enter();
    const P2PArgs& P = *pa;
    const unsigned int step = P.sc->step;
    const long long nsend = P.pix_send_ptr[P.world];
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nsend * F; e += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(e / nsend);
        const long long k = e - (long long)f * nsend;
        int p = 0;
        while (k >= P.pix_send_ptr[p + 1]) ++p;
        P.field_ghost[p][(long long)f * P.field_stride[p] + (k - P.pix_send_ptr[p])] = field[(long long)f * my_stride + send_idx[k]];
    }
    if (!p2p_last_block(P.sc, 2)) return;
    if ((int)threadIdx.x < P.world) {
        const int p = threadIdx.x;
        if (P.pix_send_ptr[p + 1] > P.pix_send_ptr[p]) st_release_sys(&P.mail[p]->flag[2][P.rank], step);
    }
}
void launch_p2p_push_pix(const P2PArgs* d_pa, const float* field, const int32_t* send_idx, int F, long long my_stride, long long nsend, cudaStream_t s) {
    long long want = (nsend * F + 255) / 256;
    launch_k(p2p_push_pix_kernel, dim3((int)(want < 1 ? 1 : (want > 32 ? 32 : want))), dim3(256), 0, s, d_pa, field, send_idx, F, my_stride);
}

// ------------------------------------------------------------------------------------------------
// K0  prep: de-interleave the solver's double[n][ncol] rows, form the field to interpolate,
//     running max of |U|^2 and |dU|^2.   PMP:267-273, SMC:386-405.
//     The squares/sum are rounded separately (no FMA) so that U_max_norm is bit-identical to
//     np.max(np.sqrt(np.square(Ux) + np.square(Uy))).
// U_max_norm, the input / output scales and the skip rule (SMC:404-419,551) from the two running maxima.
__device__ __forceinline__ void publish_scalars(const ScalarArgs& sa, unsigned long long um2, unsigned long long dm2, float& s0, float& s1) {
    Scalars* sc = sa.sc;
    const double um = sqrt(__longlong_as_double((long long)um2));
    const double dm = sqrt(__longlong_as_double((long long)dm2));
    s0 = (float)(1.0 / (um * sa.max_abs_ux)); s1 = (float)(1.0 / (um * sa.max_abs_uy));
    sc->U_max_norm = um;
    sc->dU_max_norm = dm;
    sc->in_scale[0] = s0;
    sc->in_scale[1] = s1;
    sc->out_scale = (float)(sa.dimensionalise ? sa.out_scale_base * um * um : sa.out_scale_base);
    int skip = 0;
    if (sa.mode != 0) {
        if (sa.skip_threshold > 0.0 && (dm / um) < sa.skip_threshold) skip = 1;     // SMC:410-415
        if (sa.mode == 2 && !sc->have_prev) skip = 1;                              // no U(t-1) yet
    }
    sc->skip = skip;
    sc->have_prev = 1;
}


// Multi-GPU fused flow: the LAST CTA of a prep kernel to finish pushes this rank's two running maxima and the cells that are
// ghost cells elsewhere straight into the peers' buffers, fences at system scope and raises the phase-0 flags -- no separate
// push launch.  Every CTA pays one device-scope fence + one atomic.
__device__ __forceinline__ void prep_load_runs(const P2PFused& fx, SendRun* s_runs) {
    if (fx.p2p && (int)threadIdx.x < fx.n_runs) s_runs[threadIdx.x] = fx.runs[threadIdx.x];      // static: before the wait
}
// cell i has just been converted by this thread: push it if a peer needs it as a ghost cell
__device__ __forceinline__ void prep_push_cell(const P2PFused& fx, const SendRun* s_runs, long long i, float2 val, bool& pushed) {
    if (fx.send_words) {
        const uint2 w = __ldg(fx.send_words + (i >> 5));
        const unsigned int bit = 1u << (i & 31);
        if (w.x & bit) {
            int k = (int)w.y + __popc(w.x & (bit - 1u));
            for (;;) {
                const int2 e = __ldg(fx.send_entries + k);
                fx.p2p->uv_ghost[e.x & 0xFF][e.y] = val;
                if (!(e.x & 0x100)) break;
                k = e.x >> 9;                                    // a cell that goes to several peers: its further entries are chained
            }
            pushed = true;
        }
        return;
    }
    for (int r = 0; r < fx.n_runs; ++r) {
        const SendRun& R = s_runs[r];
        if (i >= R.begin && i < R.end) { fx.p2p->uv_ghost[R.peer][R.dst + (i - R.begin)] = val; pushed = true; }
    }
}
__device__ __forceinline__ void prep_p2p_tail(const P2PFused& fx, const float2* uv, Scalars* sc, bool pushed) {
    if (!fx.p2p) return;
    __shared__ bool s_last;
    if (__syncthreads_or(pushed ? 1 : 0)) __threadfence_system();       // this CTA stored into peer memory: order it before the counter
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int prev = atomicAdd(&sc->push_done[0], 1u);
        s_last = (prev == gridDim.x - 1);
        if (s_last) { sc->push_done[0] = 0u; __threadfence(); }
    }
    __syncthreads();
    if (!s_last) return;
    const P2PArgs& P = *fx.p2p;
    const unsigned int step = P.sc->step + 1u;
    if (fx.n_runs == 0 && !fx.send_words) {
        // fragmented send list, no send map: this one CTA moves it, eight independent index -> value chains per thread in flight
        const long long nsend = P.cell_send_ptr[P.world];
        for (long long e0 = threadIdx.x; e0 < nsend; e0 += (long long)blockDim.x * 8) {
            int idx[8]; float2 val[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) { const long long e = e0 + (long long)k * blockDim.x; idx[k] = (e < nsend) ? __ldg(fx.cell_send_idx + e) : 0; }
#pragma unroll
            for (int k = 0; k < 8; ++k) val[k] = __ldcg(uv + idx[k]);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const long long e = e0 + (long long)k * blockDim.x;
                if (e < nsend) {
                    int p = 0;
                    while (e >= P.cell_send_ptr[p + 1]) ++p;
                    P.uv_ghost[p][e - P.cell_send_ptr[p]] = val[k];
                }
            }
        }
        __threadfence_system();
        __syncthreads();
    }
    if (threadIdx.x == 0) P.sc->step = step;
    if ((int)threadIdx.x < P.world) {
        const int p = threadIdx.x;
        PeerMail* m = P.mail[p];
        m->maxima[P.rank][0] = *reinterpret_cast<volatile unsigned long long*>(&sc->umax2_bits);
        m->maxima[P.rank][1] = *reinterpret_cast<volatile unsigned long long*>(&sc->dumax2_bits);
        __threadfence_system();
        st_release_sys(&m->flag[0][P.rank], step);
    }
}

template <int MODE, int NCOL>
__global__ void __launch_bounds__(256) prep_kernel(PrepArgs a) {
    pdl_enter();
    const double* __restrict__ cells = a.cells;
    double* __restrict__ p_prev = a.p_prev;
    float2* __restrict__ uv = a.uv;
    double2* __restrict__ u_prev = reinterpret_cast<double2*>(a.u_prev);
    double m_u = 0.0, m_d = 0.0;
    __shared__ SendRun s_runs[kMaxRuns];
    bool pushed = false;
    prep_load_runs(a.fx, s_runs);
    if (a.fx.p2p) __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
        const double* row = cells + i * NCOL;
        // all loads of the row first (independent), then the arithmetic
        const double ux = __ldcs(row + 0), uy = __ldcs(row + 1), pp = __ldcs(row + 4);
        double fx, fy;
        if (MODE == 0) { fx = ux; fy = uy; }
        else if (MODE == 1) { fx = __ldcs(row + 5); fy = __ldcs(row + 6); }
        else {
            const double2 prev = u_prev[i];
            fx = ux - prev.x; fy = uy - prev.y;
            u_prev[i] = make_double2(ux, uy);
        }
        p_prev[i] = pp;
        m_u = fmax(m_u, __dadd_rn(__dmul_rn(ux, ux), __dmul_rn(uy, uy)));
        if (MODE != 0) m_d = fmax(m_d, __dadd_rn(__dmul_rn(fx, fx), __dmul_rn(fy, fy)));
        const float2 val = make_float2((float)fx, (float)fy);
        uv[i] = val;
        if (a.fx.p2p) prep_push_cell(a.fx, s_runs, i, val, pushed);
    }
    m_u = warp_max(m_u);
    m_d = warp_max(m_d);
    __shared__ double s_u[8], s_d[8];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { s_u[w] = m_u; s_d[w] = m_d; }
    __syncthreads();
    if (w == 0) {
        m_u = (l < 8) ? s_u[l] : 0.0;
        m_d = (l < 8) ? s_d[l] : 0.0;
        m_u = warp_max(m_u);
        m_d = warp_max(m_d);
        if (l == 0) {   // non-negative doubles order like their bit patterns; NaN inputs are not ordered
            atomicMax(&a.sc->umax2_bits, (unsigned long long)__double_as_longlong(m_u));
            atomicMax(&a.sc->dumax2_bits, (unsigned long long)__double_as_longlong(m_d));
        }
    }
    prep_p2p_tail(a.fx, a.uv, a.sc, pushed);
}

// Bulk-copy variant (opt-in, see launch_prep): persistent CTAs stream 256-row tiles of the
// solver's row-major buffer (and of the resident U(t-1)) into shared memory with cp.async.bulk + mbarrier, four tiles
// in flight per CTA, so HBM sees long contiguous bursts instead of 8-byte column picks at a 40-byte stride; the
// threads then pick their row's columns from shared memory.  Same arithmetic and rounding as prep_kernel.
namespace {
constexpr int kPrepRows = 256, kPrepStages = 4;
__device__ __forceinline__ uint32_t prep_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void prep_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void prep_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
}  // namespace

template <int MODE, int NCOL>
__global__ void __launch_bounds__(kPrepRows) prep_bulk_kernel(PrepArgs a) {
    pdl_enter();
    constexpr int ROWS = kPrepRows, ST = kPrepStages;
    constexpr uint32_t TILE = ROWS * NCOL * 8, UP = (MODE == 2) ? ROWS * 16 : 0;
    extern __shared__ __align__(128) unsigned char prep_smem[];
    double* tiles = reinterpret_cast<double*>(prep_smem);                        // [ST][ROWS*NCOL]
    double2* ups = reinterpret_cast<double2*>(prep_smem + (size_t)ST * TILE);    // [ST][ROWS]
    __shared__ unsigned long long bars[ST];
    const long long n_tiles = a.n / ROWS;                                        // full tiles; the tail goes the plain way
    if (threadIdx.x == 0) {
        for (int s = 0; s < ST; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(prep_smem_u32(&bars[s])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](long long t, int s) {
        const uint32_t bar = prep_smem_u32(&bars[s]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(TILE + UP) : "memory");
        prep_bulk_load(prep_smem_u32(tiles + (size_t)s * ROWS * NCOL), a.cells + t * ROWS * NCOL, TILE, bar);
        if (MODE == 2) prep_bulk_load(prep_smem_u32(ups + (size_t)s * ROWS), a.u_prev + t * ROWS * 2, UP, bar);
    };
    if (threadIdx.x == 0)
        for (int k = 0; k < ST; ++k) {
            const long long t = blockIdx.x + (long long)k * gridDim.x;
            if (t < n_tiles) issue(t, k);
        }
    double m_u = 0.0, m_d = 0.0;
    double* __restrict__ p_prev = a.p_prev;
    float2* __restrict__ uv = a.uv;
    double2* __restrict__ u_prev = reinterpret_cast<double2*>(a.u_prev);
    __shared__ SendRun s_runs[kMaxRuns];
    bool pushed = false;
    prep_load_runs(a.fx, s_runs);
    if (a.fx.p2p) __syncthreads();
    auto row_work = [&](long long i, double ux, double uy, double pp, double gx, double gy, double2 prev) {
        double fx, fy;
        if (MODE == 0) { fx = ux; fy = uy; }
        else if (MODE == 1) { fx = gx; fy = gy; }
        else { fx = ux - prev.x; fy = uy - prev.y; u_prev[i] = make_double2(ux, uy); }
        p_prev[i] = pp;
        m_u = fmax(m_u, __dadd_rn(__dmul_rn(ux, ux), __dmul_rn(uy, uy)));
        if (MODE != 0) m_d = fmax(m_d, __dadd_rn(__dmul_rn(fx, fx), __dmul_rn(fy, fy)));
        const float2 val = make_float2((float)fx, (float)fy);
        uv[i] = val;
        if (a.fx.p2p) prep_push_cell(a.fx, s_runs, i, val, pushed);
    };
    int k = 0;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, ++k) {
        const int s = k % ST;
        prep_mbar_wait(prep_smem_u32(&bars[s]), (k / ST) & 1);
        const double* row = tiles + (size_t)s * ROWS * NCOL + threadIdx.x * NCOL;
        const double ux = row[0], uy = row[1], pp = row[4];
        const double gx = (MODE == 1) ? row[NCOL - 2] : 0.0, gy = (MODE == 1) ? row[NCOL - 1] : 0.0;
        const double2 prev = (MODE == 2) ? ups[(size_t)s * ROWS + threadIdx.x] : make_double2(0.0, 0.0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads of this stage before the async refill
        __syncthreads();                                         // every thread has its row: the stage can be refilled
        if (threadIdx.x == 0) {
            const long long tn = t + (long long)ST * gridDim.x;
            if (tn < n_tiles) issue(tn, s);
        }
        row_work(t * ROWS + threadIdx.x, ux, uy, pp, gx, gy, prev);
    }
    // tail rows (fewer than one tile): plain loads, spread over the CTAs
    for (long long i = n_tiles * ROWS + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x) {
        const double* row = a.cells + i * NCOL;
        const double2 prev = (MODE == 2) ? u_prev[i] : make_double2(0.0, 0.0);
        row_work(i, row[0], row[1], row[4], MODE == 1 ? row[NCOL - 2] : 0.0, MODE == 1 ? row[NCOL - 1] : 0.0, prev);
    }
    m_u = warp_max(m_u);
    m_d = warp_max(m_d);
    __shared__ double s_u[8], s_d[8];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { s_u[w] = m_u; s_d[w] = m_d; }
    __syncthreads();
    if (w == 0) {
        m_u = (l < 8) ? s_u[l] : 0.0;
        m_d = (l < 8) ? s_d[l] : 0.0;
        m_u = warp_max(m_u);
        m_d = warp_max(m_d);
        if (l == 0) {
            atomicMax(&a.sc->umax2_bits, (unsigned long long)__double_as_longlong(m_u));
            atomicMax(&a.sc->dumax2_bits, (unsigned long long)__double_as_longlong(m_d));
        }
    }
    prep_p2p_tail(a.fx, a.uv, a.sc, pushed);
}
template <int MODE, int NCOL>
static void launch_prep_bulk(const PrepArgs& a, cudaStream_t s) {
    const size_t smem = (size_t)kPrepStages * (kPrepRows * NCOL * 8 + (MODE == 2 ? kPrepRows * 16 : 0));
    static bool opted[64] = {};                                                    // the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!opted[dev & 63]) { cudaFuncSetAttribute(prep_bulk_kernel<MODE, NCOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); opted[dev & 63] = true; }
    const long long tiles = a.n / kPrepRows;
    const int per_sm = (int)(200 * 1024 / (smem + 1024));                         // resident CTAs per SM by shared memory
    long long blocks = (long long)kSMs * (per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm));
    if (blocks > tiles) blocks = tiles > 0 ? tiles : 1;
    launch_k(prep_bulk_kernel<MODE, NCOL>, dim3((int)blocks), dim3(kPrepRows), smem, s, a);
}

void launch_prep(const PrepArgs& a, cudaStream_t s) {
    // Default: the bulk-copy kernel (contiguous cp.async.bulk bursts instead of 8-byte column picks at a 40-byte stride; 6.6 TB/s
    // at 4 M cells).  Its round-1 race -- a stage refilled through the async proxy while generic reads of it were still in
    // flight -- is closed by the fence.proxy.async before the CTA barrier (tests/test_gpu_fullsize.py replays steps bit for
    // bit).  PSM_PREP_BULK=0 selects the plain kernel.
    static const bool bulk_ok = [] { const char* v = getenv("PSM_PREP_BULK"); return !(v && v[0] == '0'); }();
    if (bulk_ok && (reinterpret_cast<uintptr_t>(a.cells) & 15) == 0 && a.n >= kPrepRows) {
        if (a.mode == 0) launch_prep_bulk<0, 5>(a, s);
        else if (a.mode == 1) launch_prep_bulk<1, 7>(a, s);
        else launch_prep_bulk<2, 5>(a, s);
        return;
    }
    long long want = (a.n + 255) / 256;
    int blocks = (int)(want < (long long)kSMs * 8 ? (want > 0 ? want : 1) : kSMs * 8);
    if (a.mode == 0) launch_k(prep_kernel<0, 5>, dim3(blocks), dim3(256), 0, s, a);
    else if (a.mode == 1) launch_k(prep_kernel<1, 7>, dim3(blocks), dim3(256), 0, s, a);
    else launch_k(prep_kernel<2, 5>, dim3(blocks), dim3(256), 0, s, a);
}

// K0 on the solver's own field arrays (see psm_kernels.cuh).  Same arithmetic and rounding as prep_kernel.
template <int MODE>
__global__ void __launch_bounds__(256) prep_fields_kernel(PrepFieldsArgs a) {
    pdl_enter();
    const double* __restrict__ U = a.U;
    const double* __restrict__ dU = a.dU;
    float2* __restrict__ uv = a.uv;
    double2* __restrict__ u_prev = reinterpret_cast<double2*>(a.u_prev);
    const int st = a.stride;
    double m_u = 0.0, m_d = 0.0;
    __shared__ SendRun s_runs[kMaxRuns];
    bool pushed = false;
    prep_load_runs(a.fx, s_runs);
    if (a.fx.p2p) __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
        const double ux = __ldcs(U + i * st), uy = __ldcs(U + i * st + 1);
        double fx, fy;
        if (MODE == 0) { fx = ux; fy = uy; }
        else if (MODE == 1) { fx = __ldcs(dU + i * st); fy = __ldcs(dU + i * st + 1); }
        else {
            const double2 prev = u_prev[i];
            fx = ux - prev.x; fy = uy - prev.y;
            u_prev[i] = make_double2(ux, uy);
        }
        m_u = fmax(m_u, __dadd_rn(__dmul_rn(ux, ux), __dmul_rn(uy, uy)));
        if (MODE != 0) m_d = fmax(m_d, __dadd_rn(__dmul_rn(fx, fx), __dmul_rn(fy, fy)));
        const float2 val = make_float2((float)fx, (float)fy);
        uv[i] = val;
        if (a.fx.p2p) prep_push_cell(a.fx, s_runs, i, val, pushed);
    }
    m_u = warp_max(m_u);
    m_d = warp_max(m_d);
    __shared__ double s_u[8], s_d[8];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { s_u[w] = m_u; s_d[w] = m_d; }
    __syncthreads();
    if (w == 0) {
        m_u = (l < 8) ? s_u[l] : 0.0;
        m_d = (l < 8) ? s_d[l] : 0.0;
        m_u = warp_max(m_u);
        m_d = warp_max(m_d);
        if (l == 0) {
            atomicMax(&a.sc->umax2_bits, (unsigned long long)__double_as_longlong(m_u));
            atomicMax(&a.sc->dumax2_bits, (unsigned long long)__double_as_longlong(m_d));
        }
    }
    prep_p2p_tail(a.fx, a.uv, a.sc, pushed);
}
void launch_prep_fields(const PrepFieldsArgs& a, cudaStream_t s) {
    long long want = (a.n + 255) / 256;
    int blocks = (int)(want < (long long)kSMs * 8 ? (want > 0 ? want : 1) : kSMs * 8);
    if (a.mode == 0) launch_k(prep_fields_kernel<0>, dim3(blocks), dim3(256), 0, s, a);
    else if (a.mode == 1) launch_k(prep_fields_kernel<1>, dim3(blocks), dim3(256), 0, s, a);
    else launch_k(prep_fields_kernel<2>, dim3(blocks), dim3(256), 0, s, a);
}

// ------------------------------------------------------------------------------------------------
// K1  cell -> grid gather.  UTL:88-89 (einsum over np.take), SMC:432-444 (scatter into the grid,
//     NaN -> 0, max-abs scaling).  Validity, the negative-weight NaN rule and the (0,0) raster
//     quirk are folded into the tables at init, so this is a pure weighted gather.
__device__ __forceinline__ float2 ldg_f2(const float2* p) { return __ldg(p); }

// The step's scalars come straight from the running maxima of prep (all-reduced over ranks in the multi-GPU
// path): every thread derives the two input scales itself; thread 0 publishes U_max_norm, the output scale
// and the skip rule (SMC:404-419,551) for the later kernels.  The maxima are re-armed by offsets_kernel.
// Scalars of the step, derived by every thread from the running maxima of prep; thread 0 of CTA 0 publishes them.
template <bool READY>
__device__ __forceinline__ void step_scales(const GatherArgs& a, float& s0, float& s1) {
    Scalars* sc = a.sa.sc;
    if (READY) { s0 = sc->in_scale[0]; s1 = sc->in_scale[1]; return; }
    unsigned long long um2 = sc->umax2_bits, dm2 = sc->dumax2_bits;
    if (a.p2p) {                                   // maxima over all ranks, pushed into my mailbox
        p2p_wait(a.p2p, 0, 0xFFu);
        const PeerMail* mine = a.p2p->mail[a.p2p->rank];
        for (int p = 0; p < a.p2p->world; ++p) {
            um2 = max(um2, *reinterpret_cast<const volatile unsigned long long*>(&mine->maxima[p][0]));
            dm2 = max(dm2, *reinterpret_cast<const volatile unsigned long long*>(&mine->maxima[p][1]));
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) publish_scalars(a.sa, um2, dm2, s0, s1);
    const double um = sqrt(__longlong_as_double((long long)um2));
    s0 = (float)(1.0 / (um * a.sa.max_abs_ux)); s1 = (float)(1.0 / (um * a.sa.max_abs_uy));
}

struct PixTables { int4 i0, i1, i2; float4 q0, q1, q2; };
__device__ __forceinline__ PixTables load_tables(const GatherArgs& a, long long g) {
    PixTables t;
    t.i0 = __ldcs(reinterpret_cast<const int4*>(a.v0) + g);
    t.i1 = __ldcs(reinterpret_cast<const int4*>(a.v1) + g);
    t.i2 = __ldcs(reinterpret_cast<const int4*>(a.v2) + g);
    t.q0 = __ldcs(reinterpret_cast<const float4*>(a.w0) + g);
    t.q1 = __ldcs(reinterpret_cast<const float4*>(a.w1) + g);
    t.q2 = __ldcs(reinterpret_cast<const float4*>(a.w2) + g);
    return t;
}
// four pixels: weighted gather (UTL:88-89), scaling (SMC:441-442), NaN -> 0 (SMC:438).  Split in two so that a kernel can
// issue the twelve gathers BEFORE it touches the step's scalars (in-order issue: a stalled scalar load would hold them back).
struct PixVals { float2 a0, b0, c0, a1, b1, c1, a2, b2, c2, a3, b3, c3; };
__device__ __forceinline__ void gather4_load(const float2* __restrict__ uv, const PixTables& t, PixVals& v) {
    const int4 i0 = t.i0, i1 = t.i1, i2 = t.i2;
    v.a0 = ldg_f2(uv + i0.x); v.b0 = ldg_f2(uv + i1.x); v.c0 = ldg_f2(uv + i2.x);
    v.a1 = ldg_f2(uv + i0.y); v.b1 = ldg_f2(uv + i1.y); v.c1 = ldg_f2(uv + i2.y);
    v.a2 = ldg_f2(uv + i0.z); v.b2 = ldg_f2(uv + i1.z); v.c2 = ldg_f2(uv + i2.z);
    v.a3 = ldg_f2(uv + i0.w); v.b3 = ldg_f2(uv + i1.w); v.c3 = ldg_f2(uv + i2.w);
}
__device__ __forceinline__ void gather4_combine(const PixVals& v, const PixTables& t, float s0, float s1, float4& ox, float4& oy) {
    const float4 q0 = t.q0, q1 = t.q1, q2 = t.q2;
    ox.x = (v.a0.x * q0.x + v.b0.x * q1.x + v.c0.x * q2.x) * s0;  oy.x = (v.a0.y * q0.x + v.b0.y * q1.x + v.c0.y * q2.x) * s1;
    ox.y = (v.a1.x * q0.y + v.b1.x * q1.y + v.c1.x * q2.y) * s0;  oy.y = (v.a1.y * q0.y + v.b1.y * q1.y + v.c1.y * q2.y) * s1;
    ox.z = (v.a2.x * q0.z + v.b2.x * q1.z + v.c2.x * q2.z) * s0;  oy.z = (v.a2.y * q0.z + v.b2.y * q1.z + v.c2.y * q2.z) * s1;
    ox.w = (v.a3.x * q0.w + v.b3.x * q1.w + v.c3.x * q2.w) * s0;  oy.w = (v.a3.y * q0.w + v.b3.y * q1.w + v.c3.y * q2.w) * s1;
    ox.x = (ox.x != ox.x) ? 0.f : ox.x; ox.y = (ox.y != ox.y) ? 0.f : ox.y;
    ox.z = (ox.z != ox.z) ? 0.f : ox.z; ox.w = (ox.w != ox.w) ? 0.f : ox.w;
    oy.x = (oy.x != oy.x) ? 0.f : oy.x; oy.y = (oy.y != oy.y) ? 0.f : oy.y;
    oy.z = (oy.z != oy.z) ? 0.f : oy.z; oy.w = (oy.w != oy.w) ? 0.f : oy.w;
}
__device__ __forceinline__ void gather4(const float2* __restrict__ uv, const PixTables& t, float s0, float s1, float4& ox, float4& oy) {
    PixVals v;
    gather4_load(uv, t, v);
    gather4_combine(v, t, s0, s1, ox, oy);
}

// The tables are static: every thread loads those of its first pixel group BEFORE waiting for the previous
// kernel (programmatic dependent launch), so the first DRAM round trip overlaps the previous kernel's tail.
template <bool READY>
__global__ void __launch_bounds__(256) gather_kernel(GatherArgs a) {
    pdl_launch_dependents();
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    PixTables t{};
    if (g < a.n_pix4) t = load_tables(a, g);
    pdl_wait();
    float s0, s1;
    step_scales<READY>(a, s0, s1);
    while (g < a.n_pix4) {
        float4 ox, oy;
        gather4(a.uv, t, s0, s1, ox, oy);
        reinterpret_cast<float4*>(a.grid0)[g] = ox;
        reinterpret_cast<float4*>(a.grid1)[g] = oy;
        g += stride;
        if (g < a.n_pix4) t = load_tables(a, g);
    }
}
void launch_gather(const GatherArgs& a, cudaStream_t s) {
    long long want = (a.n_pix4 + 255) / 256;
    int blocks = (int)(want < (long long)kSMs * 8 ? (want > 0 ? want : 1) : kSMs * 8);
    if (a.scales_ready || a.sa.replay) launch_k(gather_kernel<true>, dim3(blocks), dim3(256), 0, s, a);
    else launch_k(gather_kernel<false>, dim3(blocks), dim3(256), 0, s, a);
}

// K1+K2 fused (see psm_kernels.cuh): identical arithmetic to gather_kernel, then one 128-bit store per
// covering block and channel.  U pixel groups per thread are in flight at once: all their (static) table loads are issued
// before the programmatic-dependent-launch wait, so a thread pays the HBM latency of the table stream once per U groups;
// the store addresses come from two small static offset tables (no block-origin lookups, no local memory).
__device__ __forceinline__ void store_cover(float* __restrict__ xu, const CoverEntry* __restrict__ rowcov, const CoverEntry* __restrict__ colcov,
                                            unsigned int g, unsigned int W4, int S2, const float4& ox, const float4& oy) {
    const unsigned int y = g / W4, xg = g - y * W4;
    const int4* rp = reinterpret_cast<const int4*>(rowcov + y);
    const int4* cp = reinterpret_cast<const int4*>(colcov + xg);
    const int4 r0 = __ldg(rp), c0 = __ldg(cp);
    const int rn = r0.x, cn = c0.x;
    int4 r1 = make_int4(0, 0, 0, 0), c1 = make_int4(0, 0, 0, 0);
    if (rn > 3) r1 = __ldg(rp + 1);
    if (cn > 3) c1 = __ldg(cp + 1);
    // rolled loops with register selects: no local-memory arrays, few live registers
    auto pick = [](const int4& lo, const int4& hi, int k) {
        int v = lo.y;
        v = (k == 1) ? lo.z : v; v = (k == 2) ? lo.w : v; v = (k == 3) ? hi.x : v;
        v = (k == 4) ? hi.y : v; v = (k == 5) ? hi.z : v; v = (k == 6) ? hi.w : v;
        return v;
    };
    for (int r = 0; r < rn; ++r) {
        float* row = xu + pick(r0, r1, r);
        for (int c = 0; c < cn; ++c) {
            float* dst = row + pick(c0, c1, c);
            *reinterpret_cast<float4*>(dst) = ox;
            *reinterpret_cast<float4*>(dst + S2) = oy;
        }
    }
}

template <bool READY, int U>
__global__ void __launch_bounds__(256, U == 1 ? 4 : 3) gather_extract_kernel(GatherExtractArgs e) {
    pdl_launch_dependents();
    const GatherArgs& a = e.g;
    const long long T = (long long)gridDim.x * blockDim.x;
    const int S2 = e.S * e.S;
    long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    PixTables t[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (g + u * T < a.n_pix4) t[u] = load_tables(a, g + u * T);
    pdl_wait();
    // The step's scalars: their LOADS are issued here, their first USE sits behind the first group's gathers (in-order issue:
    // a warp parked on the scalar round trip would otherwise have no gather in flight).  With the peer-memory wait in front
    // (multi-GPU) the ghost cells must land first, so the old order is kept there.
    Scalars* sc = a.sa.sc;
    float s0 = 0.f, s1 = 0.f;
    unsigned long long um2 = 0ull, dm2 = 0ull;
    bool have = false;
    if (READY || a.p2p) { step_scales<READY>(a, s0, s1); have = true; }
    else { um2 = sc->umax2_bits; dm2 = sc->dumax2_bits; }
    while (g < a.n_pix4) {
        float4 ox[U], oy[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (g + u * T < a.n_pix4) {
                PixVals pv;
                gather4_load(a.uv, t[u], pv);
                if (!have) {
                    if (blockIdx.x == 0 && threadIdx.x == 0) publish_scalars(a.sa, um2, dm2, s0, s1);
                    const double um = sqrt(__longlong_as_double((long long)um2));
                    s0 = (float)(1.0 / (um * a.sa.max_abs_ux)); s1 = (float)(1.0 / (um * a.sa.max_abs_uy));
                    have = true;
                }
                gather4_combine(pv, t[u], s0, s1, ox[u], oy[u]);
            }
        const long long gn = g + (long long)U * T;
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (gn + u * T < a.n_pix4) t[u] = load_tables(a, gn + u * T);            // next pass: tables in flight during the stores
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long gu = g + u * T;
            if (gu < a.n_pix4) {
                if (e.store_grid) {
                    reinterpret_cast<float4*>(a.grid0)[gu] = ox[u];
                    reinterpret_cast<float4*>(a.grid1)[gu] = oy[u];
                }
                if (!e.skip_xu) store_cover(e.xu, e.rowcov, e.colcov, (unsigned int)gu, (unsigned int)e.W4, S2, ox[u], oy[u]);
            }
        }
        g = gn;
    }
}
void launch_gather_extract(const GatherExtractArgs& a, cudaStream_t s) {
    // One pixel group per thread and four CTAs per SM is the default.  Two groups in flight per thread (PSM_GATHER_U=2: twelve table
    // loads before the wait, 3 CTAs / SM) measured equal at 1 M cells (99.1 vs 99.8 us per step) and slower at 4 M (gather 43.8 vs
    // 38.7 us): at 80 registers the second group's tables spill.
    static const int force_u = [] { const char* v = getenv("PSM_GATHER_U"); return v ? atoi(v) : 0; }();
    const long long wave2 = (long long)kSMs * 3 * 256, wave1 = (long long)kSMs * 4 * 256;
    const int U = (force_u == 2) ? 2 : 1;
    long long want = (a.g.n_pix4 + 255) / 256;
    if (U == 2) {
        want = (a.g.n_pix4 + 511) / 512;
        int blocks = (int)(want < wave2 / 256 ? (want > 0 ? want : 1) : wave2 / 256);
        if (a.g.scales_ready) launch_k(gather_extract_kernel<true, 2>, dim3(blocks), dim3(256), 0, s, a);
        else launch_k(gather_extract_kernel<false, 2>, dim3(blocks), dim3(256), 0, s, a);
        return;
    }
    int blocks = (int)(want < wave1 / 256 ? (want > 0 ? want : 1) : wave1 / 256);     // one resident wave (4 CTAs / SM), grid-stride
    if (a.g.scales_ready) launch_k(gather_extract_kernel<true, 1>, dim3(blocks), dim3(256), 0, s, a);
    else launch_k(gather_extract_kernel<false, 1>, dim3(blocks), dim3(256), 0, s, a);
}

// ------------------------------------------------------------------------------------------------
// K2  block extraction.  SMC:464-492: slices grid[y0:y0+S, x0:x0+S, 0:2] per plan entry; the
//     operand is stored planar (c, ly, lx) -- the PCA matrix is permuted to match at init.
__global__ void __launch_bounds__(256) extract_kernel(ExtractArgs a) {
    pdl_enter();
    // one warp per (block, channel, row): 128 floats = 32 lanes x 4
    const int S = a.S;
    const long long rows = (long long)a.B * a.nch * S;
    const int lane = threadIdx.x & 31;
    const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += wstride) {
        const int ly = (int)(r % S);
        const int c = (int)((r / S) % a.nch);
        const int b = (int)(r / ((long long)a.nch * S));
        const float* src = (c ? a.grid1 : a.grid0) + (long long)(a.by0[b] + ly) * a.W + a.bx0[b] + lane * 4;
        float4 v = make_float4(src[0], src[1], src[2], src[3]);
        reinterpret_cast<float4*>(a.xu + ((long long)b * a.nch + c) * S * S + (long long)ly * S)[lane] = v;
    }
}
void launch_extract(const ExtractArgs& a, cudaStream_t s) {
    long long rows = (long long)a.B * a.nch * a.S;
    long long want = (rows + 7) / 8;
    int blocks = (int)(want < (long long)kSMs * 8 ? want : kSMs * 8);
    launch_k(extract_kernel, dim3(blocks), dim3(256), 0, s, a);
}

// ------------------------------------------------------------------------------------------------
// FP32 GEMM (CUDA cores), C = A * B^T, 64x64x16 tiles, 4x4 register blocking.
template <int EPI>
__global__ void __launch_bounds__(256) sgemm_nt_kernel(GemmArgs a) {
    pdl_enter();
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    int k_begin = 0, k_end = a.K;
    if (EPI == EPI_PARTIAL) {
        const int kper = ((a.K / BK + a.splits - 1) / a.splits) * BK;
        k_begin = blockIdx.z * kper;
        k_end = min(a.K, k_begin + kper);
    }
    const int lr = tid >> 2, lc = (tid & 3) * 4;      // loader: row lr (0..63), k offset lc
    const float* Ap = a.A + (long long)(m0 + lr) * a.lda + lc;
    const float* Bp = a.B + (long long)(n0 + lr) * a.ldb + lc;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
        const float4 av = *reinterpret_cast<const float4*>(Ap + k0);
        const float4 bv = *reinterpret_cast<const float4*>(Bp + k0);
        As[lc + 0][lr] = av.x; As[lc + 1][lr] = av.y; As[lc + 2][lr] = av.z; As[lc + 3][lr] = av.w;
        Bs[lc + 0][lr] = bv.x; Bs[lc + 1][lr] = bv.y; Bs[lc + 2][lr] = bv.z; Bs[lc + 3][lr] = bv.w;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 ra = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 rb = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float am[4] = {ra.x, ra.y, ra.z, ra.w};
            const float bn[4] = {rb.x, rb.y, rb.z, rb.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(am[i], bn[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* Cp = a.C;
    if (EPI == EPI_PARTIAL) Cp += (long long)blockIdx.z * a.M * a.ldc;
    const int n = n0 + tx * 4;
    float o_scale = 1.f;
    if (EPI == EPI_PCA_INV) o_scale = a.sc->out_scale;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        float4 o = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        if (EPI == EPI_BIAS_RELU) {
            const float4 b = *reinterpret_cast<const float4*>(a.v0 + n);
            o.x = fmaxf(o.x + b.x, 0.f); o.y = fmaxf(o.y + b.y, 0.f); o.z = fmaxf(o.z + b.z, 0.f); o.w = fmaxf(o.w + b.w, 0.f);
        } else if (EPI == EPI_BIAS_AFFINE) {
            const float4 b = *reinterpret_cast<const float4*>(a.v0 + n);
            const float4 sc = *reinterpret_cast<const float4*>(a.v1 + n);
            const float4 sh = *reinterpret_cast<const float4*>(a.v2 + n);
            o.x = (o.x + b.x) * sc.x + sh.x; o.y = (o.y + b.y) * sc.y + sh.y;
            o.z = (o.z + b.z) * sc.z + sh.z; o.w = (o.w + b.w) * sc.w + sh.w;
        } else if (EPI == EPI_PCA_INV) {
            const float4 b = *reinterpret_cast<const float4*>(a.v0 + n);
            o.x = (o.x + b.x) * o_scale; o.y = (o.y + b.y) * o_scale; o.z = (o.z + b.z) * o_scale; o.w = (o.w + b.w) * o_scale;
        }
        *reinterpret_cast<float4*>(Cp + (long long)m * a.ldc + n) = o;
    }
}

void launch_sgemm(const GemmArgs& a, cudaStream_t s) {
    dim3 grid(a.N / 64, a.M / 64, a.epi == EPI_PARTIAL ? a.splits : 1);
    switch (a.epi) {
        case EPI_PARTIAL:     launch_k(sgemm_nt_kernel<EPI_PARTIAL>, grid, dim3(256), 0, s, a); break;
        case EPI_BIAS_RELU:   launch_k(sgemm_nt_kernel<EPI_BIAS_RELU>, grid, dim3(256), 0, s, a); break;
        case EPI_BIAS_AFFINE: launch_k(sgemm_nt_kernel<EPI_BIAS_AFFINE>, grid, dim3(256), 0, s, a); break;
        case EPI_PCA_INV:     launch_k(sgemm_nt_kernel<EPI_PCA_INV>, grid, dim3(256), 0, s, a); break;
        default:              launch_k(sgemm_nt_kernel<EPI_PLAIN>, grid, dim3(256), 0, s, a); break;
    }
}

// 256 threads = 32 consecutive outputs x 8 split groups: every thread has splits/8 independent
// coalesced loads in flight, then the 8 groups fold through shared memory in a fixed order.
__global__ void __launch_bounds__(256) reduce_standardise_kernel(ReduceArgs a) {
    pdl_enter();
    const long long total = (long long)a.M * a.N;
    const int ox = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const long long i = (long long)blockIdx.x * 32 + ox;
    float sum = 0.f;
    if (i < total) {
        // the partials may be in another row order (A operand fetched from the grid planes in box order)
        long long ptotal = total, pi = i;
        bool live = true;
        if (a.row_src) {
            const int m = (int)(i / a.N), src = a.row_src[m];
            ptotal = (long long)a.part_rows * a.N;
            live = src >= 0;
            pi = (long long)(live ? src : 0) * a.N + (i - (long long)m * a.N);
        }
        // up to 16 partials per thread, all loads in flight at once (the partials sit in L2: the kernel is latency-bound)
        float pv[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) { const int sp = grp + 8 * k; pv[k] = (live && sp < a.splits) ? __ldcs(a.part + (long long)sp * ptotal + pi) : 0.f; }
#pragma unroll
        for (int k = 0; k < 16; ++k) sum += pv[k];
        if (live) for (int sp = grp + 128; sp < a.splits; sp += 8) sum += __ldcs(a.part + (long long)sp * ptotal + pi);
    }
    __shared__ float red[8][33];
    red[grp][ox] = sum;
    __syncthreads();
    if (grp == 0 && i < total) {
        float tot = red[0][ox];
#pragma unroll
        for (int g2 = 1; g2 < 8; ++g2) tot += red[g2][ox];
        const int n = (int)(i % a.N);
        float o;
        if (a.kind == RED_STANDARDISE) o = (tot + a.zc[i]) * a.a[n] + a.b[n];
        else if (a.kind == RED_BIAS_RELU) o = fmaxf(tot + a.bias[n], 0.f);
        else o = (tot + a.bias[n]) * a.a[n] + a.b[n];
        a.x[i] = o;
        if (a.x_lo) {
            const float hi = __uint_as_float(__float_as_uint(o) & 0xFFFFE000u);
            if (a.x_hi) a.x_hi[i] = hi;
            a.x_lo[i] = o - hi;
        }
    }
}
void launch_reduce_standardise(const ReduceArgs& a, cudaStream_t s) {
    long long total = (long long)a.M * a.N;
    launch_k(reduce_standardise_kernel, dim3((unsigned)((total + 31) / 32)), dim3(256), 0, s, a);
}

// ------------------------------------------------------------------------------------------------
// K6a  masked rectangle means.  Every np.mean(pred[rect][flow_bool[rect] != 0]) of SMC:233-316 /
//      GRAD:300-340 on the UNCORRECTED blocks, plus the plain line sums the global shift needs
//      (SMC:350 / GRAD:358-361); FP64 accumulation.
//      Two deterministic passes (no atomics): one warp per (task, row) writes a row partial, then one
//      warp per task folds its rows -- rectangles range from 1 x 1 to 120 x 128 pixels, so the work
//      is balanced per row, not per task.
// One CTA (8 warps) per task: warp w takes rows y0+w, y0+w+8, ... of the rectangle (each lane 4 pixels of the
// 128-pixel block row, FP64), then the 8 warp partials are folded by a fixed butterfly -- deterministic, no
// atomics, one launch.
// The task list and the flow mask are static: they are fetched before waiting for the predicted blocks.
__device__ __forceinline__ double ld_cg_f64(const double* p) { return __ldcg(p); }

__device__ void offsets_body(const OffsetsArgs& a, const DevRec* rec, const DevShiftTerm* terms, unsigned char* scratch);

// MODE 0: means only; 1: single GPU, the last CTA also runs the offsets.  (Pushing the means to the peers from here and
// waiting for theirs in the last CTA was measured slower at 2 GPUs -- one system-scope fence per CTA -- than the separate
// push kernel + offsets kernel, so the multi-GPU path keeps those.)
template <int MODE>
__global__ void __launch_bounds__(256, 3) task_means_kernel(MeansArgs a, OffsetsArgs oa) {
    pdl_launch_dependents();
    constexpr int NW = 8, NR = 16;                               // warps per CTA; rectangles have up to 128 rows -> <= 16 rows per warp
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const DevTask t = a.tasks[blockIdx.x];
    const int x = lane * 4;
    const bool c0 = x + 0 >= t.x0 && x + 0 < t.x1, c1 = x + 1 >= t.x0 && x + 1 < t.x1;
    const bool c2 = x + 2 >= t.x0 && x + 2 < t.x1, c3 = x + 3 >= t.x0 && x + 3 < t.x1;
    // bit 4k+j: pixel j of this lane's row k counts (inside the rectangle columns, and inside the flow mask for a
    // masked mean) -- static, evaluated before the wait
    const unsigned int cols = (c0 ? 1u : 0u) | (c1 ? 2u : 0u) | (c2 ? 4u : 0u) | (c3 ? 8u : 0u);
    unsigned long long bits = 0ull;
    const uint8_t* msk = a.gmask + (long long)t.my0 * a.W + t.mx0 + x;
#pragma unroll
    for (int k = 0; k < NR; ++k) {
        const int y = t.y0 + w + k * NW;
        if (y < t.y1) {
            unsigned int mrow = 0xFu;
            if (t.kind == 0) {
                const uint8_t* m = msk + (long long)y * a.W;      // byte loads: W and the block origin need not be multiples of 4
                mrow = (m[0] ? 1u : 0u) | (m[1] ? 2u : 0u) | (m[2] ? 4u : 0u) | (m[3] ? 8u : 0u);
            }
            bits |= (unsigned long long)(mrow & cols) << (4 * k);
        }
    }
    // fused offsets: the static recurrence records and shift terms go to shared memory NOW (before the wait), so the
    // serial tail of the last CTA does not start with two dependent round trips to L2 / HBM
    extern __shared__ __align__(16) unsigned char mk_smem[];
    DevRec* s_rec = nullptr; DevShiftTerm* s_terms = nullptr;
    if (MODE != 0) {
        const int n = oa.B * oa.F, nt = oa.term_start[oa.F];
        s_rec = reinterpret_cast<DevRec*>(mk_smem + (((size_t)n * 24 + 15) & ~(size_t)15));
        s_terms = reinterpret_cast<DevShiftTerm*>(s_rec + n);
        for (int i = threadIdx.x; i < n; i += blockDim.x) s_rec[i] = oa.rec[i];
        for (int i = threadIdx.x; i < nt; i += blockDim.x) s_terms[i] = oa.terms[i];
    }
    pdl_wait();
    const float* src = a.blocks + ((long long)t.src * a.C + t.ch) * a.S * a.S + x;
    double sum = 0.0;
    for (int k0 = 0; k0 < NR; k0 += 4) {                         // batches of 4 rows in flight; most rectangles need one batch
        if (t.y0 + w + k0 * NW >= t.y1) break;
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int y = min(t.y0 + w + (k0 + k) * NW, t.y1 - 1);   // rows past the rectangle: re-read its last row, not counted
            v[k] = __ldcg(reinterpret_cast<const float4*>(src + (long long)y * a.S));
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned int b = (unsigned int)(bits >> (4 * (k0 + k)));
            if (b & 1u) sum += (double)v[k].x;
            if (b & 2u) sum += (double)v[k].y;
            if (b & 4u) sum += (double)v[k].z;
            if (b & 8u) sum += (double)v[k].w;
        }
    }
    sum = warp_sum(sum);
    __shared__ double part[NW];
    if (lane == 0) part[w] = sum;
    __syncthreads();
    if (w == 0) {
        const double tot = warp_sum(lane < NW ? part[lane] : 0.0);   // fixed butterfly: deterministic
        const double val = (t.kind == 1) ? tot : ((t.count > 0) ? tot / (double)t.count : CUDART_NAN);   // 1: line sum; 0 / 2: masked / plain mean
        if (lane == 0) a.means[t.out] = val;
    }
    if (MODE != 0) {
        // the LAST CTA to publish its mean runs the offset recurrence (a few hundred scalars) right here
        __shared__ bool s_last;
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned int prev = atomicAdd(&oa.sc->means_done, 1u);
            s_last = (prev == gridDim.x - 1);
            if (s_last) { oa.sc->means_done = 0u; __threadfence(); }
        }
        __syncthreads();
        if (!s_last) return;
        offsets_body(oa, s_rec, s_terms, mk_smem);
    }
}
void launch_means(const MeansArgs& a, const OffsetsArgs* fused, cudaStream_t s) {
    if (a.n_tasks <= 0) return;
    if (fused) {
        const size_t n = (size_t)fused->B * fused->F;
        const size_t smem = ((n * 24 + 15) & ~(size_t)15) + n * sizeof(DevRec) + (size_t)fused->term_start[fused->F] * sizeof(DevShiftTerm);
        launch_k(task_means_kernel<1>, dim3(a.n_tasks), dim3(256), smem, s, a, *fused);
    }
    else launch_k(task_means_kernel<0>, dim3(a.n_tasks), dim3(256), 0, s, a, OffsetsArgs{});
}

// K6a'  the masked means from the row partials of the PCA-inverse epilogue (StripRows in psm_kernels.cuh): one warp per
//       task sums its 4 * rows FP32 partials in FP64 -- lane-strided, then a fixed butterfly: deterministic -- so the 8 MB of
//       predicted blocks are not read a second time.  Same MODE switch as task_means_kernel.
template <int MODE>
__global__ void __launch_bounds__(256) task_fold_kernel(MeansArgs a, const float* __restrict__ rowpart, OffsetsArgs oa, P2PFused fx, int n_task_ctas) {
    pdl_launch_dependents();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const bool push_cta = (MODE == 2) && (int)blockIdx.x >= n_task_ctas;          // MODE 2: CTAs behind the task CTAs push ghost pixels
    const int ti = blockIdx.x * 8 + w;
    DevTask t{};
    if (!push_cta && ti < a.n_tasks) t = a.tasks[ti];
    extern __shared__ __align__(16) unsigned char mk_smem[];
    DevRec* s_rec = nullptr; DevShiftTerm* s_terms = nullptr;
    if (MODE != 0) {
        const int n = oa.B * oa.F, nt = oa.term_start[oa.F];
        s_rec = reinterpret_cast<DevRec*>(mk_smem + (((size_t)n * 24 + 15) & ~(size_t)15));
        s_terms = reinterpret_cast<DevShiftTerm*>(s_rec + n);
        for (int i = threadIdx.x; i < n; i += blockDim.x) s_rec[i] = oa.rec[i];
        for (int i = threadIdx.x; i < nt; i += blockDim.x) s_terms[i] = oa.terms[i];
    }
    pdl_wait();
    if (push_cta) {
        // my predicted pixels that other ranks' grid->cell tables reference: straight from the blocks into their ghost regions
        const P2PArgs& P = *fx.p2p;
        const long long total = fx.n_pix_send * fx.C;
        const long long stride = (long long)(gridDim.x - n_task_ctas) * blockDim.x;
        for (long long e = (long long)(blockIdx.x - n_task_ctas) * blockDim.x + threadIdx.x; e < total; e += stride) {
            const int f = (int)(e / fx.n_pix_send);
            const long long k = e - (long long)f * fx.n_pix_send;
            int p = 0;
            while (k >= P.pix_send_ptr[p + 1]) ++p;
            const long long slot = P.pix_slot0[p] + (k - P.pix_send_ptr[p]);
            P.blk_ghost[p][((slot / fx.S2) * fx.C + f) * fx.S2 + slot % fx.S2] = __ldcg(fx.blocks + fx.pix_send_blk[k] + (long long)f * fx.S2);
        }
        __threadfence_system();
    } else if (ti < a.n_tasks) {
        // <= 4 * 128 partials per task: 16 per lane, all loads in flight before the (fixed-order) FP64 adds
        const float* p = rowpart + (long long)t.part_base * 4;
        const int n = 4 * (t.y1 - t.y0);
        float pv[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) { const int i = lane + 32 * k; pv[k] = (i < n) ? __ldcg(p + i) : 0.f; }
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < 16; ++k) acc += (double)pv[k];
        for (int i = lane + 512; i < n; i += 32) acc += (double)__ldcg(p + i);
        acc = warp_sum(acc);
        const double val = (t.kind == 1) ? acc : ((t.count > 0) ? acc / (double)t.count : CUDART_NAN);
        if (lane == 0) a.means[t.out] = val;
    }
    if (MODE != 0) {
        __shared__ bool s_last;
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned int prev = atomicAdd(&oa.sc->means_done, 1u);
            s_last = (prev == gridDim.x - 1);
            if (s_last) { oa.sc->means_done = 0u; __threadfence(); }
        }
        __syncthreads();
        if (!s_last) return;
        if (MODE == 2) {
            // exchange of the strip means, in this CTA: my slots to every peer, flags for the means (phase 1) and for the ghost
            // pixels the other CTAs pushed (phase 2; their system-scope fences precede the counter), then wait for the peers
            const P2PArgs& P = *fx.p2p;
            const unsigned int step = P.sc->step;
            // every local mean is read ONCE (four independent slot -> value chains per thread in flight) and stored to all peers:
            // the serial tail of this CTA does not grow with the number of ranks
            for (int i0 = threadIdx.x; i0 < a.n_tasks; i0 += blockDim.x * 4) {
                int slot[4]; double v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) { const int i = i0 + k * (int)blockDim.x; slot[k] = (i < a.n_tasks) ? a.tasks[i].out : -1; }
#pragma unroll
                for (int k = 0; k < 4; ++k) v[k] = slot[k] >= 0 ? ld_cg_f64(a.means + slot[k]) : 0.0;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (slot[k] >= 0)
                        for (int p = 0; p < P.world; ++p)
                            if (p != P.rank) P.means[p][slot[k]] = v[k];
            }
            __threadfence_system();
            __syncthreads();
            if ((int)threadIdx.x < P.world && (int)threadIdx.x != P.rank) {
                const int p = threadIdx.x;
                st_release_sys(&P.mail[p]->flag[1][P.rank], step);
                if (P.pix_send_ptr[p + 1] > P.pix_send_ptr[p]) st_release_sys(&P.mail[p]->flag[2][P.rank], step);
            }
            p2p_wait(fx.p2p, 1, 0xFFu, 1);           // only the last CTA waits here: it is the sample
        }
        offsets_body(oa, s_rec, s_terms, mk_smem);
    }
}
void launch_fold(const MeansArgs& a, const float* rowpart, const OffsetsArgs* fused, const P2PFused& fx, cudaStream_t s) {
    if (a.n_tasks <= 0) return;
    const int blocks = (a.n_tasks + 7) / 8;
    if (fused) {
        const size_t n = (size_t)fused->B * fused->F;
        const size_t smem = ((n * 24 + 15) & ~(size_t)15) + n * sizeof(DevRec) + (size_t)fused->term_start[fused->F] * sizeof(DevShiftTerm);
        if (fx.p2p) {
            static bool opted[64] = {};
            int dev = 0;
            cudaGetDevice(&dev);
            if (!opted[dev & 63]) { cudaFuncSetAttribute(task_fold_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); opted[dev & 63] = true; }
            long long want = (fx.n_pix_send * fx.C + 255) / 256;
            const int push = (int)(want > 32 ? 32 : want);
            launch_k(task_fold_kernel<2>, dim3(blocks + push), dim3(256), smem, s, a, rowpart, *fused, fx, blocks);
        } else {
            static bool opted[64] = {};
            int dev = 0;
            cudaGetDevice(&dev);
            if (!opted[dev & 63]) { cudaFuncSetAttribute(task_fold_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); opted[dev & 63] = true; }
            launch_k(task_fold_kernel<1>, dim3(blocks), dim3(256), smem, s, a, rowpart, *fused, fx, blocks);
        }
    }
    else launch_k(task_fold_kernel<0>, dim3(blocks), dim3(256), 0, s, a, rowpart, OffsetsArgs{}, fx, blocks);
}

// ------------------------------------------------------------------------------------------------
// K6b  offsets.  c_k = m[a] - (m[b] - c[parent])  is a sum along the parent chain of
//      d_k = m[a] - m[b] (roots: m[a] - Ref_BC), evaluated by pointer jumping in
//      ceil(log2(depth)) rounds; then the global shift of SMC:350 / GRAD:358-361 from the per-run
//      line sums.  Single CTA: the whole problem is a few thousand scalars.  In the multi-GPU path
//      every rank evaluates this redundantly on the all-reduced means.
__device__ void offsets_body(const OffsetsArgs& a, const DevRec* rec, const DevShiftTerm* terms, unsigned char* off_smem) {
    const int n = a.B * a.F;
    // pointer jumping in shared memory when the forest fits (n <= kOffsetsSmemMax), else in the global scratch
    const bool in_smem = n <= kOffsetsSmemMax;
    double* d_cur = in_smem ? reinterpret_cast<double*>(off_smem) : a.dbuf0;
    double* d_nxt = in_smem ? d_cur + n : a.dbuf1;
    int32_t* p_cur = in_smem ? reinterpret_cast<int32_t*>(d_nxt + n) : a.pbuf0;
    int32_t* p_nxt = in_smem ? p_cur + n : a.pbuf1;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const DevRec r = rec[i];
        const int f = i / a.B;
        double d = ld_cg_f64(a.means + r.ta) - (r.tb >= 0 ? ld_cg_f64(a.means + r.tb) : a.ref_bc);
        d_cur[i] = d;
        p_cur[i] = (r.parent >= 0) ? f * a.B + r.parent : -1;
    }
    __syncthreads();
    for (int round = 0; round < a.rounds; ++round) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int p = p_cur[i];
            double d = d_cur[i];
            int pn = -1;
            if (p >= 0) { d += d_cur[p]; pn = p_cur[p]; }
            d_nxt[i] = d; p_nxt[i] = pn;
        }
        __syncthreads();
        double* td = d_cur; d_cur = d_nxt; d_nxt = td;
        int32_t* tp = p_cur; p_cur = p_nxt; p_nxt = tp;
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) a.offsets[i] = d_cur[i];
    __shared__ double red[32];
    __shared__ double s_shift[2];
    for (int f = 0; f < a.F; ++f) {
        double acc = 0.0;
        for (int i = a.term_start[f] + threadIdx.x; i < a.term_start[f + 1]; i += blockDim.x) {
            const DevShiftTerm t = terms[i];
            acc += (double)t.coef * (ld_cg_f64(a.means + t.task) - (double)t.n * d_cur[f * a.B + t.block]);
        }
        acc = warp_sum(acc);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x < 32) {
            double t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
            t = warp_sum(t);
            if (threadIdx.x == 0) { s_shift[f] = t / (double)a.shift_len[f] / 3.0; a.sc->shift[f] = s_shift[f]; }
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) a.coff[i] = (float)(d_cur[i] + s_shift[i / a.B]);
    if (threadIdx.x == 0) {
        a.sc->umax2_bits = 0ull; a.sc->dumax2_bits = 0ull;      // re-arm the running maxima of prep
        a.sc->dense_barrier = 0u;                               // ... and the grid barrier of the Dense stack
        // status word in mapped host memory: visible to the host once the step's kernels have completed (stream sync)
        if (a.host_skip) *reinterpret_cast<volatile int*>(a.host_skip) = a.sc->skip | (a.sc->comm_error << 8);
    }
}
__global__ void __launch_bounds__(1024) offsets_kernel(OffsetsArgs a) {
    pdl_enter();
    if (a.p2p) p2p_wait(a.p2p, 1, 0xFFu);          // every rank's strip means have been pushed into a.means
    extern __shared__ unsigned char off_smem[];
    offsets_body(a, a.rec, a.terms, off_smem);
}
void launch_offsets(const OffsetsArgs& a, cudaStream_t s) {
    const int n = a.B * a.F;
    const size_t smem = n <= kOffsetsSmemMax ? (size_t)n * 24 : 0;
    static bool opted[64] = {};                                                    // the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!opted[dev & 63]) { cudaFuncSetAttribute(offsets_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kOffsetsSmemMax * 24); opted[dev & 63] = true; }
    launch_k(offsets_kernel, dim3(1), dim3(1024), smem, s, a);
}

// ------------------------------------------------------------------------------------------------
// K7  placement.  SMC:332-348 / GRAD:345-356 as a gather through the last-writer map, with the
//     block correction and the global shift subtracted on the fly.
__global__ void __launch_bounds__(256) place_kernel(PlaceArgs a) {
    pdl_enter();
    // 4 consecutive pixels of a row per thread (W % 4 == 0 fast path): one 8-byte owner load, and when the
    // four pixels share their owner and the block column is 16-byte aligned, one 128-bit block load
    const int W4 = a.W >> 2;
    const long long groups = (long long)a.H * W4;
    const long long total = groups * a.F;
    if ((a.W & 3) == 0) {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
            const int f = (int)(i / groups);
            const long long g = i - (long long)f * groups;
            const int y = (int)(g / W4), x = (int)(g - (long long)y * W4) << 2;
            const long long q = (long long)y * a.W + x;
            const ushort4 o4 = *reinterpret_cast<const ushort4*>(a.owner + q);
            float4 out;
            const int o = o4.x;
            const int lx = x - a.bx0[o];
            const float* brow = a.blocks + (((long long)o * a.C + f) * a.S + (y - a.by0[o])) * a.S;
            if (o4.y == o && o4.z == o && o4.w == o && (lx & 3) == 0) {
                const float4 v = *reinterpret_cast<const float4*>(brow + lx);
                const float c = a.coff[f * a.B_glob + a.kb0 + o];
                out = make_float4(v.x - c, v.y - c, v.z - c, v.w - c);
            } else {
                const int os[4] = {o4.x, o4.y, o4.z, o4.w};
                float r[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int ok = os[k];
                    r[k] = a.blocks[(((long long)ok * a.C + f) * a.S + (y - a.by0[ok])) * a.S + (x + k - a.bx0[ok])] -
                           a.coff[f * a.B_glob + a.kb0 + ok];
                }
                out = make_float4(r[0], r[1], r[2], r[3]);
            }
            *reinterpret_cast<float4*>(a.field + (long long)f * a.plane_stride + q) = out;
        }
        return;
    }
    const long long plane = (long long)a.H * a.W;
    const long long tot1 = plane * a.F;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot1; i += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(i / plane);
        const long long q = i - (long long)f * plane;
        const int y = (int)(q / a.W), x = (int)(q - (long long)y * a.W);
        const int o = a.owner[q];
        const float v = a.blocks[(((long long)o * a.C + f) * a.S + (y - a.by0[o])) * a.S + (x - a.bx0[o])];
        a.field[(long long)f * a.plane_stride + q] = v - a.coff[f * a.B_glob + a.kb0 + o];
    }
}
void launch_place(const PlaceArgs& a, cudaStream_t s) {
    long long total = (long long)a.H * a.W * a.F / (((a.W & 3) == 0) ? 4 : 1);
    long long want = (total + 255) / 256;
    int blocks = (int)(want < (long long)kSMs * 8 ? want : kSMs * 8);
    launch_k(place_kernel, dim3(blocks), dim3(256), 0, s, a);
}

// ------------------------------------------------------------------------------------------------
// Optional Gaussian post-filter, one separable pass.  SciPy's 'reflect' boundary (half-sample symmetric:
// d c b a | a b c d | d c b a), periodic with period 2n for kernels longer than the axis.
__device__ __forceinline__ int reflect_index(int i, int n) {
    const int p = 2 * n;
    i %= p;
    if (i < 0) i += p;
    return i < n ? i : p - 1 - i;
}
__global__ void __launch_bounds__(256) gauss_kernel(GaussArgs a) {
    pdl_enter();
    extern __shared__ float gw[];
    for (int i = threadIdx.x; i < 2 * a.radius + 1; i += blockDim.x) gw[i] = a.w[i];
    __syncthreads();
    const long long total = (long long)a.H * a.W;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(q / a.W), x = (int)(q - (long long)y * a.W);
        float acc = 0.f;
        if (a.axis == 0) {
            const bool inside = y - a.radius >= 0 && y + a.radius < a.H;
            for (int k = -a.radius; k <= a.radius; ++k) {
                const int yy = inside ? y + k : reflect_index(y + k, a.H);
                acc = fmaf(gw[k + a.radius], __ldg(a.in + (long long)yy * a.W + x), acc);
            }
        } else {
            const float* row = a.in + (long long)y * a.W;
            const bool inside = x - a.radius >= 0 && x + a.radius < a.W;
            for (int k = -a.radius; k <= a.radius; ++k) {
                const int xx = inside ? x + k : reflect_index(x + k, a.W);
                acc = fmaf(gw[k + a.radius], __ldg(row + xx), acc);
            }
        }
        a.out[q] = acc;
    }
}
void launch_gauss(const GaussArgs& a, cudaStream_t s) {
    long long want = ((long long)a.H * a.W + 255) / 256;
    int blocks = (int)(want < (long long)kSMs * 8 ? (want > 0 ? want : 1) : kSMs * 8);
    launch_k(gauss_kernel, dim3(blocks), dim3(256), (size_t)(2 * a.radius + 1) * sizeof(float), s, a);
}

// ------------------------------------------------------------------------------------------------
// K8  grid -> cell gather.  PMP:481-496: result[indices] (folded into the vertex ids at init),
//     interpolate_fill, previous-pressure fallback for NaN / near-wall cells; SMC:644-645
//     (p = p_prev + delta_p) for the deltaU variant.
template <bool FROM_BLOCKS>
__global__ void __launch_bounds__(256) back_kernel(BackArgs a) {
    pdl_launch_dependents();
    // static tables of the first cell before the wait (see gather_kernel)
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int i0 = 0, i1 = 0, i2 = 0; float q0 = 0.f, q1 = 0.f, q2 = 0.f;
    int b0 = 0, b1 = 0, b2 = 0;
    auto load = [&]() {
        i0 = __ldcs(a.v0 + i); i1 = __ldcs(a.v1 + i); i2 = __ldcs(a.v2 + i);
        q0 = __ldcs(a.w0 + i); q1 = __ldcs(a.w1 + i); q2 = __ldcs(a.w2 + i);
        if (FROM_BLOCKS) { b0 = __ldcs(a.o0 + i); b1 = __ldcs(a.o1 + i); b2 = __ldcs(a.o2 + i); }
    };
    if (i < a.n) load();
    pdl_wait();
    if (a.p2p) p2p_wait(a.p2p, 2, a.p2p->pix_recv_mask);   // ghost pixels pushed by their owners
    const int skip = a.sc->skip;
    const float* __restrict__ src = a.field;               // the assembled field, or the predicted blocks
    while (i < a.n) {
        if (a.n_fields == 1) {
            const double pp = a.p_prev[i];
            double out = pp;
            if (i0 >= 0 && !skip) {
                float f0 = __ldg(src + i0), f1 = __ldg(src + i1), f2 = __ldg(src + i2);
                if (FROM_BLOCKS) { f0 -= __ldg(a.coff + b0); f1 -= __ldg(a.coff + b1); f2 -= __ldg(a.coff + b2); }   // SMC:243,350
                const float v = f0 * q0 + f1 * q1 + f2 * q2;
                if (v == v) out = a.additive ? pp + (double)v : (double)v;
            }
            a.out[i] = out;
        } else {
            double o0 = CUDART_NAN, o1 = CUDART_NAN;
            if (i0 >= 0) {
                const float* s1 = src + a.plane;
                float f0 = __ldg(src + i0), f1 = __ldg(src + i1), f2 = __ldg(src + i2);
                float g0 = __ldg(s1 + i0), g1 = __ldg(s1 + i1), g2 = __ldg(s1 + i2);
                if (FROM_BLOCKS) {
                    const float* c1 = a.coff + a.n_blocks;
                    f0 -= __ldg(a.coff + b0); f1 -= __ldg(a.coff + b1); f2 -= __ldg(a.coff + b2);
                    g0 -= __ldg(c1 + b0); g1 -= __ldg(c1 + b1); g2 -= __ldg(c1 + b2);
                }
                o0 = (double)(f0 * q0 + f1 * q1 + f2 * q2);
                o1 = (double)(g0 * q0 + g1 * q1 + g2 * q2);
            }
            reinterpret_cast<double2*>(a.out)[i] = make_double2(o0, o1);
        }
        i += stride;
        if (i < a.n) load();
    }
}
// Four consecutive cells per thread: nine 64/128-bit streaming table loads (issued before the wait) instead of thirty-six
// scalar ones, two 128-bit p_prev loads, two 128-bit stores.  Per-cell arithmetic identical to back_kernel (bit-identical
// results); the last n % 4 cells go through the scalar expressions in CTA 0.
template <bool FROM_BLOCKS>
__device__ __forceinline__ double back_one(const BackArgs& a, const float* __restrict__ src, int i0, int i1, int i2, float q0, float q1, float q2,
                                           int b0, int b1, int b2, double pp, int skip) {
    // the gathers depend on the tables only: they are issued whatever the step's skip word says (it is applied to the result)
    double out = pp;
    if (i0 >= 0) {
        float f0 = __ldg(src + i0), f1 = __ldg(src + i1), f2 = __ldg(src + i2);
        if (FROM_BLOCKS) { f0 -= __ldg(a.coff + b0); f1 -= __ldg(a.coff + b1); f2 -= __ldg(a.coff + b2); }   // SMC:243,350
        const float v = f0 * q0 + f1 * q1 + f2 * q2;
        if (v == v && !skip) out = a.additive ? pp + (double)v : (double)v;
    }
    return out;
}
template <bool FROM_BLOCKS>
__device__ __forceinline__ double2 back_two(const BackArgs& a, const float* __restrict__ src, int i0, int i1, int i2, float q0, float q1, float q2,
                                            int b0, int b1, int b2) {
    double o0 = CUDART_NAN, o1 = CUDART_NAN;
    if (i0 >= 0) {
        const float* s1 = src + a.plane;
        float f0 = __ldg(src + i0), f1 = __ldg(src + i1), f2 = __ldg(src + i2);
        float g0 = __ldg(s1 + i0), g1 = __ldg(s1 + i1), g2 = __ldg(s1 + i2);
        if (FROM_BLOCKS) {
            const float* c1 = a.coff + a.n_blocks;
            f0 -= __ldg(a.coff + b0); f1 -= __ldg(a.coff + b1); f2 -= __ldg(a.coff + b2);
            g0 -= __ldg(c1 + b0); g1 -= __ldg(c1 + b1); g2 -= __ldg(c1 + b2);
        }
        o0 = (double)(f0 * q0 + f1 * q1 + f2 * q2);
        o1 = (double)(g0 * q0 + g1 * q1 + g2 * q2);
    }
    return make_double2(o0, o1);
}

template <bool FROM_BLOCKS>
__global__ void __launch_bounds__(256) back4_kernel(BackArgs a) {
    pdl_launch_dependents();
    const long long n4 = a.n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int4 v0 = make_int4(0, 0, 0, 0), v1 = v0, v2 = v0;
    float4 w0 = make_float4(0.f, 0.f, 0.f, 0.f), w1 = w0, w2 = w0;
    uint2 o0 = make_uint2(0u, 0u), o1 = o0, o2 = o0;
    auto load = [&]() {
        v0 = __ldcs(reinterpret_cast<const int4*>(a.v0) + i); v1 = __ldcs(reinterpret_cast<const int4*>(a.v1) + i);
        v2 = __ldcs(reinterpret_cast<const int4*>(a.v2) + i);
        w0 = __ldcs(reinterpret_cast<const float4*>(a.w0) + i); w1 = __ldcs(reinterpret_cast<const float4*>(a.w1) + i);
        w2 = __ldcs(reinterpret_cast<const float4*>(a.w2) + i);
        if (FROM_BLOCKS) {
            o0 = __ldcs(reinterpret_cast<const uint2*>(a.o0) + i); o1 = __ldcs(reinterpret_cast<const uint2*>(a.o1) + i);
            o2 = __ldcs(reinterpret_cast<const uint2*>(a.o2) + i);
        }
    };
    if (i < n4) load();
    pdl_wait();
    if (a.p2p) p2p_wait(a.p2p, 2, a.p2p->pix_recv_mask);   // ghost pixels pushed by their owners
    const int* skip_p = &a.sc->skip;
    const float* __restrict__ src = a.field;               // the assembled field, or the predicted blocks
    int skip = 0;
    bool have_skip = false;
    while (i < n4) {
        const int bA[4] = {(int)(o0.x & 0xFFFFu), (int)(o0.x >> 16), (int)(o0.y & 0xFFFFu), (int)(o0.y >> 16)};
        const int bB[4] = {(int)(o1.x & 0xFFFFu), (int)(o1.x >> 16), (int)(o1.y & 0xFFFFu), (int)(o1.y >> 16)};
        const int bC[4] = {(int)(o2.x & 0xFFFFu), (int)(o2.x >> 16), (int)(o2.y & 0xFFFFu), (int)(o2.y >> 16)};
        const int iA[4] = {v0.x, v0.y, v0.z, v0.w}, iB[4] = {v1.x, v1.y, v1.z, v1.w}, iC[4] = {v2.x, v2.y, v2.z, v2.w};
        const float qA[4] = {w0.x, w0.y, w0.z, w0.w}, qB[4] = {w1.x, w1.y, w1.z, w1.w}, qC[4] = {w2.x, w2.y, w2.z, w2.w};
        if (a.n_fields == 1) {
            const double2 pa = __ldcs(reinterpret_cast<const double2*>(a.p_prev) + 2 * i);
            const double2 pb = __ldcs(reinterpret_cast<const double2*>(a.p_prev) + 2 * i + 1);
            const double pp[4] = {pa.x, pa.y, pb.x, pb.y};
            // gathers first (they depend on the static tables only) ...
            float v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                v[k] = CUDART_NAN_F;
                if (iA[k] >= 0) {
                    float f0 = __ldg(src + iA[k]), f1 = __ldg(src + iB[k]), f2 = __ldg(src + iC[k]);
                    if (FROM_BLOCKS) { f0 -= __ldg(a.coff + bA[k]); f1 -= __ldg(a.coff + bB[k]); f2 -= __ldg(a.coff + bC[k]); }   // SMC:243,350
                    v[k] = f0 * qA[k] + f1 * qB[k] + f2 * qC[k];
                }
            }
            // ... then the step's skip word (once; a compiler barrier keeps its load behind the gathers: in-order issue would
            // otherwise park the whole warp on this one L2 round trip before any gather is in flight)
            if (!have_skip) {
                asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(skip) : "l"(skip_p) : "memory");
                have_skip = true;
            }
            double r[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) r[k] = (v[k] == v[k] && !skip) ? (a.additive ? pp[k] + (double)v[k] : (double)v[k]) : pp[k];
            __stcs(reinterpret_cast<double2*>(a.out) + 2 * i, make_double2(r[0], r[1]));
            __stcs(reinterpret_cast<double2*>(a.out) + 2 * i + 1, make_double2(r[2], r[3]));
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                __stcs(reinterpret_cast<double2*>(a.out) + 4 * i + k, back_two<FROM_BLOCKS>(a, src, iA[k], iB[k], iC[k], qA[k], qB[k], qC[k], bA[k], bB[k], bC[k]));
        }
        i += stride;
        if (i < n4) load();
    }
    if (blockIdx.x == 0 && (long long)threadIdx.x < (a.n & 3)) {     // the last n % 4 cells
        if (!have_skip) skip = *skip_p;
        const long long c = (n4 << 2) + threadIdx.x;
        const int i0 = a.v0[c], i1 = a.v1[c], i2 = a.v2[c];
        const float q0 = a.w0[c], q1 = a.w1[c], q2 = a.w2[c];
        int b0 = 0, b1 = 0, b2 = 0;
        if (FROM_BLOCKS) { b0 = a.o0[c]; b1 = a.o1[c]; b2 = a.o2[c]; }
        if (a.n_fields == 1) a.out[c] = back_one<FROM_BLOCKS>(a, src, i0, i1, i2, q0, q1, q2, b0, b1, b2, a.p_prev[c], skip);
        else reinterpret_cast<double2*>(a.out)[c] = back_two<FROM_BLOCKS>(a, src, i0, i1, i2, q0, q1, q2, b0, b1, b2);
    }
}

void launch_back(const BackArgs& a, cudaStream_t s) {
    static const bool scalar = [] { const char* v = getenv("PSM_BACK_SCALAR"); return v && v[0] == '1'; }();
    const bool aligned = ((reinterpret_cast<uintptr_t>(a.out) | reinterpret_cast<uintptr_t>(a.p_prev)) & 15) == 0;
    if (!scalar && aligned && a.n >= 4) {
        long long want = ((a.n >> 2) + 255) / 256;
        int blocks = (int)(want < (long long)kSMs * 6 ? (want > 0 ? want : 1) : kSMs * 6);
        if (a.o0) launch_k(back4_kernel<true>, dim3(blocks), dim3(256), 0, s, a);
        else launch_k(back4_kernel<false>, dim3(blocks), dim3(256), 0, s, a);
        return;
    }
    long long want = (a.n + 255) / 256;
    int blocks = (int)(want < (long long)kSMs * 8 ? (want > 0 ? want : 1) : kSMs * 8);
    if (a.o0) launch_k(back_kernel<true>, dim3(blocks), dim3(256), 0, s, a);
    else launch_k(back_kernel<false>, dim3(blocks), dim3(256), 0, s, a);
}

// Cell routing kernels (see psm_kernels.cuh): plain permutations of 8-byte words, one row per thread.
__global__ void __launch_bounds__(256) route_pack_kernel(RoutePackArgs a) {
    for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < a.n; s += (long long)gridDim.x * blockDim.x) {
        const long long i = a.perm[s];
        double* o = a.send + s * a.k;
        int c = 0;
        o[c++] = a.U[i * a.stride]; o[c++] = a.U[i * a.stride + 1];
        if (a.dU) { o[c++] = a.dU[i * a.stride]; o[c++] = a.dU[i * a.stride + 1]; }
        o[c] = a.p ? a.p[i] : 0.0;
    }
}
__global__ void __launch_bounds__(256) route_scatter_kernel(RouteScatterArgs a) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < a.n; r += (long long)gridDim.x * blockDim.x) {
        const long long j = a.idx[r];
        const double* in = a.recv + r * a.k;
        int c = 0;
        a.U2[2 * j] = in[c++]; a.U2[2 * j + 1] = in[c++];
        if (a.has_du) { a.dU2[2 * j] = in[c++]; a.dU2[2 * j + 1] = in[c++]; }
        a.p[j] = in[c];
    }
}
__global__ void __launch_bounds__(256) route_back_kernel(RouteBackArgs a) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < a.n; r += (long long)gridDim.x * blockDim.x) {
        const long long j = a.idx[r];
        for (int f = 0; f < a.F; ++f) {
            if (a.gather) a.dst[r * a.F + f] = a.src[j * a.F + f];
            else a.dst[j * a.F + f] = a.src[r * a.F + f];
        }
    }
}
static int route_blocks(long long n) { long long w = (n + 255) / 256; return (int)(w < 1 ? 1 : (w > kSMs * 8 ? kSMs * 8 : w)); }
void launch_route_pack(const RoutePackArgs& a, cudaStream_t s) { if (a.n > 0) route_pack_kernel<<<route_blocks(a.n), 256, 0, s>>>(a); }
void launch_route_scatter(const RouteScatterArgs& a, cudaStream_t s) { if (a.n > 0) route_scatter_kernel<<<route_blocks(a.n), 256, 0, s>>>(a); }
void launch_route_back(const RouteBackArgs& a, cudaStream_t s) { if (a.n > 0) route_back_kernel<<<route_blocks(a.n), 256, 0, s>>>(a); }

// Static sparse exchange (multi-GPU): contiguous send buffer from an index list.
__global__ void __launch_bounds__(256) pack_kernel(PackArgs a) {
    pdl_enter();
    const long long total = a.n * a.width;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long e = i / a.width; const int c = (int)(i - e * a.width);
        // src_stride == 0: interleaved records of `width` floats (dst interleaved too);
        // otherwise `width` planes src_stride apart (dst planar: [width][n])
        if (a.src_stride) a.dst[(long long)c * a.n + e] = a.src[(long long)c * a.src_stride + a.idx[e]];
        else a.dst[i] = a.src[(long long)a.idx[e] * a.width + c];
    }
}
void launch_pack(const PackArgs& a, cudaStream_t s) {
    if (a.n <= 0) return;
    long long want = (a.n * a.width + 255) / 256;
    launch_k(pack_kernel, dim3((int)(want < 1184 ? want : 1184)), dim3(256), 0, s, a);
}

}  // namespace psm
