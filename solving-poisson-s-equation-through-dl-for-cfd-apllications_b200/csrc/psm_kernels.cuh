// Device-side data structures and kernel launchers of the surrogate hot path (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace psm {

// Programmatic dependent launch: every kernel of the step is launched with the programmatic-serialisation
// attribute and starts with pdl_enter(): its CTAs may become resident while the previous kernel drains
// (launch latency and prologue overlap), and touch memory only after the previous kernel has completed.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_launch_dependents(); pdl_wait(); }
#endif
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// Per-step scalars kept on the device so that no host round-trip sits inside a step.
struct Scalars {
    unsigned long long umax2_bits;   // running max of Ux^2+Uy^2 (bit pattern of a non-negative double)
    unsigned long long dumax2_bits;  // running max of dUx^2+dUy^2
    double U_max_norm;               // SMC:404
    double dU_max_norm;              // SMC:405
    double shift[2];                 // global shift per field, SMC:350 / GRAD:358-361
    float  in_scale[2];              // 1 / (U_max_norm * max_abs_U{x,y})        SMC:418-419,441-442
    float  out_scale;                // max_abs_p * U_max_norm^2 (SMC:551) or 1 (GRAD:537-538)
    int    skip;                     // 1: irrelevant time step (SMC:410-415) or no previous U yet
    int    have_prev;                // 5-column deltaU mode: U(t-1) is resident
    unsigned int step;               // multi-GPU: id of the current step (flags of the peer-memory exchanges)
    int    comm_error;               // multi-GPU: a peer flag did not arrive in time
    unsigned int push_done[3];       // multi-GPU: CTAs of a push kernel that have finished their stores
    unsigned int means_done;         // CTAs of task_means_kernel that have published their mean (the last one runs the offsets)
    unsigned int dense_barrier;      // grid barrier of dense_stack_kernel (arrivals of the current step; re-armed by offsets_kernel)
    // multi-GPU: time CTA 0 of each consuming kernel spent in p2p_wait, per phase (0 ghost cells + maxima, 1 strip means,
    // 2 ghost pixels), and the number of waits -- the per-phase wait histogram psm_get_wait_ns reports
    unsigned long long wait_ns[3];
    unsigned int wait_n[3];
};

// ---- multi-GPU exchanges over peer memory (NVLink P2P, cudaIpc-mapped buffers) ---------------------------
// Every exchange of a step is a PUSH: the producing rank stores straight into the consumer's buffers
// (ghost cells, strip-mean slots, ghost pixels, the two running maxima), fences at system scope and then
// raises flag[phase][src] = step in the consumer's mailbox; the consuming kernel spins on its own mailbox
// before touching the data.  No NCCL kernel sits on the step's critical path.
constexpr int kMaxPeers = 8;
struct PeerMail {
    unsigned long long maxima[kMaxPeers][2];     // [src]: bit patterns of max|U|^2, max|dU|^2 of rank src
    unsigned int flag[3][kMaxPeers];             // [phase][src]: step id the pushed data belongs to
};
struct P2PArgs {
    int rank, world;
    PeerMail* mail[kMaxPeers];                   // mail[p]: rank p's mailbox (own one included)
    float2* uv_ghost[kMaxPeers];                 // where MY cells land in rank p's ghost-cell region
    double* means[kMaxPeers];                    // rank p's strip-mean array (global slots)
    float* field_ghost[kMaxPeers];               // where MY pixels land in rank p's ghost-pixel region (field 0)
    long long field_stride[kMaxPeers];           // rank p's floats per field plane
    long long cell_send_ptr[kMaxPeers + 1], pix_send_ptr[kMaxPeers + 1];
    unsigned int pix_recv_mask;                  // peers this rank receives ghost pixels from
    Scalars* sc;
    // fused flow (no separate push kernels, no assembled field): ghost pixels go straight from MY predicted blocks into the
    // ghost region behind rank p's blocks array, laid out [chunk][C][S*S] so that the channels of a slot are S*S apart
    float* blk_ghost[kMaxPeers];                 // rank p's ghost region (start)
    long long pix_slot0[kMaxPeers];              // first ghost slot of MY pixels in rank p's region
};
void launch_p2p_push_cells(const P2PArgs* d_pa, const float2* uv, const int32_t* send_idx, long long nsend, cudaStream_t s);
void launch_p2p_push_means(const P2PArgs* d_pa, const struct DevTask* tasks, int n_tasks, int world, const double* means, cudaStream_t s);
void launch_p2p_push_pix(const P2PArgs* d_pa, const float* field, const int32_t* send_idx, int F, long long my_stride, long long nsend, cudaStream_t s);
// fused multi-GPU flow: what the tails of prep / the fold kernel push
// A maximal run of consecutive owned cells [begin, end) that land consecutively in peer `peer`'s ghost-cell region from slot
// `dst` on: the ghost cells of a block-row shard are (nearly) whole cell rows next to a rank boundary, so a mesh numbered row
// by row has a handful of runs and every prep CTA can push the cells it has just converted itself (coalesced remote stores).
struct SendRun { int32_t begin, end, peer, pad; long long dst; };
constexpr int kMaxRuns = 32;
struct P2PFused {
    const P2PArgs* p2p;                          // NULL: single GPU / legacy flow
    const int32_t* cell_send_idx;                // owned cells that are ghost cells elsewhere (pushed by the last CTA when n_runs == 0)
    const SendRun* runs; int n_runs;             // the same list as runs (n_runs <= kMaxRuns), or n_runs = 0: too fragmented
    const float* blocks; const int32_t* pix_send_blk; int C, S2;   // fold kernel: ghost pixels from the predicted blocks (channel-0 offsets)
    long long n_pix_send;
    // send map (any cell order, any list length): word i/32 = {bit c set <=> owned cell 32*(i/32) + c is a ghost cell elsewhere,
    // number of marked cells before this word}; entry k (k = rank of the cell among the marked ones) = {peer | 0x100 if the cell has
    // a further entry | index of that entry << 9, slot in that peer's ghost region}.  Every prep thread looks its own cell up: one
    // 8-byte broadcast load per warp, a popcount, then the stores.
    const uint2* send_words; const int2* send_entries;
};

struct ScalarArgs {
    Scalars* sc;
    double max_abs_ux, max_abs_uy, out_scale_base;  // out_scale = base * (dimensionalise ? U^2 : 1)
    int dimensionalise;
    double skip_threshold;
    int mode;
    int replay;           // 1: re-evaluate the gather of the LAST step from the scales it published (psm_get_stage)
};

struct PrepArgs {
    const double* cells;  // [n][ncol]
    long long n;
    int ncol;
    int mode;             // 0: field = (col0, col1) [U_to_gradP]; 1: field = (col5, col6); 2: field = U - U_prev
    float2* uv;           // [n] field to interpolate (not yet scaled)
    double* p_prev;       // [n] column 4
    double* u_prev;       // [n][2] (mode 2)
    Scalars* sc;
    P2PFused fx;          // multi-GPU fused flow: the last CTA pushes maxima + ghost cells and raises the phase-0 flags
};
void launch_prep(const PrepArgs& a, cudaStream_t s);

// K0 for the solver's NATIVE field storage (psm_predict_fields): U as double[n][u_stride] (OpenFOAM `vector`: stride 3; only
// components 0, 1 are read), optional dU in the same layout.  p_prev is not touched here: the caller's p array is copied
// straight into the handle's p_prev buffer (concurrently with the kernels -- only the last kernel of the step reads it).
struct PrepFieldsArgs {
    const double* U; const double* dU;   // [n][stride]; dU may be NULL (mode 0 / 2)
    int stride; long long n;
    int mode;             // 0: field = U; 1: field = dU; 2: field = U - U_prev (resident)
    float2* uv; double* u_prev; Scalars* sc;
    P2PFused fx;
};
void launch_prep_fields(const PrepFieldsArgs& a, cudaStream_t s);


// K1: cell -> grid barycentric gather over the folded tables (SoA, padded to a multiple of 4).
struct GatherArgs {
    const int32_t* v0; const int32_t* v1; const int32_t* v2;
    const float* w0; const float* w1; const float* w2;
    const float2* uv;
    float* grid0; float* grid1;   // planes [H*W] (padded)
    long long n_pix4;             // number of 4-pixel groups
    ScalarArgs sa;                // thread 0 also publishes the step's scalars (U_max_norm, scales, skip rule)
    int scales_ready;             // 1: use the scales already published (replay of the last step, psm_get_stage)
    const P2PArgs* p2p;           // multi-GPU over peer memory: wait for the pushed maxima + ghost cells first
};
void launch_gather(const GatherArgs& a, cudaStream_t s);

// K1+K2 fused: the gathered 4-pixel group goes straight into every overlapping block's operand row
// (SMC:464-492), so the grid is never re-read.  Needs W % 4 == 0 and block columns on multiples of 4.
// rowcov[y] / colcov[x/4]: {count, up to 7 operand offsets}: a pixel group (y, xg) lands at
// xu + rowcov[y].off[r] + colcov[xg].off[c] for every covering block row r and block column c, where
// rowcov off = (first block of the row * 2 * S * S + local row * S) and colcov off = (position in the row * 2 * S * S + local x)
// -- both static, so the store addresses need no block-origin lookups.
struct CoverEntry { int32_t n; int32_t off[7]; };
struct GatherExtractArgs {
    GatherArgs g;
    const CoverEntry* rowcov;     // [rows gathered]
    const CoverEntry* colcov;     // [W/4]
    float* xu; int W4; int S;
    int store_grid;               // 0: only the block operand is written (the grid planes are rebuilt on demand by psm_get_stage)
    int skip_xu;                  // 1: the projection reads the grid planes itself (GridA): no block operand is written
};
void launch_gather_extract(const GatherExtractArgs& a, cudaStream_t s);

// K2: overlapping block extraction into the K-major operand x_u[B_pad][2*S*S] (planar c, ly, lx).
struct ExtractArgs {
    const float* grid0; const float* grid1;   // channel planes [H*W]; nch == 1 uses grid0 only
    const int32_t* by0; const int32_t* bx0;
    float* xu; int B; int W; int S; int nch;
};
void launch_extract(const ExtractArgs& a, cudaStream_t s);

// FP32 CUDA-core GEMM  C[M,N] = A[M,K] * B[N,K]^T  (both operands K-major), fused epilogues.
// Used for init-time constant folding and as the debug cross-check of the tensor-core path.
enum EpiKind { EPI_PARTIAL = 0, EPI_BIAS_RELU = 1, EPI_BIAS_AFFINE = 2, EPI_PCA_INV = 3, EPI_PLAIN = 4 };
struct GemmArgs {
    const float* A; const float* B; float* C;
    int M, N, K;          // M % 64 == 0, N % 64 == 0, K % 16 == 0 (buffers are padded at init)
    int lda, ldb, ldc;
    int splits;           // EPI_PARTIAL: split-K factor, C holds [splits][M][N]
    int epi;
    const float* v0;      // bias[n] / pca mean[n]
    const float* v1;      // scale[n]
    const float* v2;      // shift[n]
    const Scalars* sc;    // EPI_PCA_INV: out_scale
};
void launch_sgemm(const GemmArgs& a, cudaStream_t s);

// tcgen05 / TMEM / TMA GEMM (psm_gemm_tc.cu): same contract as GemmArgs, operands described by
// TMA tensor maps built once at init.  N % 64 == 0, K % 32 == 0, M % 128 == 0.
struct alignas(64) TensorMap128 { unsigned char bytes[128]; };
struct TcGemmArgs {
    float* C; int M, N, K, ldc;
    int splits, epi;
    int three_pass;       // 1: hi/lo split (3xTF32, FP32-class accuracy), 0: single-pass TF32
    const float* v0; const float* v1; const float* v2;
    const Scalars* sc;
    float* C_hi; float* C_lo;  // dense_cluster_kernel: optional tf32 hi / lo split of the result (same layout as C)
    int b_static;              // 1: the B operand does not depend on earlier kernels of the step (weights): its first tiles are
                               //    requested BEFORE the programmatic-dependent-launch wait
    unsigned long long* trace; // dense_cluster_kernel, PSM_TRACE_LAYERS=1: [<= 64 CTAs][8] globaltimer stamps (entry, after the wait,
                               //    operands landed, accumulator ready, partial pushed, cluster barrier passed, stored)
    int presplit;              // dense_cluster_kernel, three_pass == 2: the lo halves of both operands exist in global memory (the
                               //    previous layer's epilogue / the parameter load wrote them): TMA stages them, no converter pass
};
struct TcGemm { TensorMap128 mapA, mapB; TcGemmArgs args; int bn; TensorMap128 mapAlo, mapBlo; };   // bn: N tile (64 or 128) the B map was built for

// PCA projection with the A operand read straight from the interpolated grid planes (SMC:464-492): the block
// extraction is a TMA address pattern, x_array is never materialised.  Blocks whose origins form an arithmetic progression (step =
// stride pixels along x, or along y) are fetched as ONE box of a 5-D tensor map over the planes -- dims (x, bx, y, by, channel)
// with byte strides (4, 4 stride, 4 W, 4 W stride, plane), i.e. overlapping windows -- so an A tile (128 block rows x 32 pixels of
// one block-local pixel row) is a handful of boxes.  The rows of a tile are in "box order"; ReduceArgs::row_src maps block -> row.
struct ASeg { int32_t map; int32_t row; int32_t x, y; };       // map: 0 row box, 1 column box, 2 single block; row: first tile row; origin
struct GridA {
    const ASeg* segs; const int32_t* seg_ptr;   // segments of tile t: [seg_ptr[t], seg_ptr[t+1])
    const int32_t* a_bytes;                     // bytes the segments of tile t deliver per k-block
    int S;                                      // block side (k-block kb <-> channel kb / (S*S/32), row (kb % (S*S/32)) / (S/32), chunk kb % (S/32))
};
struct TcGemmGrid { TensorMap128 mapRow, mapCol, mapOne, mapB; TcGemmArgs args; GridA ga; int tiles; };
int make_grid_maps(TcGemmGrid* out, const float* planes, int W, int H, long long plane_stride, int stride_px, int gx, int gy);
int tc_gemm_grid_prepare();
void launch_tc_gemm_grid(const TcGemmGrid& t, cudaStream_t s);
int make_kmajor_map(TensorMap128* out, const float* ptr, int rows, int cols, int ld, int box_rows);
int tc_gemm_bn(int N);
int tc_gemm_prepare();
void launch_tc_gemm(const TcGemm& t, cudaStream_t s);

// PCA projection in ONE launch (psm_gemm_tc.cu): split-K over about one wave of CTAs, the `ks` CTAs of a thread-block cluster fold
// their partial accumulators through distributed shared memory (row slab z of the tile goes to CTA z), the folded cluster partial
// goes to `part`, and for every row slab the LAST cluster to arrive (one counter per slab) adds the `ncl` cluster partials in a
// fixed order, adds the static distance-channel term and standardises (SMC:494,505-523).  Replaces tc_gemm_kernel + the
// reduce_standardise launch and their [splits][M][N] partials in HBM.  N == 128, M % 128 == 0, K % 32 == 0.
struct ProjArgs {
    int M, N, K;
    int splits, ks;               // K splits (a multiple of ks) and the cluster size along z (1, 2, 4 or 8)
    int three_pass, b_static;
    float* part;                  // [splits / ks][M][N] cluster partials
    unsigned int* counters;       // [M / 128][ks], zero between steps (re-armed by the last arrival)
    const float* zc; const float* a; const float* b;   // x = (sum + zc[m][n]) * a[n] + b[n]
    float* x;                     // [M][N]
    float* x_lo;                  // optional: x - tf32(x) (pre-split operand of the first Dense layer)
};
struct ProjGemm { TensorMap128 mapA, mapB; ProjArgs args; };
int proj_cluster_prepare(int tiles, int ks, int* max_clusters);
int launch_proj_cluster(const ProjGemm& t, cudaStream_t s);

// Dense layer in one launch: cluster split-K + distributed-shared-memory reduction + fused epilogue
// (psm_gemm_tc.cu).  t.args.splits = cluster size along K (1, 2, 4 or 8); N % 64 == 0.
int dense_cluster_prepare();
int launch_dense_cluster(const TcGemm& t, cudaStream_t s);

// PCA inverse, transposed (psm_gemm_tc.cu): one CTA per 128 output pixels, all blocks as the MMA N dimension.
// Masked strip sums out of the PCA-inverse epilogue: every CTA of pca_inverse_t_kernel holds one pixel row
// (channel c, local row ly) of ALL predicted blocks in registers, so the row partial of every masked mean / line sum that
// covers that row (SMC:233-316, GRAD:300-340, SMC:350) is one warp reduction away -- the blocks are not re-read.
// Entries are static (built once per mesh), grouped by CTA row and sorted by source block.
struct StripRows {
    const int32_t* row_ptr;   // [C*S][ceil(B/32) + 1]: entries of pixel row r = c*S + ly whose source block lies in 32-block chunk k are
                              // [row_ptr[r][k], row_ptr[r][k+1]); NULL: disabled
    const int32_t* src;       // [n_ent] local source block (ascending within a row)
    const int32_t* slot;      // [n_ent] row-partial slot
    const uint32_t* w;        // [4][n_ent] lane masks per warp quarter: bit l of w[q] <=> pixel lx = 32 q + l counts
    int n_ent;
    float* rowpart;           // [n_slots][4] FP32 partial sums per (task, row, warp quarter); slots never counted stay 0
};
struct InvTArgs {
    float* blocks;            // [B][C][S][S]: element (block b, planar pixel P) at b * block_stride + P
    long long block_stride;   // C*S*S
    int B, Mb;                // real / padded (multiple of 128) block count
    int K;                    // padded pc_p (<= pca_inverse_t_max_k())
    int n_pix;                // C*S*S (multiple of 128)
    int three_pass;
    const float* pmean;       // [n_pix] planar
    const Scalars* sc;        // out_scale
    StripRows strips;
};
struct InvT { TensorMap128 mapA, mapBhi, mapBlo; InvTArgs args; };
int pca_inverse_t_prepare();
int pca_inverse_t_max_k();
void launch_pca_inverse_t(const InvT& t, cudaStream_t s);

// The whole Dense stack (NNS:8-38: Dense(relu) x (n-1) -> Dense(linear), de-standardisation SMC:533) in ONE launch
// (psm_gemm_tc.cu).  Persistent clusters of 8 CTAs: a cluster evaluates one 128 x 64 output tile of the current
// layer, its CTAs split K, park their partial accumulators in shared memory and fold them through distributed
// shared memory (fixed order); bias / ReLU / affine fused; activations are written pre-split into hi = tf32(x) and
// lo = x - hi so that the next layer's 3xTF32 passes read them straight through TMA (weights are split at load).
// Layers are separated by a grid barrier (all CTAs are co-resident: grid <= what the device holds at once).
constexpr int kMaxDense = 16;
struct DenseLayerDesc {
    int K, N, epi;                    // padded widths (multiples of 128); EPI_BIAS_RELU or EPI_BIAS_AFFINE
    const float* bias; const float* v1; const float* v2;
    float* out;                       // fp32 result [M][N] (may be NULL for hidden layers)
    float* out_hi; float* out_lo;     // split result [M][N] read by the next layer (may be NULL after the last)
};
struct DenseStackArgs {
    const TensorMap128* maps;         // DEVICE memory, [n_layers][4]: A_hi, A_lo (box 128 x 32), W_hi, W_lo (box 64 x 32)
    int n_layers, M, three_pass;
    unsigned int* barrier;            // zero before the launch
    int* error;                       // set to 1 if the grid barrier timed out (never on a healthy device)
    unsigned long long* trace;        // optional (debug): [CTA][64] globaltimer stamps
    DenseLayerDesc L[kMaxDense];
};
int dense_stack_prepare(int* max_clusters);
int launch_dense_stack(const DenseStackArgs& a, int clusters, cudaStream_t s);

// Split-K reduction + per-block constant + standardisation (SMC:494,512).
// Also the split-K epilogue of the Dense layers: relu(sum + bias[n]) or (sum + bias[n]) * s[n] + m[n].
enum ReduceKind { RED_STANDARDISE = 0, RED_BIAS_RELU = 1, RED_BIAS_AFFINE = 2 };
struct ReduceArgs {
    const float* part; int splits; int M; int N;
    const float* zc;      // RED_STANDARDISE: [M][N] static sdf-channel contribution per block
    const float* a; const float* b;   // STANDARDISE: x = (sum + zc) * a[n] + b[n]; AFFINE: scale a[n], shift b[n]
    float* x;
    int kind;
    const float* bias;    // RED_BIAS_*: [N]
    float* x_hi; float* x_lo;   // optional: the result split into tf32(x) and x - tf32(x) (operand of dense_stack_kernel)
    const int32_t* row_src;     // optional: output row m sums partial row row_src[m] (< 0: nothing) of [splits][part_rows][N]
    int part_rows;
};
void launch_reduce_standardise(const ReduceArgs& a, cudaStream_t s);

// K6a: masked rectangle means (and the plain line sums of the global shift) of the predicted blocks.
struct DevTask {
    int32_t src;                      // LOCAL block the values come from
    int32_t kind;                     // 0: masked mean, 1: plain sum (shift-line run), 2: plain mean (no mask, PMP:437)
    int32_t ch, y0, y1, x0, x1;       // channel, block-local rectangle
    int32_t count;                    // mask pixels in the rectangle (0 -> NaN mean)
    int32_t my0, mx0;                 // GLOBAL grid origin of the block whose flow mask applies
    int32_t out;                      // slot in the GLOBAL means array
    int32_t part_base;                // first row-partial slot of this task (rows y0..y1-1 follow), see StripRows
};
struct MeansArgs {
    const DevTask* tasks; int n_tasks;        // tasks evaluated by this rank
    const float* blocks;              // [B_loc_pad][C][S][S]
    const uint8_t* gmask;             // GLOBAL [H][W]
    int C, S, W;
    double* means;                    // GLOBAL slots (this rank writes only its own)
};

// K6b: offset recurrence (pointer jumping over the parent forest) + global shift from the line sums.
struct DevRec { int32_t ta, tb, parent, is_nan; };
struct DevShiftTerm { int32_t task, block, coef, n; };     // coef * (means[task] - n * c[block])
struct OffsetsArgs {
    const DevRec* rec; int B; int F; int rounds; double ref_bc;
    const double* means;
    double* dbuf0; double* dbuf1; int32_t* pbuf0; int32_t* pbuf1;   // scratch [F*B]
    double* offsets;                  // [F][B]  c_k
    float* coff;                      // [F][B]  c_k + shift_f (consumed by the placement)
    const DevShiftTerm* terms; int term_start[3]; int shift_len[2];
    Scalars* sc;
    int* host_skip;                   // mapped pinned host word: the step's status without a copy node
    const P2PArgs* p2p;               // multi-GPU over peer memory: wait for every rank's strip means first
};
void launch_offsets(const OffsetsArgs& a, cudaStream_t s);
// `fused` != NULL (single GPU, B*F <= 1024): the last CTA of the means kernel also runs the offsets (no second launch)
void launch_means(const MeansArgs& a, const OffsetsArgs* fused, cudaStream_t s);
// K6a': the same means from the row partials the PCA-inverse epilogue left (StripRows): one warp per task, FP64, fixed order
// fx.p2p != NULL (multi-GPU fused flow, `fused` required): extra CTAs push the ghost pixels, the last CTA pushes this rank's
// means, raises the phase-1 / phase-2 flags, waits for every peer's means and runs the offsets.
void launch_fold(const MeansArgs& a, const float* rowpart, const OffsetsArgs* fused, const P2PFused& fx, cudaStream_t s);

// K7: placement through the owner map.
struct PlaceArgs {
    const float* blocks; const uint16_t* owner; const int32_t* by0; const int32_t* bx0;   // LOCAL blocks / rows
    const float* coff;                // GLOBAL [F][B_glob]
    float* field;                     // [F][plane_stride]
    int B_glob, kb0, C, F, S, H, W;   // H = local rows placed; kb0 = first global block of this rank
    long long plane_stride;
};
void launch_place(const PlaceArgs& a, cudaStream_t s);

// Optional post-filter (SMC:353-356 / GRAD:366-367): scipy.ndimage.gaussian_filter(field, sigma, mode='reflect'),
// i.e. correlate1d along axis 0, then along axis 1, with the normalised kernel exp(-x^2 / 2 sigma^2), |x| <= radius.
struct GaussArgs {
    const float* in; float* out; int H, W;
    const float* w;       // [2 * radius + 1] weights (computed in FP64 on the host like SciPy, rounded to FP32)
    int radius; int axis; // 0: along y, 1: along x
};
void launch_gauss(const GaussArgs& a, cudaStream_t s);

// K8: grid -> cell gather, fallbacks (PMP:481-496).
struct BackArgs {
    const int32_t* v0; const int32_t* v1; const int32_t* v2;   // flat pixel ids, v0 < 0: keep p_prev
    const float* w0; const float* w1; const float* w2;
    const float* field; long long n; long long plane;           // plane = H*W
    const double* p_prev; double* out;
    int n_fields; int additive;
    const Scalars* sc;
    const P2PArgs* p2p;               // multi-GPU over peer memory: wait for the pushed ghost pixels first
    // K7 folded into K8 (single GPU): v0..v2 index the predicted BLOCKS (pixel -> last-writer block and its local
    // offset are static) and the block correction is subtracted on the fly -- the field is never materialised
    const uint16_t* o0; const uint16_t* o1; const uint16_t* o2;   // owner block of each vertex pixel (NULL: gather from `field`)
    const float* coff;                // [F][n_blocks]
    int n_blocks; int block_plane;    // S*S: distance between the channels of one block
};
void launch_back(const BackArgs& a, cudaStream_t s);

// U_to_gradP pressure recovery (psm_integrate.cu): GRAD:371-416 integrate_field on four quadrants + the stitch GRAD:585-628.
struct IntegrateFix { int32_t n; int32_t pos[4]; int32_t prev[4]; };     // per block-local row: entries the reference's "reset" overwrites
struct IntegrateArgs {
    const float* dpdx; const float* dpdy;     // assembled gradient fields [H][W]
    const uint8_t* mask;                      // [H][W] sdfunct != 0
    const IntegrateFix* fix;                  // [max(cy, H - cy)]
    int H, W, cx, cy;                         // centre column / row (GRAD:591-592)
    double dx, dy;                            // np.diff(xl)[0], np.diff(yl)[0]
    double* sdpx;                             // scratch [2][H][W]
    double* anchor;                           // scratch [4][H]
    double* corr;                             // scratch [2]
    int* status;                              // [1]: 1 = the two stitch masks have different counts (numpy would raise)
    double* out;                              // [H][W]
};
void launch_integrate(const IntegrateArgs& a, cudaStream_t s);

// Cell routing (psm_predict_routed): a solver rank holds an arbitrary subset of the cells; rows travel to the rank whose block rows
// contain them and the pressures travel back.  pack: send[s][:] = {U[perm[s]].x, .y, (dU.x, .y,) p[perm[s]]}; scatter: the received
// row r becomes cell idx[r] of the handle's native-field buffers; the two reverse kernels move the result.
struct RoutePackArgs { const double* U; const double* dU; const double* p; int stride; const int32_t* perm; long long n; int k; double* send; };
struct RouteScatterArgs { const double* recv; const int32_t* idx; long long n; int k; int has_du; double* U2; double* dU2; double* p; };
struct RouteBackArgs { const double* src; const int32_t* idx; long long n; int F; double* dst; int gather; };   // gather: dst[r] = src[idx[r]]; else dst[idx[r]] = src[r]
void launch_route_pack(const RoutePackArgs& a, cudaStream_t s);
void launch_route_scatter(const RouteScatterArgs& a, cudaStream_t s);
void launch_route_back(const RouteBackArgs& a, cudaStream_t s);

// Static sparse exchange (multi-GPU): dst[i] = src[idx[i]] for the elements other ranks need.
struct PackArgs { const float* src; const int32_t* idx; float* dst; long long n; int width; long long src_stride; };
void launch_pack(const PackArgs& a, cudaStream_t s);

}  // namespace psm
