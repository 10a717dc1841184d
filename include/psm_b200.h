/* psm_b200.h -- C-ABI of the B200-native pressure-surrogate hot path.
 *
 * Drop-in boundary for the per-timestep surrogate call that the reference makes through an
 * embedded CPython interpreter.  Each entry point names the reference interface it replaces
 * (paths relative to the reference root):
 *
 *   FOAM = Thesis_Work/Chapter5/parallelized/DLPoissonSolver
 *   PMP  = Thesis_Work/Chapter5/parallelized/test_case/python_module.py
 *   SMC  = Improved_SM/deltaU_to_deltaP/source/pressureSM_deltas/SM_call.py
 *   GRAD = Improved_SM/U_to_gradP/evaluation/Eval_dual_Dense_onlycil.py
 *
 * Conventions: plain pointers and sizes only; every function returns an int status
 * (0 = PSM_OK, >0 informational, <0 error) and never throws; psm_last_error() gives the
 * message of the last failure on a handle (or of the last failed psm_create when handle is
 * NULL).  A handle owns all device state, is bound to one CUDA device and one stream, and is
 * not thread-safe.  No allocation happens inside psm_predict*.
 */
#ifndef PSM_B200_H
#define PSM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSM_API_VERSION 5

typedef struct psm_handle psm_handle;

enum psm_status_code {
    PSM_OK = 0,
    PSM_SKIPPED = 1,          /* 'irrelevant' time step, SMC:407-415: p_out = p_prev            */
    PSM_ERR_INVALID = -1,     /* bad argument                                                    */
    PSM_ERR_CUDA = -2,        /* CUDA runtime failure (no device, out of memory, launch error)   */
    PSM_ERR_GEOMETRY = -3,    /* grid on which the reference itself is undefined (p_i == 0, ...) */
    PSM_ERR_STATE = -4,       /* call order (predict before init, ...)                           */
    PSM_ERR_COMM = -5         /* NCCL failure in the multi-GPU path                              */
};

enum psm_variant_code {
    PSM_DELTAU_TO_DELTAP = 0, /* SMC: right->left plan, 1 output channel, p = p_prev + delta_p   */
    PSM_U_TO_GRADP = 1,       /* GRAD: left->right plan, 2 output channels {dp/dx, dp/dy}        */
    PSM_THESIS_U_TO_P = 2     /* PMP (the solver module the reference ships, python_module.py): U -> p, shared columns
                                 `avance` = 12 (psm_config.overlap), extra left-most block column, scalar max-abs PCA scaling,
                                 distance channel unscaled (PMP:292), forward interpolation without the NaN fill (PMP:64-65),
                                 p_out = p (not p_prev + dp).  Single-GPU handles only.             */
};

enum psm_standardization_code {
    PSM_STD = 0,              /* (z - mean_in) / std_in ; r * std_out + mean_out   SMC:505-512,532-533 */
    PSM_MAX_ABS = 1,          /* z / max_abs_input_PCA ; r * max_abs_output_PCA    SMC:521-523, GRAD:525,531 */
    PSM_MIN_MAX = 2           /* (z - min_in) / (max_in - min_in) ; r * (max_out - min_out) + min_out   SMC:513-520,535-536.
                                 min_* travel in the mean_* slots of psm_params, max_* in the std_* slots                 */
};

/* Constants the reference hard-codes or takes from argparse (EP:89-98; PMP:195,303-304). */
typedef struct psm_config {
    int32_t variant;          /* psm_variant_code                                                */
    int32_t device;           /* CUDA device ordinal                                             */
    double  delta;            /* grid spacing, 5e-3                                              */
    int32_t shape;            /* block edge S, 128 (only 128 is supported)                       */
    int32_t overlap;          /* SMC `overlap` (32) or GRAD `avance` (96): shared columns of
                                 neighbouring blocks; stride = shape - overlap                   */
    int32_t input_cols;       /* 5: {Ux,Uy,Cx,Cy,p} as FOAM/PythonComm.H:2-9 fills it;
                                 7: + {dUx,dUy} (deltaU_to_deltaP only).  With 5 columns the
                                 deltaU variant keeps U(t-1) resident and forms dU on device.   */
    int32_t additive;         /* deltaU variant: 1 -> p_out = p_prev + delta_p (SMC:644-645),
                                 0 -> raw delta_p with p_prev fallback                           */
    double  ref_bc;           /* outlet reference value, SMC:570 (0)                             */
    double  skip_threshold;   /* SMC:410-411 (1e-4); <= 0 disables the skip rule                 */
    double  near_wall_sdf;    /* PMP:492-494: cells with interpolated distance < this keep
                                 p_prev (0.05 in PMP); <= 0 disables (PMS:430-432)               */
    int32_t enable_timings;   /* 1 -> record per-stage CUDA events (psm_get_timings)             */
    int32_t gemm_mode;        /* psm_gemm_mode_code; 0 (default) = tcgen05 3xTF32                */
    double  filter_sigma;     /* > 0: Gaussian post-filter of the assembled field(s), the reference's
                                 `apply_filter` (SMC:353-356, GRAD:366-367: scipy.ndimage.gaussian_filter,
                                 sigma (10, 10), mode 'reflect', truncate 4); 0 = off (EP:105 default).
                                 Single-GPU handles only.                                        */
} psm_config;

/* How the three dense contractions (PCA projection, Dense stack, PCA inverse) are evaluated.
 * All modes are this library's own sm_100a kernels; 1 and 2 exist for cross-checking mode 0. */
enum psm_gemm_mode_code {
    PSM_GEMM_TC_3XTF32 = 0,   /* tcgen05.mma kind::tf32, hi/lo operand split: FP32-class accuracy  */
    PSM_GEMM_TC_TF32 = 1,     /* tcgen05.mma kind::tf32, single pass (10-bit mantissa operands)    */
    PSM_GEMM_FP32_SIMT = 2    /* CUDA-core FP32 FFMA kernel (debug cross-check)                   */
};

/* Artefacts the reference loads at import time (SMC:70-87,505-511; PMP:103-118,168-170).
 * PCA matrices are in the reference's own layout: row = component, column index
 * k = (ly*S + lx)*n_channels + c  (channel-last blocks flattened, SMC:491-492, GRAD:513-516). */
typedef struct psm_params {
    double  maxs[5];              /* max_abs_Ux, Uy, dist, p [, second output]   SMC:70-72, GRAD:49-52 */
    int32_t pc_in;                /* kept input components  (SMC:87)                            */
    int32_t pc_p;                 /* kept output components (SMC:86)                            */
    int32_t n_out_channels;       /* 1 (deltaU_to_deltaP) or 2 (U_to_gradP)                     */
    int32_t standardization;      /* psm_standardization_code                                   */
    const double* pca_in_components;   /* [pc_in][S*S*3]                                        */
    const double* pca_in_mean;         /* [S*S*3]                                               */
    const double* pca_out_components;  /* [pc_p][S*S*n_out_channels]                            */
    const double* pca_out_mean;        /* [S*S*n_out_channels]                                  */
    const double* mean_in;  const double* std_in;    /* [pc_in]  PSM_STD: mean, std;  PSM_MIN_MAX: min, max  */
    const double* mean_out; const double* std_out;   /* [pc_p]   PSM_STD: mean, std;  PSM_MIN_MAX: min, max  */
    double  max_abs_input_PCA;    /* PSM_MAX_ABS                                                */
    double  max_abs_output_PCA;
    int32_t n_dense;              /* Dense layers incl. the linear output layer (4 for MLP_small, UTL:437-439) */
    int32_t reserved;
    const int32_t* layer_dims;    /* [n_dense+1]: pc_in, 512, 512, 512, pc_p                    */
    const float* const* dense_kernels;  /* n_dense pointers, Keras layout [in][out] (NNS:24-33) */
    const float* const* dense_biases;   /* n_dense pointers, [out]                              */
} psm_params;

/* Interpolation tables and raster built once per mesh (SMC:110-178; PMP:203-243).  The Python
 * shim builds them with SciPy's Qhull -- the library the reference calls -- so they are
 * bit-identical to the reference's; psm_init_with_tables copies them to the device. */
typedef struct psm_tables {
    int64_t n_cells;
    int32_t grid_h, grid_w;       /* grid_shape_y, grid_shape_x  (SMC:148-149)                  */
    const int32_t* vert;          /* [H*W][3] cell ids of the enclosing simplex   (UTL:40)      */
    const double*  weights;       /* [H*W][3] barycentric weights                 (UTL:42-44)   */
    const int32_t* vert_back;     /* [n_cells][3] grid-point ids (PMP:211); NULL -> grid output only */
    const double*  weights_back;  /* [n_cells][3]                                               */
    const int64_t* indices;       /* [H*W][2] (ii, jj) raster, (0,0) for invalid points (SMC:161-178) */
    const double*  sdfunct;       /* [H][W] distance field, 0 outside the flow domain (SMC:163,175) */
} psm_tables;

/* One rank's share of a block-row partitioned mesh (multi-GPU, replaces the gather-to-root of
 * PMP:179-185,258,501-511).  Built by psm_b200/shard.py (partition() from global tables, or band tables
 * for meshes too large to triangulate on one host).  Tables are LOCAL and PRE-FOLDED:
 *   - forward table rows cover pixel rows [row0,row1) only; invalid pixels have zero weights; the
 *     (0,0)-pixel quirk (SMC:161,432) is already applied on the rank that holds pixel (0,0);
 *   - cell ids are local: [0,n_owned) owned cells (the rows the caller passes to psm_predict, and the
 *     cells it gets pressures for), [n_owned, n_owned+n_ghost) ghost cells received from their owners;
 *   - back table ids are local field ids: (row-row0)*W+col for own pixels, n_pix_own+slot for ghost
 *     pixels received from their owners; vert_back[c][0] = -1 keeps p_prev (PMP:492-496). */
typedef struct psm_shard {
    int32_t rank, world;
    int32_t grid_h, grid_w;            /* GLOBAL grid                                                */
    int32_t row0, row1;                /* pixel rows this rank gathers and places                    */
    int32_t ext_rows;                  /* overlap rows [row1,row1+ext_rows) received from rank+1
                                          before block extraction (the halo strip); 0 on the last   */
    int32_t send_rows;                 /* rows [row0,row0+send_rows) sent to rank-1; 0 on rank 0     */
    int32_t blk_row0, blk_row1;        /* block rows [blk_row0,blk_row1) of the global plan          */
    int32_t local_ext_rows;            /* overlap rows [row1,row1+local_ext_rows) this rank gathers ITSELF
                                          (the forward table covers them; their cells arrive as ghost
                                          cells) -- the halo strip travels in cell space; use either this
                                          or ext_rows                                                  */
    const uint8_t* mask_global;        /* [H][W] flow mask (sdfunct != 0) of the WHOLE grid          */
    int64_t n_owned, n_ghost, n_ghost_pix;
    const int32_t* vert;               /* [(row1-row0+local_ext_rows)*W][3]                          */
    const double*  weights;            /* [(row1-row0+local_ext_rows)*W][3]                          */
    const double*  sdfunct;            /* [row1-row0+local_ext_rows+ext_rows][W]                     */
    const int32_t* vert_back;          /* [n_owned][3] or NULL                                       */
    const double*  weights_back;       /* [n_owned][3]                                               */
    /* static sparse exchanges, CSR over peer ranks ([world+1] offsets) */
    const int64_t* cell_send_ptr; const int32_t* cell_send_idx;   /* owned cell ids each peer needs   */
    const int64_t* cell_recv_ptr;                                 /* ghost slots, grouped by owner    */
    const int64_t* pix_send_ptr;  const int32_t* pix_send_idx;    /* own pixel ids each peer needs    */
    const int64_t* pix_recv_ptr;                                  /* ghost pixel slots, by owner      */
    const int64_t* ghost_pix;          /* [n_ghost_pix] GLOBAL pixel id (row * W + col) of every ghost pixel slot, or NULL.
                                          With it (and the peer-memory transport) the grid->cell gather reads the owners' predicted
                                          BLOCKS directly -- ghost pixels are pushed from the blocks, no field is assembled      */
} psm_shard;

/* Static geometry report (filled by psm_get_geometry). */
typedef struct psm_geometry {
    int32_t grid_h, grid_w, shape, overlap;
    int32_t n_x, n_y, p_i, p_j;   /* SMC:461-462,213,216 / GRAD:479-480,277-278                 */
    int32_t n_blocks, n_fields;
    int64_t n_cells;
    int32_t n_tasks;              /* masked strip means evaluated per step (whole mesh)         */
    int32_t peer_memory_exchange; /* multi-GPU: 1 = exchanges pushed over cudaIpc-mapped peer memory
                                     (NVLink), 0 = NCCL send/recv + all-reduce (PSM_COMM=nccl)        */
    /* this rank's share (equal to the whole mesh on a single-GPU handle) */
    int32_t row0, row1, ext_rows, first_block, n_local_blocks, world;
    int64_t n_ghost_cells, n_ghost_pix;
} psm_geometry;

/* Device-resident intermediates that parity tests read back (psm_get_stage). */
/* On a sharded handle H means this rank's rows (row1-row0, + ext_rows for GRID) and B its local
 * blocks (GRID additionally holds the overlap rows below), except OFFSETS and MEANS which are global. */
enum psm_stage_code {
    PSM_STAGE_GRID = 0,       /* float [2][H][W]   scaled input channels 0,1 (SMC:430-444)       */
    PSM_STAGE_XINPUT = 1,     /* float [B][pc_in]  standardised PCA coordinates (SMC:512)        */
    PSM_STAGE_MLPOUT = 2,     /* float [B][pc_p]   de-standardised MLP output (SMC:533)          */
    PSM_STAGE_BLOCKS = 3,     /* float [B][C][S][S] predicted blocks before correction (SMC:551) */
    PSM_STAGE_OFFSETS = 4,    /* double [F][B]     per-block corrections BC_coor (SMC:243)       */
    PSM_STAGE_FIELD = 5,      /* float [F][H][W]   assembled field(s) (SMC:350)                  */
    PSM_STAGE_SCALARS = 6,    /* double [4]        U_max_norm, dU_max_norm, shift[0], shift[1]   */
    PSM_STAGE_MEANS = 7,      /* double [n_tasks]  masked strip means                            */
    PSM_STAGE_XU = 8          /* float [B][2][S][S] extracted blocks, channels 0,1 (SMC:464-492), planar */
};

/* ---- lifetime -------------------------------------------------------------------------- */

/* Replaces Py_Initialize + import python_module (FOAM/PythonComm_init.H:3-19).
 * Fails with PSM_ERR_CUDA when no usable sm_100 device exists: there is no CPU fallback. */
int psm_create(psm_handle** out, const psm_config* cfg);

/* Replaces the module-level artefact loading (PMP:103-118,168-170; SMC:70-87). */
int psm_load_params(psm_handle* h, const psm_params* params);

/* Replaces init_func(array, top_boundary, obst_boundary) (PMP:172-247; FOAM/PythonComm_init.H:94)
 * for callers that already hold the Qhull tables and raster (the Python shim). */
int psm_init_with_tables(psm_handle* h, const psm_tables* tables);

/* Same from a flat binary file written by `python -m psm_b200.tables_file` (or psm_save_tables), so that a
 * C/C++ caller needs neither SciPy nor an interpreter at run time (INTEGRATION.md route B). */
int psm_init_from_file(psm_handle* h, const char* path);
int psm_load_params_file(psm_handle* h, const char* path);
int psm_save_tables(const psm_tables* tables, const char* path);
int psm_save_params(const psm_params* params, int32_t shape, const char* path);

/* init_func(array, top_boundary, obst_boundary) (PMP:172-247; SMC:89-180) for a C / C++ caller, without an interpreter.
 * The library computes the rounded bounding box and the uniform grid (SMC:102-106, UTL:111-125), the flow mask and the distance
 * field on the GPU (SMC:117-143), the index raster (SMC:161-178) and -- opt-in -- the grid -> cell tables in closed form.  The
 * cells -> grid Delaunay tables stay with Qhull (the library the reference calls through SciPy): they are handed in as `vert` /
 * `weights`, or come out of the table cache: with `cache_dir` set, the finished tables are stored under a hash of the mesh
 * (psm_mesh_hash) and the next psm_init_mesh of the same mesh is one file read -- no Delaunay, no Python. */
typedef struct psm_mesh {
    int64_t n_cells;
    const double* cells_xy;       /* cell centres: x at [i * xy_stride], y at [i * xy_stride + 1]; the solver's double[n][5] rows
                                     (FOAM/PythonComm_init.H:53-75) are passed as (rows + 2, 5)                                  */
    int32_t xy_stride;
    int32_t back_closed_form;     /* 1: grid -> cell tables in closed form when vert_back is NULL (psm_b200/tables.py); 0: none  */
    const double* top;  int64_t n_top;     /* "top" patch points  [n_top][2]   (FOAM/PythonComm_init.H:33-52)                   */
    const double* obst; int64_t n_obst;    /* "obstacle" patch points [n_obst][2]                                               */
    const double* probe;          /* [n_cells] field whose interpolation decides pixel validity (SMC:165: p; PMP:230: Ux), or NULL */
    const int32_t* vert;          /* [H*W][3] Qhull simplex vertices (UTL:40), or NULL when the cache holds this mesh           */
    const double*  weights;       /* [H*W][3]                                                                                    */
    const int32_t* vert_back;     /* [n_cells][3] (PMP:211) or NULL                                                              */
    const double*  weights_back;
    const char* cache_dir;        /* directory of the table cache, or NULL                                                       */
} psm_mesh;
int psm_init_mesh(psm_handle* h, const psm_mesh* mesh);

/* Host-only helpers of psm_init_mesh (no GPU needed; the CPU test-suite checks them against the NumPy shim). */
int psm_mesh_hash(int32_t variant, double delta, const double* cells_xy, int32_t xy_stride, int64_t n_cells, const double* top,
                  int64_t n_top, const double* obst, int64_t n_obst, const double* probe, char out[17]);
int psm_mesh_grid(int32_t variant, double delta, const double* cells_xy, int32_t xy_stride, int64_t n_cells, double bbox[4],
                  int32_t* grid_h, int32_t* grid_w);
int psm_back_tables_closed_form(const double* cells_xy, int32_t xy_stride, int64_t n_cells, const double* X0_row, int32_t W,
                                const double* Y0_col, int32_t H, int32_t* vert_back, double* weights_back);

/* ---- multi-GPU: one process per GPU, block rows over ranks ------------- */

#define PSM_UNIQUE_ID_BYTES 128
/* Rank 0 creates the NCCL unique id; the caller ships the bytes to the other ranks with its own
 * transport (MPI_Bcast / Pstream / torch.distributed) -- replaces MPI.COMM_WORLD of PMP:14-17. */
int psm_comm_get_unique_id(void* id_out /* PSM_UNIQUE_ID_BYTES */);
/* Collective over all ranks: creates the communicator the handle's step uses. */
int psm_comm_init(psm_handle* h, const void* unique_id, int32_t rank, int32_t world);
/* Collective: this rank's share of the mesh.  After it, psm_predict / psm_predict_device are
 * collective calls: cells = this rank's n_owned rows, p_out = its n_owned pressures. */
int psm_init_sharded(psm_handle* h, const psm_shard* shard);

/* The block-row partitioner for C / C++ callers (host only, no GPU): this rank's psm_shard from the GLOBAL tables -- the C++ twin of
 * psm_b200/shard.py `partition` (halo in cell space), array for array.  y_min: rounded lower edge of the cell bounding box (bbox[2] of
 * psm_mesh_grid).  The object owns the arrays psm_shard_view points into; free it after psm_init_sharded. */
typedef struct psm_shard_owned psm_shard_owned;
int psm_shard_build(const psm_tables* tables, const double* cells_xy, int32_t xy_stride, double y_min, double delta, int32_t variant,
                    int32_t shape, int32_t overlap, double near_wall_sdf, int32_t rank, int32_t world, psm_shard_owned** out);
const psm_shard* psm_shard_view(const psm_shard_owned* s);
/* owned_ids int64[n_owned]: global ids of the cells this rank passes to psm_predict (ascending); cell_rank int32[n_cells]: owner of
 * every global cell (input of the cell routing). */
int psm_shard_cells(const psm_shard_owned* s, const int64_t** owned_ids, const int32_t** cell_rank);
int psm_shard_free(psm_shard_owned* s);

/* Cell routing: the solver's own domain decomposition need not be the block-row partition.  The reference takes ANY per-rank cell
 * sets (scotch, system/decomposeParDict) by gathering every rank's rows to rank 0 and scattering the pressures back (PMP:179-185, 258,
 * 501-511); here a rank describes once where each of its cells belongs and psm_predict_routed moves the rows to the owning GPU
 * (grouped ncclSend / ncclRecv) and the pressures back.  dest_rank / dest_index come from the partitioner (psm_b200.shard.route):
 * the rank whose block rows contain the cell, and the cell's position among that rank's owned cells. */
typedef struct psm_route {
    int64_t n_local;              /* cells this solver rank holds: any subset of the mesh, any order; every cell on exactly one rank */
    const int32_t* dest_rank;     /* [n_local]                                                                                   */
    const int32_t* dest_index;    /* [n_local] row of the cell in the owner's psm_predict input (0 .. its n_owned - 1)             */
} psm_route;
/* Collective over the ranks of psm_comm_init, after psm_init_sharded (also valid on a single-GPU handle: a pure permutation). */
int psm_route_init(psm_handle* h, const psm_route* route);
/* psm_predict_fields on the rank's OWN cells (host arrays, n_local rows, the order of psm_route); collective. */
int psm_predict_routed(psm_handle* h, const double* U, int32_t u_stride, const double* dU, const double* p, int64_t n_local,
                       double* out);

/* Idempotent; NULL is accepted. */
int psm_destroy(psm_handle* h);

/* ---- per time step --------------------------------------------------------------------- */

/* Replaces py_func(array_in) (PMP:249-517; FOAM/PythonComm.H:24-35).
 * cells : host, row-major double[n_cells][input_cols]
 * p_out : host, double[n_cells] (deltaU_to_deltaP) or double[n_cells][2] (U_to_gradP)
 * Both buffers stay owned by the caller.  Returns PSM_OK or PSM_SKIPPED. */
int psm_predict(psm_handle* h, const double* cells, int64_t n_cells, double* p_out);

/* Same with DEVICE pointers (caller already holds the fields on this GPU); asynchronous on
 * the handle's stream unless `sync` is non-zero.  The status of an asynchronous call is
 * reported by the next synchronous call or psm_synchronize. */
int psm_predict_device(psm_handle* h, const double* d_cells, int64_t n_cells, double* d_p_out, int32_t sync);

/* The same step on the solver's NATIVE field storage -- no row packing (the forAll fill loop of FOAM/PythonComm.H:2-9 and its
 * double[nCells][5] buffer, FOAM/PythonComm_init.H:53, disappear; Cx, Cy are static and were consumed by init):
 *   U   : host double[n_cells][u_stride], u_stride = 3 for OpenFOAM's `vector` (U.primitiveField().cdata()) or 2; only
 *         components 0 and 1 are read.
 *   dU  : deltaU_to_deltaP only, same layout, or NULL: the handle keeps U(t-1) resident and forms dU on the device (first
 *         call returns PSM_SKIPPED), exactly as with 5-column rows.
 *   p   : host double[n_cells] previous pressure (p.primitiveField().cdata()), or NULL: the output is then the raw prediction
 *         (delta_p; 0 where the reference keeps p_prev, PMP:492-496) and the solver adds it to its own field (SMC:644-645).
 *   out : host double[n_cells] (deltaU_to_deltaP / thesis) or double[n_cells][2] (U_to_gradP).
 * 24 B (U) + 8 B (p) per cell go up instead of 40; the copy of p overlaps the kernels (only the last one reads it). */
int psm_predict_fields(psm_handle* h, const double* U, int32_t u_stride, const double* dU, const double* p, int64_t n_cells,
                       double* out);
/* Device-pointer form of psm_predict_fields (p is read in place, nothing is copied); asynchronous unless `sync`. */
int psm_predict_fields_device(psm_handle* h, const double* d_U, int32_t u_stride, const double* d_dU, const double* d_p,
                              int64_t n_cells, double* d_out, int32_t sync);

int psm_synchronize(psm_handle* h);

/* The CUDA stream (cudaStream_t) all work of this handle is launched on, so that a caller can
 * order its own work or record its own events against it. */
int psm_get_stream(const psm_handle* h, void** stream);

/* Page-lock / unlock a caller buffer that lives for the whole run (FOAM/PythonComm_init.H:53
 * allocates input_vals once), so that psm_predict copies at full PCIe speed. */
int psm_register_host_buffer(void* ptr, int64_t bytes);
int psm_unregister_host_buffer(void* ptr);

/* ---- introspection (tests, tracing) ----------------------------------------------------- */

const char* psm_last_error(const psm_handle* h);
int psm_get_geometry(const psm_handle* h, psm_geometry* out);

/* Block plan exactly as the reference loops produce it (SMC:464-479 / GRAD:486-500):
 * origins int32[B][2] = (y0, x0), indices_list int32[B][2] = (idx_i, idx_j); either may be NULL. */
int psm_get_plan(const psm_handle* h, int32_t* origins, int32_t* indices_list);

/* Last-writer map of the placement step (SMC:332-348 / GRAD:345-356): int32[H][W] block id. */
int psm_get_owner_map(const psm_handle* h, int32_t* owner);

/* Device tables after the validity fold: vert int32[H*W][3], weights float[H*W][3]. */
int psm_get_forward_table(const psm_handle* h, int32_t* vert, float* weights);

/* Copy one device-resident intermediate of the LAST step to the host; n_bytes must match. */
int psm_get_stage(psm_handle* h, int32_t stage, void* out, int64_t n_bytes);

/* U_to_gradP only (single-GPU handle): the pressure field recovered from the two assembled gradient fields of the LAST step, the way
 * the reference's evaluation does it -- Evaluation.integrate_field on four quadrants around the obstacle (GRAD:371-416) and the
 * stitch of timeStep (GRAD:585-628), quirks included (see oracle/integrate.py).  p_field: host double[H][W]. */
typedef struct psm_integrate_geometry {
    double min_x, max_x, min_y, max_y;    /* bounding box of the `top` boundary points (GRAD:205; xl, yl of GRAD:585-586)    */
    double x0_min;                        /* X0.min(): x of the first grid column (GRAD:591)                                 */
    int32_t center_row;                   /* the reference hard-codes 200 (GRAD:592); must cross the obstacle                 */
    int32_t reserved;
} psm_integrate_geometry;
int psm_integrate_gradp(psm_handle* h, const psm_integrate_geometry* g, double* p_field);

/* Per-stage milliseconds of the last step (enable_timings=1).  Order: h2d, prep, gather,
 * extract, pca_project, mlp, pca_inverse, strip_means, offsets, place, back_gather, d2h.
 * Replaces the hand-rolled time.time() pairs of PMP:262-499. */
#define PSM_N_TIMINGS 12
int psm_get_timings(psm_handle* h, float* ms, int32_t n);
/* Switch the per-stage events on/off at run time (they cost a few microseconds per step). */
int psm_set_timings(psm_handle* h, int32_t on);

/* Number of kernels launched by the last psm_predict* call. */
int psm_get_launch_count(const psm_handle* h);

/* Multi-GPU wait histogram: nanoseconds the consuming kernels of THIS rank have spent waiting for their peers' pushes
 * since the last reset, per exchange phase (0 = ghost cells + maxima, awaited by the gather; 1 = strip means, awaited
 * by the fold kernel's last CTA; 2 = ghost pixels, awaited by the grid->cell gather), and the number of waits.
 * One sample per kernel launch (its first CTA).  reset != 0 zeroes the counters after reading.  Synchronises the stream.
 * The reference has no counterpart: its ranks block in MPI gather / scatter (PMP:258, 501-511). */
int psm_get_wait_ns(psm_handle* h, uint64_t ns[3], uint32_t count[3], int32_t reset);

/* ---- host-only plan compiler (no GPU needed; used by the CPU test-suite) ---------------- */

/* Compile the static block/assembly plan for a grid and return its sizes.
 * mask : uint8[H][W], 1 where the distance field is non-zero (SMC:224 `flow_bool`).
 * Outputs may be NULL.  origins/indices_list int32[B][2]; owner int32[H][W];
 * rec int32[F][B][4] = (task_a, task_b, parent_block, is_nan) of c_k = m[a] - (m[b] - c[parent]);
 * tasks int32[n_tasks][8] = (src_block, mask_block, channel, y0, y1, x0, x1, count).        */
int psm_plan_sizes(int32_t variant, int32_t grid_h, int32_t grid_w, int32_t shape, int32_t overlap,
                   const uint8_t* mask, int32_t* n_blocks, int32_t* n_fields, int32_t* n_tasks);
int psm_plan_compile(int32_t variant, int32_t grid_h, int32_t grid_w, int32_t shape, int32_t overlap,
                     const uint8_t* mask, int32_t* origins, int32_t* indices_list, int32_t* owner,
                     int32_t* rec, int32_t* tasks);

/* Runs of the two global-shift lines (SMC:350 / GRAD:358-361) through the owner map:
 * lines int32[n][8] = (field, block, y0, y1, x0, x1, coef, n_pixels), block-local rectangles; the shift of a
 * field is sum(coef * (sum(raw block values over the run) - n_pixels * c[block])) / (3 * line length).
 * `lines` may be NULL to query n_lines. */
int psm_plan_shift_lines(int32_t variant, int32_t grid_h, int32_t grid_w, int32_t shape, int32_t overlap,
                         const uint8_t* mask, int32_t* n_lines, int32_t* lines);

/* Box plan of the PCA projection's A operand when it is fetched from the grid planes (replaces the block
 * extraction SMC:464-492 / GRAD:479-516): blocks whose origins (by0, bx0) form an arithmetic progression of `stride` pixels are one TMA
 * box.  tiles = 128-row operand tiles (0: the layout does not pay and the handle keeps the extracted operand), gx / gy = blocks per
 * row / column box, row_src int32[n_blocks] = operand row of every block, segs int32[n_segs][5] = (tile, kind 0 row box / 1 column
 * box / 2 single block, first row inside the tile, x origin, y origin).  row_src / segs may be NULL. */
int psm_grid_operand_plan(int32_t n_blocks, const int32_t* by0, const int32_t* bx0, int32_t stride, int32_t* tiles, int32_t* gx,
                          int32_t* gy, int32_t* n_segs, int32_t* row_src, int32_t* segs, int32_t max_segs);

/* Ghost-cell send map of a sharded handle (replaces the gather to rank 0 of PMP:258): send_idx[send_ptr[p] ..
 * send_ptr[p+1]) are the owned cells rank p needs.  words uint32[(n_cells + 31) / 32 + 1][2] = (bitmap of the marked cells, marked
 * cells before the word); entries int32[n_entries][2] = (peer | 0x100 if the cell has a further entry | its index << 9, slot in the
 * peer's ghost region); entry k < number of marked cells belongs to the k-th marked cell.  words / entries may be NULL. */
int psm_send_map_build(int64_t n_cells, int32_t world, const int64_t* send_ptr, const int32_t* send_idx, uint32_t* words,
                       int32_t* entries, int64_t* n_entries);

/* Unit-test entry for the GEMM kernels: C[splits][M][N] (host) = A[M][K] * B[N][K]^T (host), split-K
 * partials left unreduced.  mode = psm_gemm_mode_code.  M % 128 == 0, N % 64 == 0, K % 32 == 0. */
int psm_debug_gemm(int32_t device, int32_t mode, int32_t M, int32_t N, int32_t K, const float* A, const float* B,
                   float* C, int32_t splits);

/* Unit-test entry for the fused Dense-stack kernel (NNS:8-38): out[M][dims[n]] = Dense(linear)(relu(... relu(x W0 + b0) ...)),
 * kernels[l] in the Keras layout [dims[l]][dims[l+1]].  mode = PSM_GEMM_TC_3XTF32 or PSM_GEMM_TC_TF32.
 * clusters > 0 caps the number of 8-CTA clusters (to exercise the multi-tile-per-cluster loop). */
int psm_debug_dense_stack(int32_t device, int32_t mode, int32_t M, int32_t n_layers, const int32_t* dims,
                          const float* const* kernels, const float* const* biases, const float* x, float* out,
                          int32_t clusters);

int psm_api_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PSM_B200_H */
