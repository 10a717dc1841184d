/* psm_driver.c -- the per-time-step surrogate call through the C-ABI alone (no Python, no PyTorch).
 *
 * What a solver does in place of the embedded-interpreter calls of the reference
 * (Thesis_Work/Chapter5/parallelized/DLPoissonSolver/PythonComm_init.H:3-94 and PythonComm.H:2-36):
 *
 *   psm_driver <params.bin> <tables.bin> <cells.bin> <n_cells> <input_cols> <variant 0|1> <steps> <p_out.bin>
 *
 * cells.bin : double[n_cells][input_cols] row-major, exactly the buffer of PythonComm_init.H:53
 * p_out.bin : double[n_cells] (variant 0) or double[n_cells][2] (variant 1) of the LAST step
 *
 * With PSM_DRIVER_CACHE=<dir> PSM_DRIVER_TOP=<top.bin> PSM_DRIVER_OBST=<obst.bin> (double[n][2] each) the mesh is initialised
 * like init_func -- psm_init_mesh from the raw arrays (cell centres = columns 2, 3 of the rows), the Delaunay tables coming out of
 * the table cache -- and <tables.bin> is ignored; with PSM_DRIVER_FIELDS=1 every step goes through psm_predict_fields on
 * U double[n][3] + p double[n] (the solver's native storage) instead of the packed rows.
 * Build     : gcc -O2 -I../../include psm_driver.c -L<dir of libpsm_b200.so> -lpsm_b200 -o psm_driver
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "psm_b200.h"

static double now_ms(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static double* read_points(const char* path, long long* n) {      /* double[n][2] */
    if (!path) return NULL;
    FILE* f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    const long bytes = ftell(f);
    fseek(f, 0, SEEK_SET);
    double* a = (double*)malloc((size_t)bytes);
    if (!a || fread(a, 1, (size_t)bytes, f) != (size_t)bytes) { fclose(f); free(a); return NULL; }
    fclose(f);
    *n = bytes / 16;
    return a;
}

#define CHECK(call)                                                                     \
    do {                                                                                \
        int rc_ = (call);                                                               \
        if (rc_ < 0) {                                                                  \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, psm_last_error(h));           \
            return 1;                                                                   \
        }                                                                               \
    } while (0)

int main(int argc, char** argv) {
    if (argc != 9) {
        fprintf(stderr, "usage: %s params.bin tables.bin cells.bin n_cells input_cols variant steps p_out.bin\n", argv[0]);
        return 2;
    }
    const long long n = atoll(argv[4]);
    const int ncol = atoi(argv[5]), variant = atoi(argv[6]), steps = atoi(argv[7]);
    const int nf = variant == PSM_U_TO_GRADP ? 2 : 1;
    psm_handle* h = NULL;

    psm_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.variant = variant;
    cfg.device = 0;
    cfg.delta = 5e-3;                                   /* EP:89 */
    cfg.shape = 128;                                    /* EP:91 */
    cfg.overlap = variant == PSM_U_TO_GRADP ? 96 : 32;  /* GRAD:708 avance / EP:90 overlap ratio 0.25 */
    cfg.input_cols = ncol;
    cfg.additive = 1;                                   /* SMC:644-645 */
    cfg.ref_bc = 0.0;                                   /* SMC:570 */
    cfg.skip_threshold = 1e-4;                          /* SMC:410-411 */
    cfg.near_wall_sdf = 0.0;
    cfg.gemm_mode = PSM_GEMM_TC_3XTF32;
    CHECK(psm_create(&h, &cfg));                        /* replaces Py_Initialize + import   (init.H:3-19)  */
    CHECK(psm_load_params_file(h, argv[1]));            /* replaces the module-level loading (PMP:103-170)  */

    double* cells = (double*)malloc(sizeof(double) * n * ncol);
    double* p_out = (double*)malloc(sizeof(double) * n * nf);
    FILE* f = fopen(argv[3], "rb");
    if (!cells || !p_out || !f || fread(cells, sizeof(double), (size_t)(n * ncol), f) != (size_t)(n * ncol)) {
        fprintf(stderr, "cannot read %s\n", argv[3]);
        return 1;
    }
    fclose(f);
    const char* cache = getenv("PSM_DRIVER_CACHE");
    if (cache) {                                        /* init_func(array, top, obst) (PMP:172-247) from the raw arrays */
        long long n_top = 0, n_obst = 0;
        double* top = read_points(getenv("PSM_DRIVER_TOP"), &n_top);
        double* obst = read_points(getenv("PSM_DRIVER_OBST"), &n_obst);
        double* probe = (double*)malloc(sizeof(double) * n);
        if (!top || !obst || !probe) { fprintf(stderr, "cannot read the boundary points\n"); return 1; }
        for (long long i = 0; i < n; ++i) probe[i] = cells[i * ncol + (variant == PSM_DELTAU_TO_DELTAP ? 4 : 0)];   /* SMC:165 p ; PMP:230 ux */
        psm_mesh m;
        memset(&m, 0, sizeof m);
        m.n_cells = n; m.cells_xy = cells + 2; m.xy_stride = ncol; m.back_closed_form = 1;
        m.top = top; m.n_top = n_top; m.obst = obst; m.n_obst = n_obst; m.probe = probe; m.cache_dir = cache;
        CHECK(psm_init_mesh(h, &m));                    /* vert / weights NULL: the table cache must hold this mesh */
        free(top); free(obst); free(probe);
    } else {
        CHECK(psm_init_from_file(h, argv[2]));          /* replaces init_func                (PMP:172-247)  */
    }
    const int use_fields = getenv("PSM_DRIVER_FIELDS") != NULL;
    double *U3 = NULL, *dU3 = NULL, *pp = NULL;
    if (use_fields) {                                   /* the solver's native storage: vectorField U, scalarField p */
        U3 = (double*)calloc((size_t)n * 3, sizeof(double)); pp = (double*)malloc(sizeof(double) * n);
        if (ncol == 7) dU3 = (double*)calloc((size_t)n * 3, sizeof(double));
        for (long long i = 0; i < n; ++i) {
            U3[3 * i] = cells[i * ncol]; U3[3 * i + 1] = cells[i * ncol + 1]; pp[i] = cells[i * ncol + 4];
            if (dU3) { dU3[3 * i] = cells[i * ncol + 5]; dU3[3 * i + 1] = cells[i * ncol + 6]; }
        }
    }
    /* the solver allocates both buffers once (init.H:53): page-lock them so that the copies run at PCIe speed */
    CHECK(psm_register_host_buffer(cells, (int64_t)sizeof(double) * n * ncol));
    CHECK(psm_register_host_buffer(p_out, (int64_t)sizeof(double) * n * nf));

    int last = 0;
    double t_sum = 0.0;
    for (int s = 0; s < steps; ++s) {
        const double t0 = now_ms();
        last = use_fields ? psm_predict_fields(h, U3, 3, dU3, pp, n, p_out)
                          : psm_predict(h, cells, n, p_out);   /* replaces PyObject_CallObject(py_func, ...) (PythonComm.H:24-35) */
        if (last < 0) { fprintf(stderr, "psm_predict -> %d: %s\n", last, psm_last_error(h)); return 1; }
        if (s > 0) t_sum += now_ms() - t0;              /* like DLPoissonFoam.C:106-111, first (capturing) call excluded */
    }
    psm_geometry g;
    CHECK(psm_get_geometry(h, &g));
    printf("psm_driver: %lld cells, grid %d x %d, %d blocks, status %d, %.3f ms/step over %d steps, %d kernel launches/step\n",
           n, g.grid_h, g.grid_w, g.n_blocks, last, steps > 1 ? t_sum / (steps - 1) : 0.0, steps, psm_get_launch_count(h));

    f = fopen(argv[8], "wb");
    if (!f || fwrite(p_out, sizeof(double), (size_t)(n * nf), f) != (size_t)(n * nf)) { fprintf(stderr, "cannot write %s\n", argv[8]); return 1; }
    fclose(f);
    psm_unregister_host_buffer(cells);
    psm_unregister_host_buffer(p_out);
    psm_destroy(h);
    free(cells);
    free(p_out);
    return 0;
}
