/* psm_driver.c -- the per-time-step surrogate call through the C-ABI alone (no Python, no PyTorch).
 *
 * What a solver does in place of the embedded-interpreter calls of the reference
 * (Thesis_Work/Chapter5/parallelized/DLPoissonSolver/PythonComm_init.H:3-94 and PythonComm.H:2-36):
 *
 *   psm_driver <params.bin> <tables.bin> <cells.bin> <n_cells> <input_cols> <variant 0|1> <steps> <p_out.bin>
 *
 * cells.bin : double[n_cells][input_cols] row-major, exactly the buffer of PythonComm_init.H:53
 * p_out.bin : double[n_cells] (variant 0) or double[n_cells][2] (variant 1) of the LAST step
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
    CHECK(psm_init_from_file(h, argv[2]));              /* replaces init_func                (PMP:172-247)  */

    double* cells = (double*)malloc(sizeof(double) * n * ncol);
    double* p_out = (double*)malloc(sizeof(double) * n * nf);
    FILE* f = fopen(argv[3], "rb");
    if (!cells || !p_out || !f || fread(cells, sizeof(double), (size_t)(n * ncol), f) != (size_t)(n * ncol)) {
        fprintf(stderr, "cannot read %s\n", argv[3]);
        return 1;
    }
    fclose(f);
    /* the solver allocates both buffers once (init.H:53): page-lock them so that the copies run at PCIe speed */
    CHECK(psm_register_host_buffer(cells, (int64_t)sizeof(double) * n * ncol));
    CHECK(psm_register_host_buffer(p_out, (int64_t)sizeof(double) * n * nf));

    int last = 0;
    double t_sum = 0.0;
    for (int s = 0; s < steps; ++s) {
        const double t0 = now_ms();
        last = psm_predict(h, cells, n, p_out);        /* replaces PyObject_CallObject(py_func, ...) (PythonComm.H:24-35) */
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
