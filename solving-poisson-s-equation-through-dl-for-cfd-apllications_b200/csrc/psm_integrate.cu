// U_to_gradP pressure recovery on the GPU: Evaluation.integrate_field (GRAD:371-416) on four quadrants around the obstacle and
// the stitch of timeStep (GRAD:585-628), restated from the closed form of its loops (oracle/integrate.py is the CPU restatement,
// pinned to the reference's own output):
//   SdPx[i, :]  = cumsum(a'_i) * dx, where a'_i is row i of dP/dx with the few entries the reference's "reset" overwrites
//                 (GRAD:389-395: index array = int(sdfunct[i_local, :]), last write wins) -- a static fix-up list per row;
//   SdPy[:, j0] = cumsum over the quadrant's rows of dP/dy in the quadrant's anchor column, * dy (only the anchor column is used);
//   P[i, j]     = SdPy[i, j0] - SdPy[i0, j0] + SdPx[i, j] - SdPx[i, j0]            (GRAD:405-414)
//   stitch      : left quadrants shifted by mean(P_left[:, last][mask_l] - P_right[:, first][mask_r])   (GRAD:606, 620)
// FP64 throughout, like the reference.  One warp per (row, side) for the row scans; the anchor columns and the two stitch
// means are a few thousand scalars (one CTA).
#include <cuda_runtime.h>

#include <cstdint>

#include "psm_kernels.cuh"

namespace psm {

namespace {
__device__ __forceinline__ double warp_scan_incl(double v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}
}  // namespace

// side 0: left quadrants, columns [0, cx); side 1: right quadrants, columns [cx-1, W).  Output sdpx[side][row][W] (block-local
// scan written at the global column).  fix: per LOCAL row (row inside its quadrant) up to 4 (position, previous position) pairs.
__global__ void __launch_bounds__(256) integrate_rows_kernel(IntegrateArgs a) {
    const int lane = threadIdx.x & 31;
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wid >= 2ll * a.H) return;
    const int side = (int)(wid / a.H), y = (int)(wid - (long long)side * a.H);
    const int c0 = side ? a.cx - 1 : 0, c1 = side ? a.W : a.cx;
    const int wq = c1 - c0;
    const int il = (y < a.cy) ? y : y - a.cy;                         // GRAD:392 uses the row index inside the block
    const float* __restrict__ src = a.dpdx + (long long)y * a.W + c0;
    double* __restrict__ dst = a.sdpx + ((long long)side * a.H + y) * a.W + c0;
    const IntegrateFix fx = a.fix[il];
    // values the "reset" writes: a'[v] = -(ccc[v] - ccc[prev])  (prev < 0: -ccc[v]); ccc = prefix sums of the ORIGINAL row
    double newv[4] = {0.0, 0.0, 0.0, 0.0};
    for (int k = 0; k < fx.n; ++k) {
        double cv = 0.0, cp = 0.0;
        for (int j = 0; j <= fx.pos[k] && j < wq; ++j) cv += (double)src[j];
        if (fx.prev[k] >= 0) for (int j = 0; j <= fx.prev[k] && j < wq; ++j) cp += (double)src[j];
        newv[k] = -(cv - cp);
    }
    double carry = 0.0;
    for (int j0 = 0; j0 < wq; j0 += 32) {
        const int j = j0 + lane;
        double v = 0.0;
        if (j < wq) {
            v = (double)src[j];
            for (int k = 0; k < fx.n; ++k) if (fx.pos[k] == j) v = newv[k];
        }
        v = warp_scan_incl(v, lane) + carry;
        if (j < wq) dst[j] = v * a.dx;
        carry = __shfl_sync(0xffffffffu, v, 31);
    }
}

// One CTA: the four anchor-column scans of dP/dy, then the two stitch means; results in a.anchor ([4][H] SdPy per quadrant at its
// rows) and a.corr[2] (upper, lower).  Quadrant q: 0 upper right, 1 upper left, 2 lower right, 3 lower left.
__device__ __forceinline__ double quad_value(const IntegrateArgs& a, int q, int y, int x) {
    const int side = (q == 0 || q == 2) ? 1 : 0;
    const bool lower = q >= 2;
    const int j0 = side ? a.W - 1 : 0;                                 // anchor column (direction_x = -1 on the right)
    const int i0 = lower ? a.H - 1 : 0;                                // anchor row    (direction_y = -1 below)
    const double* sx = a.sdpx + ((long long)side * a.H + y) * a.W;
    const double* an = a.anchor + (long long)q * a.H;
    return ((an[y] + (-an[i0])) + sx[x]) + (-sx[j0]);
}
__global__ void __launch_bounds__(128) integrate_anchor_kernel(IntegrateArgs a) {
    const int q = threadIdx.x;
    if (q < 4) {
        const int side = (q == 0 || q == 2) ? 1 : 0;
        const bool lower = q >= 2;
        const int col = side ? a.W - 1 : 0;
        const int r0 = lower ? a.cy : 0, r1 = lower ? a.H : a.cy;
        double acc = 0.0;
        double* an = a.anchor + (long long)q * a.H;
        for (int y = r0; y < r1; ++y) { acc += (double)a.dpdy[(long long)y * a.W + col]; an[y] = acc * a.dy; }
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        // (P_left[:, last][mask_l] - P_right[:, first][mask_r]).mean(): both columns compressed by their own mask, then
        // subtracted element by element (GRAD:606 / 620).  mask_l: sdfunct[rows, cx] != 0, mask_r: sdfunct[rows, cx - 1] != 0.
        const bool lower = threadIdx.x == 1;
        const int r0 = lower ? a.cy : 0, r1 = lower ? a.H : a.cy;
        const int ql = lower ? 3 : 1, qr = lower ? 2 : 0;
        int yl = r0, yr = r0, n = 0;
        double acc = 0.0;
        while (true) {
            while (yl < r1 && !a.mask[(long long)yl * a.W + a.cx]) ++yl;
            while (yr < r1 && !a.mask[(long long)yr * a.W + a.cx - 1]) ++yr;
            if (yl >= r1 || yr >= r1) break;
            acc += quad_value(a, ql, yl, a.cx - 1) - quad_value(a, qr, yr, a.cx - 1);
            ++n; ++yl; ++yr;
        }
        // unequal counts: numpy raises (shape mismatch); report it
        bool rest = false;
        for (int y = yl; y < r1; ++y) rest = rest || a.mask[(long long)y * a.W + a.cx];
        for (int y = yr; y < r1; ++y) rest = rest || a.mask[(long long)y * a.W + a.cx - 1];
        a.corr[threadIdx.x] = n > 0 ? acc / (double)n : __longlong_as_double(0x7ff8000000000000ll);
        if (rest) a.status[0] = 1;
    }
}

__global__ void __launch_bounds__(256) integrate_combine_kernel(IntegrateArgs a) {
    const long long total = (long long)a.H * a.W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / a.W), x = (int)(i - (long long)y * a.W);
        const bool lower = y >= a.cy;
        const bool left = x < a.cx;                                    // the left quadrant is written last: it owns column cx-1
        const int q = lower ? (left ? 3 : 2) : (left ? 1 : 0);
        double v = quad_value(a, q, y, x);
        if (left) v -= a.corr[lower ? 1 : 0];
        a.out[i] = v;
    }
}

void launch_integrate(const IntegrateArgs& a, cudaStream_t s) {
    const long long warps = 2ll * a.H;
    integrate_rows_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, s>>>(a);
    integrate_anchor_kernel<<<1, 128, 0, s>>>(a);
    long long want = ((long long)a.H * a.W + 255) / 256;
    integrate_combine_kernel<<<(unsigned)(want < 148 * 8 ? want : 148 * 8), 256, 0, s>>>(a);
}

}  // namespace psm
