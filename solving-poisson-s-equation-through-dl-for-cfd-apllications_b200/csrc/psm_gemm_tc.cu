// tcgen05 GEMM for the dense contractions of the surrogate (PCA projection, Dense stack, PCA
// inverse):  C[M,N] = A[M,K] * B[N,K]^T, both operands K-major FP32 in global memory.
//
//   * TMA (cp.async.bulk.tensor, SWIZZLE_128B) stages 128 x 32 (A) and BN x 32 (B) FP32 tiles;
//   * tcgen05.mma kind::tf32, M=128, N=BN, K=8 per instruction, FP32 accumulators in TMEM;
//   * "3xTF32": four converter warps split every staged tile in place into hi = tf32(x) and
//     lo = x - hi, and the issuing thread accumulates  hi*hi + hi*lo + lo*hi  -- FP32-class accuracy
//     (the reference is float64 around a float32 Dense stack) at no extra HBM traffic; the kernels
//     are HBM-bound, so the 3x tensor work is free;
//   * the same four warps then run the epilogue: tcgen05.ld -> bias / ReLU / affine / PCA mean
//     and re-dimensionalisation -> 128-bit stores.  Split-K partials go to a workspace.
//
// Warp roles (192 threads): 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..5 = converters
// and epilogue (TMEM lane quadrant = warp_idx % 4).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#include "psm_kernels.cuh"

namespace psm {

namespace {

constexpr int BM = 128;           // UMMA_M
constexpr int BK = 32;            // 32 fp32 = 128 B = one swizzle row
constexpr int UMMA_K = 8;         // tf32
constexpr int kThreads = 192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4,
// LBO = 1 (unused for swizzled K-major), SBO = 1024 B (8 rows x 128 B), version 1, layout type 2.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// cute::UMMA::InstrDescriptor for kind::tf32: c_format F32 (1) @4, a/b format TF32 (2) @7/@10,
// K-major A and B, n_dim = N>>3 @17, m_dim = M>>4 @24.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// dynamic shared memory: STAGES x [A_hi | A_lo | B_hi | B_lo] after manual 1024 B alignment, then barriers
template <int BN> struct Cfg {
    static constexpr int STAGES = (BN == 128) ? 3 : 4;
    static constexpr int STAGE_BYTES = 2 * (BM + BN) * BK * 4;
    static constexpr int SMEM_TOTAL = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

}  // namespace

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcGemmArgs g) {
    constexpr int STAGES = Cfg<BN>::STAGES;
    constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4, STAGE = 2 * (A_BYTES + B_BYTES);
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    // stage s: [A_hi | A_lo | B_hi | B_lo]
    const uint32_t bars = base + STAGES * STAGE;                 // full[S], conv[S], empty[S], accum, tmem slot
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto conv_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    auto empty_bar = [&](int s) { return bars + 8u * (2 * STAGES + s); };
    const uint32_t accum_bar = bars + 8u * (3 * STAGES);
    const uint32_t tmem_slot = accum_bar + 8u;
    uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));   // generic pointer to the aligned base

    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kb_total = g.K / BK;
    const int kb_per = (kb_total + g.splits - 1) / g.splits;
    const int kb0 = blockIdx.z * kb_per;
    const int kb1 = min(kb_total, kb0 + kb_per);
    const int nkb = kb1 - kb0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(conv_bar(s), 128); mbar_init(empty_bar(s), 1); }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // TMEM allocation: BN fp32 accumulator columns (power of two >= 32)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - base));
    // The B operand of every caller is STATIC (PCA matrix / Dense kernel): the producer requests the B tiles of the first
    // ring pass before waiting for the previous kernel, so their HBM / L2 latency overlaps its tail.
    const int n_pre = g.b_static ? min(nkb, STAGES) : 0;
    if (threadIdx.x == 0) {
        for (int it = 0; it < n_pre; ++it) {
            mbar_expect_tx(full_bar(it), A_BYTES + B_BYTES);
            tma_load_2d(base + it * STAGE + 2 * A_BYTES, &tmB, (kb0 + it) * BK, n0, full_bar(it));
        }
    }
    pdl_wait();                                                  // barriers + TMEM are set up while the previous kernel drains

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            for (int it = 0; it < nkb; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                const uint32_t st = base + s * STAGE;
                if (it >= n_pre) {
                    mbar_wait(empty_bar(s), ph ^ 1);
                    mbar_expect_tx(full_bar(s), A_BYTES + B_BYTES);
                    tma_load_2d(st + 2 * A_BYTES, &tmB, (kb0 + it) * BK, n0, full_bar(s));
                }
                tma_load_2d(st, &tmA, (kb0 + it) * BK, m0, full_bar(s));
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(BM, BN);
            for (int it = 0; it < nkb; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(g.three_pass ? conv_bar(s) : full_bar(s), ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = base + s * STAGE;
                const uint32_t a_hi = st, a_lo = st + A_BYTES, b_hi = st + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    const uint32_t ko = k * UMMA_K * 4;        // byte offset inside the 128 B swizzle row
                    umma_tf32(tmem_base, make_smem_desc(a_hi + ko), make_smem_desc(b_hi + ko), idesc, (it | k) != 0);
                    if (g.three_pass) {
                        umma_tf32(tmem_base, make_smem_desc(a_hi + ko), make_smem_desc(b_lo + ko), idesc, 1);
                        umma_tf32(tmem_base, make_smem_desc(a_lo + ko), make_smem_desc(b_hi + ko), idesc, 1);
                    }
                }
                umma_commit(empty_bar(s));                     // frees the stage once these MMAs retire
            }
            umma_commit(accum_bar);                            // accumulator complete
        }
    } else {
        // ===================== converters (3xTF32 split), then epilogue =====================
        const int t = threadIdx.x - 64;                        // 0..127
        const bool keep_hi = g.three_pass != 2;                // 2: the tensor core truncates the raw FP32 tile to TF32 itself
        if (g.three_pass) {
            for (int it = 0; it < nkb; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(full_bar(s), ph);
                uint8_t* st = gen_base + s * STAGE;
                float4* a_hi = reinterpret_cast<float4*>(st);
                float4* a_lo = reinterpret_cast<float4*>(st + A_BYTES);
                float4* b_hi = reinterpret_cast<float4*>(st + 2 * A_BYTES);
                float4* b_lo = reinterpret_cast<float4*>(st + 2 * A_BYTES + B_BYTES);
                auto split4 = [](float4 v, float4& hi, float4& lo) {
                    hi.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); lo.x = v.x - hi.x;
                    hi.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); lo.y = v.y - hi.y;
                    hi.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); lo.z = v.z - hi.z;
                    hi.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); lo.w = v.w - hi.w;
                };
#pragma unroll 4
                for (int i = t; i < A_BYTES / 16; i += 128) {
                    float4 hi, lo;
                    split4(a_hi[i], hi, lo);
                    if (keep_hi) a_hi[i] = hi;
                    a_lo[i] = lo;
                }
#pragma unroll 4
                for (int i = t; i < B_BYTES / 16; i += 128) {
                    float4 hi, lo;
                    split4(b_hi[i], hi, lo);
                    if (keep_hi) b_hi[i] = hi;
                    b_lo[i] = lo;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> async proxy (UMMA)
                mbar_arrive(conv_bar(s));
            }
        }
        // ---- epilogue: thread <-> accumulator row (TMEM lane), 16 columns per tcgen05.ld ----
        if (nkb > 0) {
            mbar_wait(accum_bar, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const int q = warp & 3;                                // TMEM lane quadrant this warp may access
        const int row = q * 32 + lane;
        const int m = m0 + row;
        float* Cp = g.C + (g.epi == EPI_PARTIAL ? (size_t)blockIdx.z * g.M * g.ldc : 0) + (size_t)m * g.ldc + n0;
        const float o_scale = (g.epi == EPI_PCA_INV) ? g.sc->out_scale : 1.f;
        for (int c = 0; c < BN; c += 16) {
            float v[16];
            if (nkb > 0) tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
            else {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = 0.f;
            }
            const int n = n0 + c;
            if (g.epi == EPI_BIAS_RELU) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i] + __ldg(g.v0 + n + i), 0.f);
            } else if (g.epi == EPI_BIAS_AFFINE) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = (v[i] + __ldg(g.v0 + n + i)) * __ldg(g.v1 + n + i) + __ldg(g.v2 + n + i);
            } else if (g.epi == EPI_PCA_INV) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = (v[i] + __ldg(g.v0 + n + i)) * o_scale;
            }
            if (m < g.M) {
#pragma unroll
                for (int i = 0; i < 16; i += 4)
                    *reinterpret_cast<float4*>(Cp + c + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// PCA projection with the A operand fetched from the grid planes (GridA in psm_kernels.cuh): tc_gemm_kernel<128> with another
// producer.  Split-K partials only (EPI_PARTIAL); rows of a tile the segments do not cover hold stale shared memory -- every
// accumulator row depends on its own A row only, and the reduce kernel never reads those rows.
namespace {
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(bar)
        : "memory");
}
}  // namespace

__global__ void __launch_bounds__(kThreads, 1)
tc_gemm_grid_kernel(const __grid_constant__ CUtensorMap tmRow, const __grid_constant__ CUtensorMap tmCol,
                    const __grid_constant__ CUtensorMap tmOne, const __grid_constant__ CUtensorMap tmB, TcGemmArgs g, GridA ga) {
    constexpr int BN = 128;
    constexpr int STAGES = Cfg<BN>::STAGES;
    constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4, STAGE = 2 * (A_BYTES + B_BYTES);
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + STAGES * STAGE;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto conv_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    auto empty_bar = [&](int s) { return bars + 8u * (2 * STAGES + s); };
    const uint32_t accum_bar = bars + 8u * (3 * STAGES);
    const uint32_t tmem_slot = accum_bar + 8u;
    uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));

    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.y, m0 = tile * BM;
    const int kb_total = g.K / BK;
    const int kb_per = (kb_total + g.splits - 1) / g.splits;
    const int kb0 = min((int)blockIdx.z * kb_per, kb_total);
    const int kb1 = min(kb_total, kb0 + kb_per);
    const int nkb = kb1 - kb0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(conv_bar(s), 128); mbar_init(empty_bar(s), 1); }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - base));
    // static per tile: its segment list and the bytes it delivers per k-block (fetched before the wait)
    const int seg0 = __ldg(ga.seg_ptr + tile), seg1 = __ldg(ga.seg_ptr + tile + 1);
    const uint32_t a_bytes = (uint32_t)__ldg(ga.a_bytes + tile);
    const int n_pre = g.b_static ? min(nkb, STAGES) : 0;
    if (threadIdx.x == 0) {
        for (int it = 0; it < n_pre; ++it) {
            mbar_expect_tx(full_bar(it), a_bytes + B_BYTES);
            tma_load_2d(base + it * STAGE + 2 * A_BYTES, &tmB, (kb0 + it) * BK, 0, full_bar(it));
        }
    }
    pdl_wait();

    if (warp == 0) {
        // the whole warp produces: lane q issues box q of the tile (one instruction moves up to 32 boxes), lane 0 also the B tile
        const int per_ch = ga.S * ga.S / BK, per_row = ga.S / BK;
        const int q0 = seg0 + lane;
        ASeg mine{};
        const CUtensorMap* mp = &tmOne;
        if (q0 < seg1) { mine = ga.segs[q0]; mp = mine.map == 0 ? &tmRow : (mine.map == 1 ? &tmCol : &tmOne); }
        for (int it = 0; it < nkb; ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            const uint32_t st = base + s * STAGE;
            if (it >= n_pre) {
                mbar_wait(empty_bar(s), ph ^ 1);
                if (lane == 0) {
                    mbar_expect_tx(full_bar(s), a_bytes + B_BYTES);
                    tma_load_2d(st + 2 * A_BYTES, &tmB, (kb0 + it) * BK, 0, full_bar(s));
                }
                __syncwarp();
            }
            const int kb = kb0 + it;
            const int ch = kb / per_ch, r = (kb % per_ch) / per_row, xo = (kb % per_row) * BK;
            if (q0 < seg1) tma_load_5d(st + (uint32_t)mine.row * (BK * 4), mp, mine.x + xo, 0, mine.y + r, 0, ch, full_bar(s));
            for (int q = q0 + 32; q < seg1; q += 32) {           // more than 32 boxes per tile: plain loop
                const ASeg sg = ga.segs[q];
                const CUtensorMap* m2 = sg.map == 0 ? &tmRow : (sg.map == 1 ? &tmCol : &tmOne);
                tma_load_5d(st + (uint32_t)sg.row * (BK * 4), m2, sg.x + xo, 0, sg.y + r, 0, ch, full_bar(s));
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(BM, BN);
            for (int it = 0; it < nkb; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(g.three_pass ? conv_bar(s) : full_bar(s), ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = base + s * STAGE;
                const uint32_t a_hi = st, a_lo = st + A_BYTES, b_hi = st + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    const uint32_t ko = k * UMMA_K * 4;
                    umma_tf32(tmem_base, make_smem_desc(a_hi + ko), make_smem_desc(b_hi + ko), idesc, (it | k) != 0);
                    if (g.three_pass) {
                        umma_tf32(tmem_base, make_smem_desc(a_hi + ko), make_smem_desc(b_lo + ko), idesc, 1);
                        umma_tf32(tmem_base, make_smem_desc(a_lo + ko), make_smem_desc(b_hi + ko), idesc, 1);
                    }
                }
                umma_commit(empty_bar(s));
            }
            umma_commit(accum_bar);
        }
    } else {
        const int t = threadIdx.x - 64;
        const bool keep_hi = g.three_pass != 2;
        if (g.three_pass) {
            for (int it = 0; it < nkb; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(full_bar(s), ph);
                uint8_t* st = gen_base + s * STAGE;
                float4* a_hi = reinterpret_cast<float4*>(st);
                float4* a_lo = reinterpret_cast<float4*>(st + A_BYTES);
                float4* b_hi = reinterpret_cast<float4*>(st + 2 * A_BYTES);
                float4* b_lo = reinterpret_cast<float4*>(st + 2 * A_BYTES + B_BYTES);
                auto split4 = [](float4 v, float4& hi, float4& lo) {
                    hi.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); lo.x = v.x - hi.x;
                    hi.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); lo.y = v.y - hi.y;
                    hi.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); lo.z = v.z - hi.z;
                    hi.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); lo.w = v.w - hi.w;
                };
#pragma unroll 4
                for (int i = t; i < A_BYTES / 16; i += 128) { float4 hi, lo; split4(a_hi[i], hi, lo); if (keep_hi) a_hi[i] = hi; a_lo[i] = lo; }
#pragma unroll 4
                for (int i = t; i < B_BYTES / 16; i += 128) { float4 hi, lo; split4(b_hi[i], hi, lo); if (keep_hi) b_hi[i] = hi; b_lo[i] = lo; }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(conv_bar(s));
            }
        }
        if (nkb > 0) {
            mbar_wait(accum_bar, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const int q = warp & 3;
        const int row = q * 32 + lane;
        float* Cp = g.C + (size_t)blockIdx.z * g.M * g.ldc + (size_t)(m0 + row) * g.ldc;
        for (int c = 0; c < BN; c += 16) {
            float v[16];
            if (nkb > 0) tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
            else {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = 0.f;
            }
#pragma unroll
            for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(Cp + c + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
    }
}

int tc_gemm_grid_prepare() {
    return cudaFuncSetAttribute(tc_gemm_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<128>::SMEM_TOTAL) == cudaSuccess ? 0 : -1;
}
void launch_tc_gemm_grid(const TcGemmGrid& t, cudaStream_t s) {
    dim3 grid(1, t.tiles, t.args.splits);
    const CUtensorMap& r = *reinterpret_cast<const CUtensorMap*>(&t.mapRow);
    const CUtensorMap& c = *reinterpret_cast<const CUtensorMap*>(&t.mapCol);
    const CUtensorMap& o = *reinterpret_cast<const CUtensorMap*>(&t.mapOne);
    const CUtensorMap& b = *reinterpret_cast<const CUtensorMap*>(&t.mapB);
    launch_k(tc_gemm_grid_kernel, grid, dim3(kThreads), Cfg<128>::SMEM_TOTAL, s, r, c, o, b, t.args, t.ga);
}

// ------------------------------------------------------------------------------------------------
// Dense layer (NNS:24-33) as ONE launch: cluster split-K with an on-chip reduction.
// The batch is only B blocks (one or a few 128-row tiles), so a layer is latency-bound: it is cut into
// 64-column tiles x KS K-slices, the KS CTAs of a tile form a thread-block cluster, every CTA runs the
// TMA -> 3xTF32 split -> tcgen05.mma pipeline over its K-slice, parks its partial accumulator in its
// own shared memory, and after a cluster barrier CTA z folds rows [z*128/KS, (z+1)*128/KS) of all KS
// partials through distributed shared memory (fixed order: deterministic), applies bias + ReLU (or
// bias + de-standardisation, SMC:533) and stores coalesced rows.  No partials in HBM, no reduce kernel.
namespace {
constexpr int D_BN = 64;
static_assert(kThreads % (D_BN / 4) == 0, "one epilogue column group per thread");
constexpr int D_STAGES = 2;
constexpr int D_A_BYTES = BM * BK * 4, D_B_BYTES = D_BN * BK * 4, D_STAGE = 2 * (D_A_BYTES + D_B_BYTES);
constexpr int D_PARK_LD = D_BN + 4;                    // floats per received row (16-byte aligned, conflict-free float4 rows)
constexpr int D_PARK = BM * D_PARK_LD * 4;             // receive buffer: [KS slices][128 / KS rows][D_PARK_LD]
constexpr int D_SMEM = D_STAGES * D_STAGE + D_PARK + 1024 + 256;

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_dsmem_v4(uint32_t local_addr, uint32_t cta, float4 v) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(cta));
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(remote), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, float (&v)[64]) {
    uint32_t r[64];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
        "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
          "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
          "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]),
          "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]),
          "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = __uint_as_float(r[i]);
}
}  // namespace

__global__ void __launch_bounds__(kThreads, 1)
dense_cluster_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmAlo, const __grid_constant__ CUtensorMap tmBlo, TcGemmArgs g) {
    constexpr int BN = D_BN, STAGES = D_STAGES, A_BYTES = D_A_BYTES, B_BYTES = D_B_BYTES, STAGE = D_STAGE;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t park = base + STAGES * STAGE;
    const uint32_t bars = park + D_PARK;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto conv_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    auto empty_bar = [&](int s) { return bars + 8u * (2 * STAGES + s); };
    const uint32_t accum_bar = bars + 8u * (3 * STAGES);
    const uint32_t tmem_slot = accum_bar + 8u;
    uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));

    pdl_launch_dependents();
    const int cta_lin = (blockIdx.y * gridDim.x + blockIdx.x) * gridDim.z + blockIdx.z;
    auto stamp = [&](int i) {
        if (g.trace && threadIdx.x == 64 && cta_lin < 64) {
            unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            g.trace[cta_lin * 8 + i] = t;
        }
    };
    stamp(0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KS = gridDim.z;                                    // cluster = (1,1,KS): rank in cluster = blockIdx.z
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kb_total = g.K / BK;
    const int kb_per = (kb_total + KS - 1) / KS;
    const int kb0 = blockIdx.z * kb_per;
    const int kb1 = min(kb_total, kb0 + kb_per);
    const int nkb = max(kb1 - kb0, 0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(conv_bar(s), 128); mbar_init(empty_bar(s), 1); }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - base));
    cluster_sync_all();                                          // every CTA of the cluster runs: its shared memory may be written remotely
    // the Dense kernels are static: their tiles of the first ring pass are requested before waiting for the previous layer
    const int n_pre = g.b_static ? min(nkb, STAGES) : 0;
    const bool pre = g.presplit != 0;                            // lo halves come from global memory: no converter pass
    const uint32_t tx = pre ? 2 * (A_BYTES + B_BYTES) : A_BYTES + B_BYTES;
    if (threadIdx.x == 0) {
        for (int it = 0; it < n_pre; ++it) {
            mbar_expect_tx(full_bar(it), tx);
            tma_load_2d(base + it * STAGE + 2 * A_BYTES, &tmB, (kb0 + it) * BK, n0, full_bar(it));
            if (pre) tma_load_2d(base + it * STAGE + 2 * A_BYTES + B_BYTES, &tmBlo, (kb0 + it) * BK, n0, full_bar(it));
        }
    }
    // the epilogue vectors of this thread's output columns are static too (every element this thread folds has the same c4):
    // fetched before the wait instead of behind the cluster barrier
    const int my_c4 = threadIdx.x % (BN / 4);
    const float4 e_b = __ldg(reinterpret_cast<const float4*>(g.v0 + n0 + my_c4 * 4));
    float4 e_sc = make_float4(1.f, 1.f, 1.f, 1.f), e_sh = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g.epi != EPI_BIAS_RELU) {
        e_sc = __ldg(reinterpret_cast<const float4*>(g.v1 + n0 + my_c4 * 4));
        e_sh = __ldg(reinterpret_cast<const float4*>(g.v2 + n0 + my_c4 * 4));
    }
    pdl_wait();                                                  // barriers + TMEM are set up while the previous kernel drains
    stamp(1);

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < nkb; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                const uint32_t st = base + s * STAGE;
                if (it >= n_pre) {
                    mbar_wait(empty_bar(s), ph ^ 1);
                    mbar_expect_tx(full_bar(s), tx);
                    tma_load_2d(st + 2 * A_BYTES, &tmB, (kb0 + it) * BK, n0, full_bar(s));
                    if (pre) tma_load_2d(st + 2 * A_BYTES + B_BYTES, &tmBlo, (kb0 + it) * BK, n0, full_bar(s));
                }
                tma_load_2d(st, &tmA, (kb0 + it) * BK, m0, full_bar(s));
                if (pre) tma_load_2d(st + A_BYTES, &tmAlo, (kb0 + it) * BK, m0, full_bar(s));
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(BM, BN);
            for (int it = 0; it < nkb; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait((g.three_pass && !pre) ? conv_bar(s) : full_bar(s), ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = base + s * STAGE;
                const uint32_t a_hi = st, a_lo = st + A_BYTES, b_hi = st + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    const uint32_t ko = k * UMMA_K * 4;
                    umma_tf32(tmem_base, make_smem_desc(a_hi + ko), make_smem_desc(b_hi + ko), idesc, (it | k) != 0);
                    if (g.three_pass) {
                        umma_tf32(tmem_base, make_smem_desc(a_hi + ko), make_smem_desc(b_lo + ko), idesc, 1);
                        umma_tf32(tmem_base, make_smem_desc(a_lo + ko), make_smem_desc(b_hi + ko), idesc, 1);
                    }
                }
                umma_commit(empty_bar(s));
            }
            umma_commit(accum_bar);
        }
        __syncwarp();
    } else {
        const int t = threadIdx.x - 64;
        const bool keep_hi = g.three_pass != 2;
        if (g.three_pass && !pre) {
            for (int it = 0; it < nkb; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(full_bar(s), ph);
                uint8_t* st = gen_base + s * STAGE;
                float4* a_hi = reinterpret_cast<float4*>(st);
                float4* a_lo = reinterpret_cast<float4*>(st + A_BYTES);
                float4* b_hi = reinterpret_cast<float4*>(st + 2 * A_BYTES);
                float4* b_lo = reinterpret_cast<float4*>(st + 2 * A_BYTES + B_BYTES);
                auto split4 = [](float4 v, float4& hi, float4& lo) {
                    hi.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); lo.x = v.x - hi.x;
                    hi.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); lo.y = v.y - hi.y;
                    hi.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); lo.z = v.z - hi.z;
                    hi.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); lo.w = v.w - hi.w;
                };
#pragma unroll 4
                for (int j = t; j < A_BYTES / 16; j += 128) { float4 hi, lo; split4(a_hi[j], hi, lo); if (keep_hi) a_hi[j] = hi; a_lo[j] = lo; }
#pragma unroll 4
                for (int j = t; j < B_BYTES / 16; j += 128) { float4 hi, lo; split4(b_hi[j], hi, lo); if (keep_hi) b_hi[j] = hi; b_lo[j] = lo; }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(conv_bar(s));
            }
        }
        // push the partial accumulator: thread <-> row (TMEM lane); row r goes to slot [z][r % rows_per] of CTA r / rows_per
        if (g.trace && nkb > 0 && nkb <= STAGES) { mbar_wait(full_bar(0), 0); stamp(2); }     // first operand tiles have landed
        if (nkb > 0) {
            mbar_wait(accum_bar, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        stamp(3);
        const int q = warp & 3;
        const int row = q * 32 + lane;
        float v[64];
        if (nkb > 0) tmem_ld64(tmem_base + ((uint32_t)(q * 32) << 16), v);
        else {
#pragma unroll
            for (int i = 0; i < 64; ++i) v[i] = 0.f;
        }
        const int rows_per = BM / KS;
        const uint32_t dst = park + (uint32_t)((blockIdx.z * rows_per + (row % rows_per)) * D_PARK_LD) * 4u;
        const uint32_t owner = (uint32_t)(row / rows_per);
#pragma unroll
        for (int i = 0; i < 16; ++i) st_dsmem_v4(dst + 16u * i, owner, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
        stamp(4);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                                          // all KS partials of my rows have landed in my buffer
    stamp(5);
    {
        // CTA z folds rows [z*rows_per, ...) from its OWN shared memory (slices 0..KS-1 in a fixed order: deterministic),
        // applies bias + ReLU (or bias + de-standardisation, SMC:533) and stores 256-byte row segments
        const int rows_per = BM / KS;
        const float* buf = reinterpret_cast<const float*>(gen_base + (park - base));
        for (int e = threadIdx.x; e < rows_per * (BN / 4); e += kThreads) {
            const int rl = e / (BN / 4), c4 = e % (BN / 4);
            float4 acc = *reinterpret_cast<const float4*>(buf + rl * D_PARK_LD + c4 * 4);
            for (int p = 1; p < KS; ++p) {
                const float4 w = *reinterpret_cast<const float4*>(buf + (p * rows_per + rl) * D_PARK_LD + c4 * 4);
                acc.x += w.x; acc.y += w.y; acc.z += w.z; acc.w += w.w;
            }
            const int n = n0 + c4 * 4, m = m0 + blockIdx.z * rows_per + rl;
            const float4 b = e_b;                                // kThreads % (BN / 4) == 0: c4 == my_c4 for every e of this thread
            float4 o;
            if (g.epi == EPI_BIAS_RELU) {
                o = make_float4(fmaxf(acc.x + b.x, 0.f), fmaxf(acc.y + b.y, 0.f), fmaxf(acc.z + b.z, 0.f), fmaxf(acc.w + b.w, 0.f));
            } else {
                const float4 sc = e_sc, sh = e_sh;
                o = make_float4((acc.x + b.x) * sc.x + sh.x, (acc.y + b.y) * sc.y + sh.y, (acc.z + b.z) * sc.z + sh.z, (acc.w + b.w) * sc.w + sh.w);
            }
            if (m < g.M) {
                const size_t off = (size_t)m * g.ldc + n;
                *reinterpret_cast<float4*>(g.C + off) = o;
                if (g.C_lo) {                                    // pre-split operand of the next contraction (C_hi only for the PCA inverse)
                    float4 hi;
                    hi.x = __uint_as_float(__float_as_uint(o.x) & 0xFFFFE000u); hi.y = __uint_as_float(__float_as_uint(o.y) & 0xFFFFE000u);
                    hi.z = __uint_as_float(__float_as_uint(o.z) & 0xFFFFE000u); hi.w = __uint_as_float(__float_as_uint(o.w) & 0xFFFFE000u);
                    if (g.C_hi) *reinterpret_cast<float4*>(g.C_hi + off) = hi;
                    *reinterpret_cast<float4*>(g.C_lo + off) = make_float4(o.x - hi.x, o.y - hi.y, o.z - hi.z, o.w - hi.w);
                }
            }
        }
    }
    __syncthreads();
    stamp(6);
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
    }
}

int dense_cluster_prepare() {
    return cudaFuncSetAttribute(dense_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, D_SMEM) == cudaSuccess ? 0 : -1;
}

// t.args.splits = KS (1, 2, 4 or 8: the cluster size along z); epi = EPI_BIAS_RELU or EPI_BIAS_AFFINE
int launch_dense_cluster(const TcGemm& t, cudaStream_t s) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(t.args.N / D_BN, (t.args.M + BM - 1) / BM, t.args.splits);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = D_SMEM;
    cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = t.args.splits;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    const CUtensorMap& a = *reinterpret_cast<const CUtensorMap*>(&t.mapA);
    const CUtensorMap& b = *reinterpret_cast<const CUtensorMap*>(&t.mapB);
    const CUtensorMap& al = *reinterpret_cast<const CUtensorMap*>(t.args.presplit ? &t.mapAlo : &t.mapA);
    const CUtensorMap& bl = *reinterpret_cast<const CUtensorMap*>(t.args.presplit ? &t.mapBlo : &t.mapB);
    return cudaLaunchKernelEx(&cfg, dense_cluster_kernel, a, b, al, bl, t.args) == cudaSuccess ? 0 : -1;
}

// ------------------------------------------------------------------------------------------------
// PCA projection (SMC:494) + standardisation (SMC:505-523) as ONE launch -- see ProjArgs in psm_kernels.cuh.
// Main loop = tc_gemm_kernel<128> (TMA -> 3xTF32 split -> tcgen05.mma, 3-stage ring); the epilogue replaces the
// [splits][M][N] partials in HBM + the reduce launch by two on-chip levels:
//   1. after a cluster barrier (all main loops of the cluster are done: the ring memory is free) every CTA pushes row slab s of
//      its 128 x 128 partial into CTA s's ring memory through distributed shared memory; CTA z then folds the ks slices of
//      slab z in a fixed order and stores the cluster partial (L2-resident, ks times smaller than before);
//   2. one counter per (tile, slab): the LAST cluster to store slab z folds the ncl cluster partials of that slab in a fixed
//      order, adds zc, standardises and writes the Dense input.  No spinning anywhere; the result does not depend on which
//      cluster arrives last (bit-identical replays).
namespace {
constexpr int P_BN = 128, P_STAGES = 3;
constexpr int P_A_BYTES = BM * BK * 4, P_B_BYTES = P_BN * BK * 4, P_STAGE = 2 * (P_A_BYTES + P_B_BYTES);
constexpr int P_PARK_LD = P_BN + 4;                                  // floats per parked row (conflict-free float4 rows)
constexpr int P_SMEM = P_STAGES * P_STAGE + 1024 + 256;
static_assert(BM * P_PARK_LD * 4 <= P_STAGES * P_STAGE, "the receive buffer re-uses the operand ring");
}  // namespace

__global__ void __launch_bounds__(kThreads, 1)
proj_cluster_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, ProjArgs g) {
    constexpr int BN = P_BN, STAGES = P_STAGES, A_BYTES = P_A_BYTES, B_BYTES = P_B_BYTES, STAGE = P_STAGE;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + STAGES * STAGE;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto conv_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    auto empty_bar = [&](int s) { return bars + 8u * (2 * STAGES + s); };
    const uint32_t accum_bar = bars + 8u * (3 * STAGES);
    const uint32_t tmem_slot = accum_bar + 8u;
    uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));

    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KS = g.ks;
    const int z = blockIdx.z % KS, cl = blockIdx.z / KS, ncl = g.splits / KS;   // cluster = (1,1,KS): consecutive z
    const int m0 = blockIdx.y * BM;
    const int kb_total = g.K / BK;
    const int kb_per = (kb_total + g.splits - 1) / g.splits;
    const int kb0 = min(blockIdx.z * kb_per, kb_total);
    const int kb1 = min(kb_total, kb0 + kb_per);
    const int nkb = kb1 - kb0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(conv_bar(s), 128); mbar_init(empty_bar(s), 1); }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - base));
    // the PCA matrix is static: its tiles of the first ring pass are requested before waiting for the gather
    const int n_pre = g.b_static ? min(nkb, STAGES) : 0;
    if (threadIdx.x == 0) {
        for (int it = 0; it < n_pre; ++it) {
            mbar_expect_tx(full_bar(it), A_BYTES + B_BYTES);
            tma_load_2d(base + it * STAGE + 2 * A_BYTES, &tmB, (kb0 + it) * BK, 0, full_bar(it));
        }
    }
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < nkb; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                const uint32_t st = base + s * STAGE;
                if (it >= n_pre) {
                    mbar_wait(empty_bar(s), ph ^ 1);
                    mbar_expect_tx(full_bar(s), A_BYTES + B_BYTES);
                    tma_load_2d(st + 2 * A_BYTES, &tmB, (kb0 + it) * BK, 0, full_bar(s));
                }
                tma_load_2d(st, &tmA, (kb0 + it) * BK, m0, full_bar(s));
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(BM, BN);
            for (int it = 0; it < nkb; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(g.three_pass ? conv_bar(s) : full_bar(s), ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = base + s * STAGE;
                const uint32_t a_hi = st, a_lo = st + A_BYTES, b_hi = st + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    const uint32_t ko = k * UMMA_K * 4;
                    umma_tf32(tmem_base, make_smem_desc(a_hi + ko), make_smem_desc(b_hi + ko), idesc, (it | k) != 0);
                    if (g.three_pass) {
                        umma_tf32(tmem_base, make_smem_desc(a_hi + ko), make_smem_desc(b_lo + ko), idesc, 1);
                        umma_tf32(tmem_base, make_smem_desc(a_lo + ko), make_smem_desc(b_hi + ko), idesc, 1);
                    }
                }
                umma_commit(empty_bar(s));
            }
            umma_commit(accum_bar);
        }
        __syncwarp();
    } else {
        const int t = threadIdx.x - 64;
        const bool keep_hi = g.three_pass != 2;
        if (g.three_pass) {
            for (int it = 0; it < nkb; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(full_bar(s), ph);
                uint8_t* st = gen_base + s * STAGE;
                float4* a_hi = reinterpret_cast<float4*>(st);
                float4* a_lo = reinterpret_cast<float4*>(st + A_BYTES);
                float4* b_hi = reinterpret_cast<float4*>(st + 2 * A_BYTES);
                float4* b_lo = reinterpret_cast<float4*>(st + 2 * A_BYTES + B_BYTES);
                auto split4 = [](float4 v, float4& hi, float4& lo) {
                    hi.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); lo.x = v.x - hi.x;
                    hi.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); lo.y = v.y - hi.y;
                    hi.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); lo.z = v.z - hi.z;
                    hi.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); lo.w = v.w - hi.w;
                };
#pragma unroll 4
                for (int j = t; j < A_BYTES / 16; j += 128) { float4 hi, lo; split4(a_hi[j], hi, lo); if (keep_hi) a_hi[j] = hi; a_lo[j] = lo; }
#pragma unroll 4
                for (int j = t; j < B_BYTES / 16; j += 128) { float4 hi, lo; split4(b_hi[j], hi, lo); if (keep_hi) b_hi[j] = hi; b_lo[j] = lo; }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(conv_bar(s));
            }
        }
        if (nkb > 0) mbar_wait(accum_bar, 0);                    // my MMAs have retired: my ring memory is free
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                                          // every main loop of the cluster is done: ring memory may be written remotely
    const int rows_per = BM / KS;
    if (warp >= 2) {
        // push: thread <-> row (TMEM lane); row r goes to slot [z][r % rows_per] of CTA r / rows_per
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t dst = base + (uint32_t)((z * rows_per + (row % rows_per)) * P_PARK_LD) * 4u;
        const uint32_t owner = (uint32_t)(row / rows_per);
#pragma unroll
        for (int c = 0; c < BN; c += 32) {
            float v[32];
            if (nkb > 0) tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
            else {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0.f;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) st_dsmem_v4(dst + 4u * c + 16u * i, owner, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    cluster_sync_all();                                          // all KS slices of my slab have landed
    {
        const float* buf = reinterpret_cast<const float*>(gen_base);
        const size_t slab = (size_t)(m0 + z * rows_per) * BN;    // first element of my slab inside one [M][N] partial
        float* mine = g.part + (size_t)cl * g.M * BN + slab;
        for (int e = threadIdx.x; e < rows_per * (BN / 4); e += kThreads) {
            const int rl = e / (BN / 4), c4 = e % (BN / 4);
            float4 acc = *reinterpret_cast<const float4*>(buf + rl * P_PARK_LD + c4 * 4);
            for (int p = 1; p < KS; ++p) {
                const float4 w = *reinterpret_cast<const float4*>(buf + (p * rows_per + rl) * P_PARK_LD + c4 * 4);
                acc.x += w.x; acc.y += w.y; acc.z += w.z; acc.w += w.w;
            }
            __stcg(reinterpret_cast<float4*>(mine + (size_t)rl * BN + c4 * 4), acc);
        }
        __shared__ bool s_last;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned int* cnt = g.counters + blockIdx.y * KS + z;
            const unsigned int prev = atomicAdd(cnt, 1u);
            s_last = (prev == (unsigned int)(ncl - 1));
            if (s_last) { *cnt = 0u; __threadfence(); }          // re-armed for the next step
        }
        __syncthreads();
        if (s_last) {
            for (int e = threadIdx.x; e < rows_per * (BN / 4); e += kThreads) {
                const int rl = e / (BN / 4), c4 = e % (BN / 4);
                const size_t off = slab + (size_t)rl * BN + c4 * 4;
                float4 acc = __ldcg(reinterpret_cast<const float4*>(g.part + off));
                for (int c = 1; c < ncl; ++c) {
                    const float4 w = __ldcg(reinterpret_cast<const float4*>(g.part + (size_t)c * g.M * BN + off));
                    acc.x += w.x; acc.y += w.y; acc.z += w.z; acc.w += w.w;
                }
                const float4 zc = __ldg(reinterpret_cast<const float4*>(g.zc + off));
                const float4 sa = __ldg(reinterpret_cast<const float4*>(g.a + c4 * 4));
                const float4 sb = __ldg(reinterpret_cast<const float4*>(g.b + c4 * 4));
                const float4 o = make_float4((acc.x + zc.x) * sa.x + sb.x, (acc.y + zc.y) * sa.y + sb.y,
                                             (acc.z + zc.z) * sa.z + sb.z, (acc.w + zc.w) * sa.w + sb.w);
                *reinterpret_cast<float4*>(g.x + off) = o;
                if (g.x_lo) {                                    // pre-split operand of the first Dense layer
                    float4 hi;
                    hi.x = __uint_as_float(__float_as_uint(o.x) & 0xFFFFE000u); hi.y = __uint_as_float(__float_as_uint(o.y) & 0xFFFFE000u);
                    hi.z = __uint_as_float(__float_as_uint(o.z) & 0xFFFFE000u); hi.w = __uint_as_float(__float_as_uint(o.w) & 0xFFFFE000u);
                    *reinterpret_cast<float4*>(g.x_lo + off) = make_float4(o.x - hi.x, o.y - hi.y, o.z - hi.z, o.w - hi.w);
                }
            }
        }
    }
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
    }
}

static void proj_launch_cfg(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* at, int tiles, int splits, int ks, cudaStream_t s, bool pdl) {
    cfg.gridDim = dim3(1, tiles, splits);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = P_SMEM;
    cfg.stream = s;
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = ks;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 2 : 1;
}

// Largest number of ks-CTA clusters the device holds at once (the split is sized to one resident wave).
int proj_cluster_prepare(int tiles, int ks, int* max_clusters) {
    if (cudaFuncSetAttribute(proj_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM) != cudaSuccess) return -1;
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute at[2];
    proj_launch_cfg(cfg, at, tiles, ks * 64, ks, 0, false);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, proj_cluster_kernel, &cfg) != cudaSuccess || n < 1) { cudaGetLastError(); return -1; }
    *max_clusters = n;
    return 0;
}

int launch_proj_cluster(const ProjGemm& t, cudaStream_t s) {
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute at[2];
    proj_launch_cfg(cfg, at, t.args.M / BM, t.args.splits, t.args.ks, s, pdl_enabled());
    const CUtensorMap& a = *reinterpret_cast<const CUtensorMap*>(&t.mapA);
    const CUtensorMap& b = *reinterpret_cast<const CUtensorMap*>(&t.mapB);
    return cudaLaunchKernelEx(&cfg, proj_cluster_kernel, a, b, t.args) == cudaSuccess ? 0 : -1;
}

// ------------------------------------------------------------------------------------------------
// The whole Dense stack in one launch (see psm_kernels.cuh).  192 threads: warp 0 = TMA producer, warp 1 = TMEM
// allocator + MMA issuer, warps 2..5 = TMEM -> shared-memory parking; all six warps fold the cluster's partials.
namespace {
constexpr int S_BN = 64;
constexpr int S_KS = 8;                                 // cluster size = K slices per output tile
constexpr int S_STAGES = 3;
constexpr int S_A_BYTES = BM * BK * 4, S_B_BYTES = S_BN * BK * 4;
constexpr int S_STAGE = 2 * (S_A_BYTES + S_B_BYTES);    // A_hi | A_lo | B_hi | B_lo
constexpr int S_PARK_LD = S_BN + 4;                     // floats per parked row: 16-byte aligned, conflict-free float4 rows
constexpr int S_PARK = BM * S_PARK_LD * 4;
constexpr int S_SMEM = S_STAGES * S_STAGE + 2 * S_PARK + 1024 + 256;

__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
}  // namespace

__global__ void __launch_bounds__(kThreads, 1) dense_stack_kernel(const __grid_constant__ DenseStackArgs g) {
    constexpr int BN = S_BN, KS = S_KS, STAGES = S_STAGES, A_BYTES = S_A_BYTES, B_BYTES = S_B_BYTES, STAGE = S_STAGE;
    constexpr int ROWS = BM / KS;                                // rows of a tile folded by each CTA of the cluster
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t park0 = base + STAGES * STAGE;
    const uint32_t bars = park0 + 2 * S_PARK;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    const uint32_t accum_bar = bars + 8u * (2 * STAGES);
    const uint32_t tmem_slot = accum_bar + 8u;
    uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));

    pdl_launch_dependents();
    int tr_n = 0;
    auto stamp = [&]() {
        if (g.trace && threadIdx.x == 64 && tr_n < 64) {
            unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            g.trace[(size_t)blockIdx.x * 64 + tr_n++] = t;
        }
    };
    stamp();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int z = blockIdx.x % KS;                               // rank in the cluster (cluster = 8 consecutive CTAs along x)
    const int cl = blockIdx.x / KS, n_cl = gridDim.x / KS;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - base));
    cluster_sync_all();          // every CTA of the cluster has started: its shared memory may be written remotely
    if (warp == 0 && lane == 0) {
        // the weights do not depend on the previous kernel: pull every tile this CTA will use into L2 while it drains
        for (int l = 0; l < g.n_layers; ++l) {
            const DenseLayerDesc& L = g.L[l];
            const int kb_total = L.K / BK, kb_per = (kb_total + KS - 1) / KS;
            const int kb0 = z * kb_per, nkb = max(min(kb_total - kb0, kb_per), 0);
            const int n_tiles = L.N / BN;
            for (int t = cl; t < n_tiles; t += n_cl)             // weight tiles repeat for every row tile
                for (int it = 0; it < nkb; ++it) {
                    tma_prefetch_l2_2d(reinterpret_cast<const CUtensorMap*>(g.maps + 4 * l + 2), (kb0 + it) * BK, t * BN);
                    tma_prefetch_l2_2d(reinterpret_cast<const CUtensorMap*>(g.maps + 4 * l + 3), (kb0 + it) * BK, t * BN);
                }
        }
    }
    stamp();
    pdl_wait();
    stamp();

    uint32_t it_pipe = 0;        // k-blocks this CTA has pushed through the pipeline (same count in producer and issuer)
    uint32_t acc_phase = 0;      // completed accumulations of this CTA
    uint32_t pb = 0;             // parking buffer of the current tile
    const uint32_t idesc = make_idesc_tf32(BM, BN);

    for (int l = 0; l < g.n_layers; ++l) {
        const DenseLayerDesc& L = g.L[l];
        const CUtensorMap* mA_hi = reinterpret_cast<const CUtensorMap*>(g.maps + 4 * l + 0);
        const CUtensorMap* mA_lo = reinterpret_cast<const CUtensorMap*>(g.maps + 4 * l + 1);
        const CUtensorMap* mB_hi = reinterpret_cast<const CUtensorMap*>(g.maps + 4 * l + 2);
        const CUtensorMap* mB_lo = reinterpret_cast<const CUtensorMap*>(g.maps + 4 * l + 3);
        const int kb_total = L.K / BK;
        const int kb_per = (kb_total + KS - 1) / KS;
        const int active = (kb_total + kb_per - 1) / kb_per;      // CTAs of the cluster that own >= 1 k-block
        const int kb0 = z * kb_per;
        const int nkb = max(min(kb_total - kb0, kb_per), 0);
        const int n_tiles = L.N / BN;
        const int tiles = (g.M / BM) * n_tiles;
        for (int t = cl; t < tiles; t += n_cl) {
            const int m0 = (t / n_tiles) * BM, n0 = (t % n_tiles) * BN;
            if (warp == 0) {
                if (lane == 0) {
                    for (int it = 0; it < nkb; ++it) {
                        const uint32_t i = it_pipe + it;
                        const int s = i % STAGES;
                        mbar_wait(empty_bar(s), ((i / STAGES) & 1) ^ 1);
                        mbar_expect_tx(full_bar(s), 2 * (A_BYTES + B_BYTES));
                        const uint32_t st = base + s * STAGE;
                        const int kc = (kb0 + it) * BK;
                        tma_load_2d(st, mA_hi, kc, m0, full_bar(s));
                        tma_load_2d(st + A_BYTES, mA_lo, kc, m0, full_bar(s));
                        tma_load_2d(st + 2 * A_BYTES, mB_hi, kc, n0, full_bar(s));
                        tma_load_2d(st + 2 * A_BYTES + B_BYTES, mB_lo, kc, n0, full_bar(s));
                    }
                }
                __syncwarp();
            } else if (warp == 1) {
                if (lane == 0) {
                    for (int it = 0; it < nkb; ++it) {
                        const uint32_t i = it_pipe + it;
                        const int s = i % STAGES;
                        mbar_wait(full_bar(s), (i / STAGES) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t st = base + s * STAGE;
                        const uint32_t a_hi = st, a_lo = st + A_BYTES, b_hi = st + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            const uint32_t ko = k * UMMA_K * 4;
                            umma_tf32(tmem_base, make_smem_desc(a_hi + ko), make_smem_desc(b_hi + ko), idesc, (it | k) != 0);
                            if (g.three_pass) {
                                umma_tf32(tmem_base, make_smem_desc(a_hi + ko), make_smem_desc(b_lo + ko), idesc, 1);
                                umma_tf32(tmem_base, make_smem_desc(a_lo + ko), make_smem_desc(b_hi + ko), idesc, 1);
                            }
                        }
                        umma_commit(empty_bar(s));
                    }
                    if (nkb > 0) umma_commit(accum_bar);
                }
                __syncwarp();
            } else if (nkb > 0) {
                mbar_wait(accum_bar, acc_phase & 1);
                stamp();
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int q = warp & 3;
                const int row = q * 32 + lane;
                float v[64];
                tmem_ld64(tmem_base + ((uint32_t)(q * 32) << 16), v);
                stamp();
                // push this row of the partial into the CTA that folds it: slot [z][row % 16] of CTA row / 16
                const uint32_t dst = park0 + pb * S_PARK + (uint32_t)((z * ROWS + (row % ROWS)) * S_PARK_LD) * 4u;
                const uint32_t owner = (uint32_t)(row / ROWS);
#pragma unroll
                for (int i = 0; i < 16; ++i) st_dsmem_v4(dst + 16u * i, owner, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            }
            it_pipe += nkb;
            if (nkb > 0) ++acc_phase;
            stamp();
            cluster_sync_all();                                  // every partial of this tile has landed in its folder's buffer
            stamp();
            {
                // CTA z folds rows [z*16, z*16+16) x 64 columns from its OWN shared memory: slots 0..active-1, fixed order
                const float* park = reinterpret_cast<const float*>(gen_base + (park0 + pb * S_PARK - base));
                for (int e = threadIdx.x; e < ROWS * (BN / 4); e += kThreads) {
                    const int rl = e / (BN / 4), c4 = e % (BN / 4);
                    const int r = z * ROWS + rl;
                    float4 acc = *reinterpret_cast<const float4*>(park + rl * S_PARK_LD + c4 * 4);
                    for (int p = 1; p < active; ++p) {
                        const float4 w = *reinterpret_cast<const float4*>(park + (p * ROWS + rl) * S_PARK_LD + c4 * 4);
                        acc.x += w.x; acc.y += w.y; acc.z += w.z; acc.w += w.w;
                    }
                    const int n = n0 + c4 * 4;
                    const size_t o = (size_t)(m0 + r) * L.N + n;
                    const float4 b = __ldg(reinterpret_cast<const float4*>(L.bias + n));
                    float4 y;
                    if (L.epi == EPI_BIAS_RELU) {
                        y = make_float4(fmaxf(acc.x + b.x, 0.f), fmaxf(acc.y + b.y, 0.f), fmaxf(acc.z + b.z, 0.f), fmaxf(acc.w + b.w, 0.f));
                    } else {
                        const float4 sc = __ldg(reinterpret_cast<const float4*>(L.v1 + n));
                        const float4 sh = __ldg(reinterpret_cast<const float4*>(L.v2 + n));
                        y = make_float4((acc.x + b.x) * sc.x + sh.x, (acc.y + b.y) * sc.y + sh.y, (acc.z + b.z) * sc.z + sh.z,
                                        (acc.w + b.w) * sc.w + sh.w);
                    }
                    if (L.out) *reinterpret_cast<float4*>(L.out + o) = y;
                    if (L.out_hi) {
                        float4 hi;
                        hi.x = __uint_as_float(__float_as_uint(y.x) & 0xFFFFE000u); hi.y = __uint_as_float(__float_as_uint(y.y) & 0xFFFFE000u);
                        hi.z = __uint_as_float(__float_as_uint(y.z) & 0xFFFFE000u); hi.w = __uint_as_float(__float_as_uint(y.w) & 0xFFFFE000u);
                        *reinterpret_cast<float4*>(L.out_hi + o) = hi;
                        *reinterpret_cast<float4*>(L.out_lo + o) = make_float4(y.x - hi.x, y.y - hi.y, y.z - hi.z, y.w - hi.w);
                    }
                }
            }
            stamp();
            pb ^= 1;             // the other buffer is free: its readers passed the cluster barrier above
        }
        if (l + 1 < g.n_layers) {
            // grid barrier: this layer's activations (generic-proxy stores) are read by the next layer's TMA loads
            asm volatile("fence.proxy.async;" ::: "memory");
            __syncthreads();
            if (threadIdx.x == 0) {
                const unsigned int target = (unsigned int)(l + 1) * gridDim.x;
                __threadfence();
                atomicAdd(g.barrier, 1u);
                const long long t0 = clock64();
                while (ld_acquire_gpu(g.barrier) < target) {
                    if (clock64() - t0 > 2000000000ll) { *g.error = 1; break; }
                }
                __threadfence();
            }
            __syncthreads();
            asm volatile("fence.proxy.async;" ::: "memory");
            stamp();
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                                          // nobody leaves while its shared memory is being read
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
    }
}

// Largest number of 8-CTA clusters of dense_stack_kernel the device holds at once (the kernel's grid barrier needs
// every CTA resident).
int dense_stack_prepare(int* max_clusters) {
    if (cudaFuncSetAttribute(dense_stack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S_SMEM) != cudaSuccess) return -1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(S_KS * 16); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = S_SMEM;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = S_KS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, dense_stack_kernel, &cfg) != cudaSuccess || n < 1) { cudaGetLastError(); return -1; }
    *max_clusters = n;
    return 0;
}

int launch_dense_stack(const DenseStackArgs& a, int clusters, cudaStream_t s) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(S_KS * clusters, 1, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = S_SMEM;
    cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = S_KS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, dense_stack_kernel, a) == cudaSuccess ? 0 : -1;
}

// ------------------------------------------------------------------------------------------------
// PCA inverse (SMC:541,551), transposed:  blocks^T[pixel][block] = comp_out_t[pixel][K] * r[block][K]^T.
// One CTA per 128 output pixels: its 128 x K slice of the output PCA matrix is read from HBM exactly once (the
// A operand, split hi/lo on chip), the de-standardised coordinates of ALL blocks stream through as the B operand
// (pre-split by the last Dense layer, L2-resident), and the accumulator -- TMEM lane = pixel, column = block -- is
// written with fully coalesced rows: for a fixed block the 32 lanes of a warp store 32 consecutive pixels.
// Four 128-column accumulators rotate in TMEM, so the epilogue of one block chunk overlaps the MMAs of the next.
namespace {
constexpr int I_KB = 4;                                  // k-blocks of the A slice held in shared memory (K <= 128)
constexpr int I_NB = 128;                                // blocks per chunk = UMMA N
constexpr int I_A_TILE = BM * BK * 4;                    // 16 KB
constexpr int I_B_TILE = I_NB * BK * 4;                  // 16 KB
constexpr int I_BSTAGES = 2;
constexpr int I_TILE_LD = 129;                           // floats per parked block row (strip sums): odd pitch, conflict-free columns
constexpr int I_TILE = 32 * I_TILE_LD * 4;
constexpr int I_SMEM = 2 * I_KB * I_A_TILE + I_BSTAGES * 2 * I_B_TILE + 1024 + 256 + I_TILE;
}  // namespace

__global__ void __launch_bounds__(kThreads, 1)
pca_inverse_t_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
                     const __grid_constant__ CUtensorMap tmBlo, InvTArgs g) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_hi = base, a_lo = base + I_KB * I_A_TILE, b_ring = base + 2 * I_KB * I_A_TILE;
    const uint32_t bars = b_ring + I_BSTAGES * 2 * I_B_TILE;
    const uint32_t a_full = bars, a_conv = bars + 8;
    auto b_full = [&](int s) { return bars + 16u + 8u * s; };
    auto b_empty = [&](int s) { return bars + 16u + 8u * (I_BSTAGES + s); };
    auto acc_full = [&](int b) { return bars + 16u + 8u * (2 * I_BSTAGES + b); };
    auto acc_empty = [&](int b) { return bars + 16u + 8u * (2 * I_BSTAGES + 4 + b); };
    const uint32_t tmem_slot = bars + 16u + 8u * (2 * I_BSTAGES + 8);
    const uint32_t tile_off = bars + 256u;                       // [32][I_TILE_LD] floats: the strip-sum tile
    uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));

    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p0 = blockIdx.x * BM;                               // first output pixel (row of comp_out_t) of this CTA
    const int kb_n = g.K / BK;                                    // <= I_KB
    const int n_chunks = g.Mb / I_NB;                             // block chunks (Mb = padded block count)

    if (threadIdx.x == 0) {
        mbar_init(a_full, 1); mbar_init(a_conv, 128);
        for (int s = 0; s < I_BSTAGES; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
        for (int b = 0; b < 4; ++b) { mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - base));

    if (warp == 0) {
        if (lane == 0) {
            // the PCA matrix is static: its slice is requested before waiting for the Dense stack
            mbar_expect_tx(a_full, (uint32_t)(kb_n * I_A_TILE));
            for (int kb = 0; kb < kb_n; ++kb) tma_load_2d(a_hi + kb * I_A_TILE, &tmA, kb * BK, p0, a_full);
            pdl_wait();
            uint32_t it = 0;
            for (int c = 0; c < n_chunks; ++c)
                for (int kb = 0; kb < kb_n; ++kb, ++it) {
                    const int s = it % I_BSTAGES;
                    mbar_wait(b_empty(s), ((it / I_BSTAGES) & 1) ^ 1);
                    mbar_expect_tx(b_full(s), 2 * I_B_TILE);
                    const uint32_t st = b_ring + s * 2 * I_B_TILE;
                    tma_load_2d(st, &tmBhi, kb * BK, c * I_NB, b_full(s));
                    tma_load_2d(st + I_B_TILE, &tmBlo, kb * BK, c * I_NB, b_full(s));
                }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(BM, I_NB);
            mbar_wait(a_conv, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t it = 0;
            for (int c = 0; c < n_chunks; ++c) {
                const int buf = c & 3;
                mbar_wait(acc_empty(buf), ((c >> 2) & 1) ^ 1);   // the epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tacc = tmem_base + (uint32_t)(buf * I_NB);
                for (int kb = 0; kb < kb_n; ++kb, ++it) {
                    const int s = it % I_BSTAGES;
                    mbar_wait(b_full(s), (it / I_BSTAGES) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = b_ring + s * 2 * I_B_TILE;
                    const uint32_t ah = a_hi + kb * I_A_TILE, al = a_lo + kb * I_A_TILE, bh = st, bl = st + I_B_TILE;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint32_t ko = k * UMMA_K * 4;
                        umma_tf32(tacc, make_smem_desc(ah + ko), make_smem_desc(bh + ko), idesc, (kb | k) != 0);
                        if (g.three_pass) {
                            umma_tf32(tacc, make_smem_desc(ah + ko), make_smem_desc(bl + ko), idesc, 1);
                            umma_tf32(tacc, make_smem_desc(al + ko), make_smem_desc(bh + ko), idesc, 1);
                        }
                    }
                    umma_commit(b_empty(s));
                }
                umma_commit(acc_full(buf));
            }
        }
        __syncwarp();
    } else {
        // ---- converters: split the A slice once (hi in place, lo beside it) ----
        const int t = threadIdx.x - 64;
        const bool keep_hi = g.three_pass != 2;
        mbar_wait(a_full, 0);
        {
            float4* hi4 = reinterpret_cast<float4*>(gen_base + (a_hi - base));
            float4* lo4 = reinterpret_cast<float4*>(gen_base + (a_lo - base));
            const int n4 = kb_n * I_A_TILE / 16;
#pragma unroll 4
            for (int i = t; i < n4; i += 128) {
                const float4 v = hi4[i];
                float4 h, l;
                h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
                h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
                h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
                h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
                if (keep_hi) hi4[i] = h;
                lo4[i] = l;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(a_conv);
        }
        // ---- epilogue: thread <-> pixel (TMEM lane); 32 blocks per tcgen05.ld; coalesced 128 B rows per block ----
        const int q = warp & 3;
        // strip sums (StripRows): the 32 blocks x 128 pixels just read from TMEM are parked in a shared-memory tile; the 128
        // epilogue threads then split this pixel row's (entry, warp-quarter) pairs of those 32 blocks among themselves -- every
        // pair is an independent masked sum of 32 tile values (fixed order: deterministic), so the sums run at full thread-level
        // parallelism instead of one dependent shuffle chain per entry.
        const StripRows& sr = g.strips;
        const bool strips = sr.row_ptr != nullptr;
        float* tile = reinterpret_cast<float*>(gen_base + (tile_off - base));
        const int n_c32 = (g.B + 31) / 32;
        const int32_t* rp = strips ? sr.row_ptr + (size_t)blockIdx.x * (n_c32 + 1) : nullptr;
        // the (entry, quarter) pairs of a 32-block chunk are static: each thread keeps the metadata of its (up to KP) pairs of the
        // NEXT chunk in registers, requested one chunk ahead (the first one before the wait), so no L2 round trip sits between
        // the two barriers of a chunk
        constexpr int KP = 4;
        uint32_t nw[KP]; int nsrc[KP], nslot[KP]; int n_pairs_next = 0, e_lo_next = 0;
        auto prefetch = [&](int cc) {
            n_pairs_next = 0;
            if (cc >= n_c32) return;
            e_lo_next = __ldg(rp + cc);
            n_pairs_next = (__ldg(rp + cc + 1) - e_lo_next) * 4;
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                const int pr = t + 128 * k;
                nw[k] = 0u; nsrc[k] = 0; nslot[k] = 0;
                if (pr < n_pairs_next) {
                    const int e = e_lo_next + (pr >> 2);
                    nw[k] = __ldg(sr.w + (size_t)(pr & 3) * sr.n_ent + e);
                    nsrc[k] = __ldg(sr.src + e); nslot[k] = __ldg(sr.slot + e);
                }
            }
        };
        if (strips) prefetch(0);
        pdl_wait();                                              // g.sc->out_scale belongs to this step
        const int lx = q * 32 + lane;
        const int P = p0 + lx;                                   // planar pixel index (c*S*S + ly*S + lx)
        const float pm = __ldg(g.pmean + P);
        const float o_scale = g.sc->out_scale;
        auto pair_sum = [&](uint32_t wq, const float* row, int qq) {
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int ii = (i + 8 * qq) & 31;               // rotated start: the four quarters of one entry hit different banks
                if ((wq >> ii) & 1u) acc += row[ii];
            }
            return acc;
        };
        for (int c = 0; c < n_chunks; ++c) {
            const int buf = c & 3;
            mbar_wait(acc_full(buf), (c >> 2) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int j0 = 0; j0 < I_NB; j0 += 32) {
                const int b0 = c * I_NB + j0;
                if (b0 >= g.B) break;                            // padding blocks are never stored
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * I_NB + j0), v);
                float* dst = g.blocks + (size_t)b0 * g.block_stride + P;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (b0 + j < g.B) {
                        const float o = (v[j] + pm) * o_scale;
                        dst[(size_t)j * g.block_stride] = o;
                        if (strips) tile[j * I_TILE_LD + lx] = o;
                    }
                }
                if (strips) {
                    // this chunk's pair metadata moves out of the prefetch registers, the next chunk's is requested
                    uint32_t cw[KP]; int csrc[KP], cslot[KP];
#pragma unroll
                    for (int k = 0; k < KP; ++k) { cw[k] = nw[k]; csrc[k] = nsrc[k]; cslot[k] = nslot[k]; }
                    const int n_pairs = n_pairs_next, e_lo = e_lo_next;
                    prefetch((b0 >> 5) + 1);
                    asm volatile("bar.sync 1, 128;" ::: "memory");               // the tile of these 32 blocks is complete
#pragma unroll
                    for (int k = 0; k < KP; ++k) {
                        const int qq = (t + 128 * k) & 3;
                        if (cw[k]) sr.rowpart[(size_t)cslot[k] * 4 + qq] = pair_sum(cw[k], tile + (csrc[k] - b0) * I_TILE_LD + qq * 32, qq);
                    }
                    for (int pr = t + 128 * KP; pr < n_pairs; pr += 128) {       // more than KP pairs per thread: plain loads
                        const int e = e_lo + (pr >> 2), qq = pr & 3;
                        const uint32_t wq = __ldg(sr.w + (size_t)qq * sr.n_ent + e);
                        if (wq) sr.rowpart[(size_t)__ldg(sr.slot + e) * 4 + qq] = pair_sum(wq, tile + (__ldg(sr.src + e) - b0) * I_TILE_LD + qq * 32, qq);
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");               // the tile may be overwritten
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(acc_empty(buf));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

int pca_inverse_t_prepare() {
    return cudaFuncSetAttribute(pca_inverse_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, I_SMEM) == cudaSuccess ? 0 : -1;
}
int pca_inverse_t_max_k() { return I_KB * BK; }

void launch_pca_inverse_t(const InvT& t, cudaStream_t s) {
    const CUtensorMap& a = *reinterpret_cast<const CUtensorMap*>(&t.mapA);
    const CUtensorMap& bh = *reinterpret_cast<const CUtensorMap*>(&t.mapBhi);
    const CUtensorMap& bl = *reinterpret_cast<const CUtensorMap*>(&t.mapBlo);
    launch_k(pca_inverse_t_kernel, dim3(t.args.n_pix / BM), dim3(kThreads), I_SMEM, s, a, bh, bl, t.args);
}

// ------------------------------------------------------------------------------------------------
// Host side: tensor maps (driver entry point fetched through the runtime: no link-time libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int make_kmajor_map(TensorMap128* out, const float* ptr, int rows, int cols, int ld, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return -1;
    static_assert(sizeof(TensorMap128) == sizeof(CUtensorMap), "tensor map size");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -(int)r - 1000;
}

// 5-D maps over the two grid planes: dims (x, bx, y, by, channel), byte strides (4,) 4*stride, 4*W, 4*W*stride, 4*plane_stride --
// overlapping windows: element (x, bx, y, by, c) is pixel (y + by*stride, x + bx*stride) of plane c.  Three boxes over the same
// tensor: gx blocks along x, gy blocks along y, one block; each box row is 32 pixels = 128 B = one swizzle row.
int make_grid_maps(TcGemmGrid* out, const float* planes, int W, int H, long long plane_stride, int stride_px, int gx, int gy) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return -1;
    const int nb = gx > gy ? gx : gy;
    cuuint64_t dims[5] = {(cuuint64_t)W, (cuuint64_t)nb, (cuuint64_t)H, (cuuint64_t)nb, 2};
    cuuint64_t strides[4] = {(cuuint64_t)stride_px * 4, (cuuint64_t)W * 4, (cuuint64_t)W * 4 * stride_px, (cuuint64_t)plane_stride * 4};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const cuuint32_t boxes[3][5] = {{(cuuint32_t)BK, (cuuint32_t)gx, 1, 1, 1}, {(cuuint32_t)BK, 1, 1, (cuuint32_t)gy, 1}, {(cuuint32_t)BK, 1, 1, 1, 1}};
    TensorMap128* maps[3] = {&out->mapRow, &out->mapCol, &out->mapOne};
    for (int i = 0; i < 3; ++i) {
        CUresult r = enc(reinterpret_cast<CUtensorMap*>(maps[i]), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(planes), dims,
                         strides, boxes[i], estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return -(int)r - 1000;
    }
    return 0;
}

int tc_gemm_bn(int N) { return (N % 128 == 0) ? 128 : 64; }

int tc_gemm_prepare() {
    cudaError_t e1 = cudaFuncSetAttribute(tc_gemm_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<128>::SMEM_TOTAL);
    cudaError_t e2 = cudaFuncSetAttribute(tc_gemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<64>::SMEM_TOTAL);
    return (e1 == cudaSuccess && e2 == cudaSuccess) ? 0 : -1;
}

void launch_tc_gemm(const TcGemm& t, cudaStream_t s) {
    const int bn = t.bn;
    dim3 grid(t.args.N / bn, (t.args.M + BM - 1) / BM, t.args.splits);
    const CUtensorMap& a = *reinterpret_cast<const CUtensorMap*>(&t.mapA);
    const CUtensorMap& b = *reinterpret_cast<const CUtensorMap*>(&t.mapB);
    if (bn == 128) launch_k(tc_gemm_kernel<128>, grid, dim3(kThreads), Cfg<128>::SMEM_TOTAL, s, a, b, t.args);
    else launch_k(tc_gemm_kernel<64>, grid, dim3(kThreads), Cfg<64>::SMEM_TOTAL, s, a, b, t.args);
}

}  // namespace psm
