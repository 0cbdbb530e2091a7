// INT8-sliced ("Ozaki scheme") evaluation of the variance triangular product on the 5th-generation tensor cores.
//
// The FP64 DMMA path (trmm_sumsq_kernel) already runs at the FP64 roofline (99 % of the measured cuBLAS DGEMM rate);
// the only way past that roof is to move the contraction onto tcgen05.mma, which has no FP64 kind.  Both operands are
// therefore split -- error-free -- into S signed 7-bit digit planes with a per-row power-of-two scale,
//     x[r][k] = 2^e_r * sum_{t=1..S} digit_t[r][k] * 2^(-7t)  (+ remainder < 2^(e_r - 7S - 1)),   digit_t in [-64, 64],
// the digit planes are multiplied EXACTLY as int8 x int8 -> int32 GEMMs (tcgen05.mma kind::i8, accumulators in TMEM;
// |sum| <= K * 2^12 * S < 2^31 for K <= 32768), and the pair products with a+b <= S+1 are recombined in FP64 in the
// epilogue, which also squares and row-reduces the result (W is never written, as in the DMMA kernel).
// With S = 6 (21 digit-plane products) the predictive std differs from the FP64 path by ~1e-9 of sqrt(c+s2)
// (tolerance 1e-7); S = 7 gives ~1e-11.  Opt-in (gptb_set_variance_mode); the default stays the FP64 DMMA kernel.
//
// Layout: digit planes are int8, row-major, K-contiguous: A planes [S][rows][Npad] (right-hand sides), B planes
// [S][Npad][Npad] (inverse factor, zeros above the diagonal so no masking is needed).  One CTA computes a
// 128 (queries) x 64 (factor rows) tile: TMA (3-D map: k-bytes, rows, plane; 64-byte swizzle) stages all S planes of a
// 64-byte k-chunk per pipeline stage, one thread issues the S(S+1)/2 * 2 MMAs of the chunk, diagonal d = a+b
// accumulates in TMEM columns [64 d, 64 d + 64).
#pragma once
#include "gemm_engine.cuh"

namespace gptb {
namespace oz {

constexpr int OM = 128;      // queries per tile (UMMA M)
constexpr int ON = 64;       // inverse-factor rows per tile (UMMA N)
constexpr int OKB = 64;      // k bytes (= k elements) per pipeline chunk, one 64-byte swizzle row
constexpr int OTHREADS = 192;   // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2..5: epilogue
constexpr int DIGIT_BITS = 7;

template <int S> struct Cfg {
    static constexpr int STAGE_BYTES = S * (OM * OKB + ON * OKB);
    static constexpr int NST = (3 * STAGE_BYTES <= 222 * 1024) ? 3 : 2;
    static constexpr int SMEM_BYTES = NST * STAGE_BYTES + 1024;   // + alignment slack
    static constexpr int TMEM_COLS = 512;                          // S * 64 <= 448 columns used
};

__device__ __forceinline__ uint64_t smem_desc_sw64(const void* p) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(p) & 0x3FFFF) >> 4);   // start address
    d |= (uint64_t)1 << 16;                           // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(512 >> 4) << 32;                  // stride byte offset: 8 rows x 64 B
    d |= (uint64_t)1 << 46;                           // descriptor version (sm_100)
    d |= (uint64_t)4 << 61;                           // SWIZZLE_64B
    return d;
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_3d_u8(void* dst, const CUtensorMap* m, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(smem_u32(dst)),
        "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// ------------------------------------------------------------------------------------------------------------
// Error-free digit split of the rows of a row-major FP64 matrix.  One CTA per row.  lower != 0: only columns
// k <= row are split (the strict upper part -- the mirror -- becomes zero digits).
// scale[row] = 2^e with |x| * 2^-e < 1/2 for the whole row.
// ------------------------------------------------------------------------------------------------------------
template <int S>
__global__ void __launch_bounds__(256) slice_rows_kernel(const double* __restrict__ src, long long ld, long long rows, int ncols,
                                                         int lower, int8_t* __restrict__ planes, long long plane_stride,
                                                         double* __restrict__ scale) {
    const long long row = blockIdx.x;
    const int kmax = lower ? (int)min((long long)ncols, row + 1) : ncols;
    const double* rp = src + row * ld;
    __shared__ double red[256];
    double mx = 0.0;
    for (int k = threadIdx.x; k < kmax; k += 256) mx = fmax(mx, fabs(rp[k]));
    red[threadIdx.x] = mx;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + w]);
        __syncthreads();
    }
    mx = red[0];
    int ex = 0;
    if (mx > 0.0) { frexp(mx, &ex); ex += 1; }          // mx = m * 2^(ex-1), m in [0.5,1)  ->  |x| * 2^-ex < 1/2
    const double down = ldexp(1.0, -ex);
    if (threadIdx.x == 0) scale[row] = ldexp(1.0, ex);
    for (int k = threadIdx.x; k < ncols; k += 256) {
        double y = (k < kmax) ? rp[k] * down : 0.0;
#pragma unroll
        for (int t = 0; t < S; ++t) {
            y *= 128.0;
            const double dgt = rint(y);
            y -= dgt;
            planes[(long long)t * plane_stride + row * (long long)ncols + k] = (int8_t)(int)dgt;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// part[ti64][row] = sum_{i in 64-row tile ti64} ( sum_{k <= i} RHS[row][k] Linv[i][k] )^2   via S(S+1)/2 int8 GEMMs.
// ------------------------------------------------------------------------------------------------------------
template <int S>
__global__ void __launch_bounds__(OTHREADS, 1) ozaki_trmm_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                const __grid_constant__ CUtensorMap mapB,
                                                                const double* __restrict__ scaleA, const double* __restrict__ scaleB,
                                                                int T64, int rowtiles, long long rows_total,
                                                                double* __restrict__ part) {
    using C = Cfg<S>;
    constexpr int NST = C::NST;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full[NST], empty[NST], acc_full;
    __shared__ uint32_t tmem_base_s;
    __shared__ double sB_scale[ON];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // tile order: groups of 12 row tiles x 24 factor-row tiles (same L2 blocking as the DMMA kernel), heaviest first
    int ti, rt;
    {
        constexpr int GR = 12, GI = 24;
        const long long idx = blockIdx.x;
        const long long per_tib = (long long)rowtiles * GI;
        const int tib = (int)(idx / per_tib);
        const long long rem = idx - (long long)tib * per_tib;
        const int ti_cnt = min(GI, T64 - tib * GI);
        const int rb = (int)(rem / ((long long)GR * ti_cnt));
        const int rem2 = (int)(rem - (long long)rb * GR * ti_cnt);
        rt = rb * GR + rem2 / ti_cnt;
        ti = T64 - 1 - (tib * GI + rem2 % ti_cnt);
    }
    const int nchunk = ti + 1;      // k extent (ti+1)*64 bytes, 64 per chunk

    if (tid == 0) {
        for (int i = 0; i < NST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(&acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (tid < ON) sB_scale[tid] = scaleB[(long long)ti * ON + tid];
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {
            for (int c = 0; c < nchunk; ++c) {
                const int st = c % NST;
                if (c >= NST) mbar_wait(&empty[st], ((c / NST) - 1) & 1);
                uint8_t* sA = smem + st * C::STAGE_BYTES;
                uint8_t* sB = sA + S * OM * OKB;
                mbar_expect_tx(&full[st], C::STAGE_BYTES);
                tma_load_3d_u8(sA, &mapA, c * OKB, rt * OM, 0, &full[st]);
                tma_load_3d_u8(sB, &mapB, c * OKB, ti * ON, 0, &full[st]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: D = S32, A = B = INT8, both K-major, N = 64, M = 128
            const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(ON >> 3) << 17) | ((uint32_t)(OM >> 4) << 24);
            for (int c = 0; c < nchunk; ++c) {
                const int st = c % NST;
                mbar_wait(&full[st], (c / NST) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint8_t* sA = smem + st * C::STAGE_BYTES;
                const uint8_t* sB = sA + S * OM * OKB;
#pragma unroll
                for (int a = 0; a < S; ++a) {
                    const uint64_t ad = smem_desc_sw64(sA + a * OM * OKB);
#pragma unroll
                    for (int b = 0; b < S - a; ++b) {
                        const uint64_t bd = smem_desc_sw64(sB + b * ON * OKB);
                        const uint32_t dcol = tmem_base + (uint32_t)((a + b) * ON);
#pragma unroll
                        for (int kk = 0; kk < OKB / 32; ++kk)
                            umma_i8(dcol, ad + (uint64_t)(kk * 2), bd + (uint64_t)(kk * 2), idesc, (a == 0 && c == 0 && kk == 0) ? 0u : 1u);
                    }
                }
                umma_commit(&empty[st]);        // the stage is free once these MMAs have read it
            }
            umma_commit(&acc_full);
        }
    } else {
        // ---------------- epilogue warps 2..5: TMEM lane quarter = warp % 4 ----------------
        const int quarter = warp & 3;
        const int q = quarter * 32 + lane;                      // query row within the tile = TMEM lane
        mbar_wait(&acc_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        double ssum = 0.0;
#pragma unroll 1
        for (int cb = 0; cb < ON / 16; ++cb) {
            double v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0.0;
#pragma unroll
            for (int d = 0; d < S; ++d) {
                uint32_t r[16];
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(d * ON + cb * 16);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
                const double w = ldexp(1.0, -DIGIT_BITS * (d + 2));       // digits are 1-based: weight 2^-7(a+b), a+b = d+2
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = fma((double)(int)r[j], w, v[j]);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const double vv = v[j] * sB_scale[cb * 16 + j];
                ssum = fma(vv, vv, ssum);
            }
        }
        const long long grow = (long long)rt * OM + q;
        const double sa = scaleA[grow];
        part[(long long)ti * rows_total + grow] = ssum * sa * sa;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
}

}  // namespace oz
}  // namespace gptb
