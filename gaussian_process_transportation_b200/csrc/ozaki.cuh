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
#include "digits.cuh"

namespace gptb {
namespace oz {

constexpr int OM = 128;      // queries per tile (UMMA M)
constexpr int ON = 64;       // inverse-factor rows per tile (UMMA N)
constexpr int OKB = 64;      // k bytes (= k elements) per pipeline chunk, one 64-byte swizzle row
constexpr int OTHREADS = 256;   // dense variant -- warp group 0: warp 0 = TMA producer, warp 1 = MMA issuer (owns TMEM); warp group 1: epilogue
// skipping variant: 384 threads -- warp 0 producer, warps 1..3 MMA issuers, warps 4..7 and 8..11 epilogue (two warps per TMEM lane quarter,
// 32 of the 64 accumulator columns each: the issuing warps wait for the epilogue of the previous tile 8-10 % of a launch at N = 4096)
constexpr int OTHREADS_SKIP = 384;
// Register budget (setmaxnreg): the control warp group gives registers back (down to 80 per thread), the epilogue warp groups grow to
// 176 (ptxas compiles each branch against its own budget, no spills).  Dense variant: __launch_bounds__(256, 2) = 128 registers at
// launch, a footprint that leaves room for one 256-thread generator CTA on the SM (the generator/product overlap of round 1,
// gptb_set_query_pipeline: kernels overlap, no net gain).  Skipping variant: __launch_bounds__(384, 1) = 168 registers at launch, 55 K of
// the SM's 64 K after the exchange; it owns the SM (216 KB of shared memory), so the overlap option only double-buffers there.
constexpr int OREG_LIGHT = 80, OREG_EPI = 176;
// digit width: 7 (balanced digits, "int8xS") or 8 (full int8 range, "int8wS"; digits.cuh) -- a template parameter of the
// slicers and a run-time argument of the product kernel (it only changes the recombination weights)

template <int S> struct Cfg {
    static constexpr int STAGE_BYTES = S * (OM * OKB + ON * OKB);
    static constexpr int NST = (3 * STAGE_BYTES <= 222 * 1024) ? 3 : 2;
    static constexpr int SMEM_BYTES = NST * STAGE_BYTES + 1024;   // + alignment slack
    static constexpr int TMEM_COLS = 512;                          // S * 64 <= 448 columns used
    // skipping variant: plane-granular operand ring (one slot = one A plane 128 x 64 B + one B plane 64 x 64 B)
    static constexpr int SLOT_A = OM * OKB, SLOT_B = ON * OKB;
    static constexpr int NSLOT = (216 * 1024) / (SLOT_A + SLOT_B);                 // 18
    static constexpr int NE = 12;                                                  // chunks in flight (barrier ring), a multiple of the 3 issuing warps
    static constexpr int SMEM_SKIP_BYTES = NSLOT * (SLOT_A + SLOT_B) + 1024;
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
// One elected lane of a CONVERGED warp (the same lane every time for a full mask).  The issuing warps below stay converged and
// make their control values warp-uniform with a redux (its result lives in a uniform register), so that shared-memory
// addresses, descriptors and TMA coordinates are computed on the uniform datapath and UTCIMMA / UTMALDG take them directly.
// With `if (lane == 0)` around the whole loop the compiler could not prove uniformity and wrapped EVERY tcgen05.mma and TMA
// issue in an ELECT / R2UR.BROADCAST / BRA waterfall loop (~12 instructions per MMA by one thread).
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ int warp_uniform(int v) { return (int)__reduce_max_sync(0xffffffffu, v); }

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// Address-based twins for the issuing warps of the skipping variant.  A shared-memory address on sm_100 carries the CTA's rank in
// its upper bits, and the compiler re-derived it (S2R SR_CgaCtaId, a ~50-cycle special-register read) in front of every
// tcgen05.commit / TMA issue inside the chunk loop -- 15 % of the MMA warp's samples sat on the R2UR that waits for it.  Here the
// base addresses are taken once per role (through an opaque mov, so they cannot be rematerialised, and a redux, so they stay on
// the uniform datapath) and everything else is integer offsets.
__device__ __forceinline__ uint32_t pinned_uniform_addr(const void* p) {
    uint32_t a;
    asm volatile("mov.u32 %0, %1;\n" : "=r"(a) : "r"(smem_u32(p)));
    return (uint32_t)__reduce_max_sync(0xffffffffu, a);
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void umma_commit_a(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_3d_u8_a(uint32_t dst, const CUtensorMap* m, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(dst),
        "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ uint64_t smem_desc_sw64_a(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);          // start address
    d |= (uint64_t)1 << 16;                           // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(512 >> 4) << 32;                  // stride byte offset: 8 rows x 64 B
    d |= (uint64_t)1 << 46;                           // descriptor version (sm_100)
    d |= (uint64_t)4 << 61;                           // SWIZZLE_64B
    return d;
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
template <int S, int BITS>
__global__ void __launch_bounds__(256) slice_rows_kernel(const double* __restrict__ src, long long ld, long long rows, int ncols,
                                                         int lower, int8_t* __restrict__ planes, long long plane_stride,
                                                         double* __restrict__ scale, unsigned long long* __restrict__ l1max = nullptr,
                                                         unsigned* __restrict__ flags = nullptr, int flags_stride = 0) {
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
    const int ex = digit_scale_exp(mx, BITS);           // |x| * 2^-ex fits the first digit for the whole row
    const double down = ldexp(1.0, -ex);
    const double sc8 = digit_scale8(mx);                // 8-bit planes: tightest scale instead of the next power of two
    if (threadIdx.x == 0) scale[row] = (BITS == 8) ? sc8 : ldexp(1.0, ex);
    if constexpr (BITS == 8) {
        // four columns per thread, one 32-bit store per plane (ncols is a multiple of 128)
        const double mul = ldexp(1.0, 8 * S) / sc8;
        unsigned long long l1 = 0;
        for (int k = 4 * threadIdx.x; k < ncols; k += 1024) {
            double v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = (k + j < kmax) ? rp[k + j] : 0.0;
            unsigned packed[S];
            digits8_pack4<S>(v, mul, packed);
#pragma unroll
            for (int t = 0; t < S; ++t) {
                *reinterpret_cast<unsigned*>(planes + (long long)t * plane_stride + row * (long long)ncols + k) = packed[t];
                l1 += __vsadu4(__vabsss4(packed[t]), 0u);          // sum of |digit| over the four bytes (|-128| saturates to 127: +1 each below)
                l1 += __popc(__vcmpeq4(packed[t], 0x80808080u) & 0x01010101u);
            }
            if (flags != nullptr) {
                // non-zero planes of the (64-row, 64-column) block this thread's four digits belong to (see DigitScales::flags)
                unsigned m = 0;
#pragma unroll
                for (int t = 0; t < S; ++t) m |= (packed[t] != 0u ? 1u : 0u) << t;
                m |= __shfl_xor_sync(0xffffffffu, m, 1);
                m |= __shfl_xor_sync(0xffffffffu, m, 2);
                m |= __shfl_xor_sync(0xffffffffu, m, 4);
                m |= __shfl_xor_sync(0xffffffffu, m, 8);
                const int chunk = k >> 6;
                if ((threadIdx.x & 15) == 0 && m != 0u) atomicOr(flags + (row >> 6) * flags_stride + (chunk >> 2), m << ((chunk & 3) * 8));
            }
        }
        if (l1max != nullptr) {
            // data-dependent exactness bound of the int32 accumulators: |sum_k a_k b_k| <= 128 * sum_k |b_k|, summed over this row's planes
            __syncthreads();
            unsigned long long* redl = reinterpret_cast<unsigned long long*>(red);
            redl[threadIdx.x] = l1;
            __syncthreads();
            for (int w = 128; w > 0; w >>= 1) {
                if (threadIdx.x < w) redl[threadIdx.x] += redl[threadIdx.x + w];
                __syncthreads();
            }
            if (threadIdx.x == 0) atomicMax(l1max, redl[0]);
        }
    } else {
        for (int k = threadIdx.x; k < ncols; k += 256) {
            double y = (k < kmax) ? rp[k] * down : 0.0;
#pragma unroll
            for (int t = 0; t < S; ++t) {
                const double dgt = digit_step<BITS>(y);
                planes[(long long)t * plane_stride + row * (long long)ncols + k] = (int8_t)(int)dgt;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// part[ti64][row] = sum_{i in 64-row tile ti64} ( sum_{k <= i} RHS[row][k] Linv[i][k] )^2   via S(S+1)/2 int8 GEMMs.
// Persistent: grid = #SMs, CTA b runs tiles b, b + grid, ... of the L2-blocked, heaviest-first tile order; TMEM is
// allocated once, the TMA producer streams straight into the next tile while the epilogue drains the accumulators
// (the MMA thread only waits for that drain, ~1 us per ~20 us tile).
// ------------------------------------------------------------------------------------------------------------
// TMA views of one operand's digit planes with 1 .. 7 planes per box: m[n-1] loads n consecutive planes starting at the
// plane coordinate, so a chunk whose leading planes are zero moves (and pays L2 bandwidth for) only the planes it needs.
struct PlaneMaps {
    CUtensorMap m[7];
};

// (Tried and backed out in round 2: the producer publishing the decoded codes in shared memory so that the issuing warps need not load
// and decode the masks themselves -- the slot then stays busy until the issuing warps finish the tile, which throttles the producer's
// run-ahead: c3 step 78.8 -> 79.7 ms, N = 16384 208.7 -> 220.0 ms; profiles/r02_pipeline_ab_v_published_codes_rejected.log.)
// Block masks of one tile, skipping variant.  Each control warp loads its two mask rows ONCE per tile (one 16-byte load per lane and
// operand: 16 mask bytes = 16 chunks), decodes them lane-parallel into (za*8 + zb) codes and then hands them out four chunks at a
// time with one shuffle -- the earlier reader (ZeroPlaneReader) fetched one 32-bit word per four chunks with __ldg and decoded it
// serially; what-if timings (tools/whatif.py, profiles/r02_whatif_*.log) showed the launch bound by the control warps' instruction
// streams (a lone warp retires ~0.2 instructions per cycle: 100+ instructions per chunk cost more than the chunk's MMAs).
struct TileMasks {
    unsigned code[4];              // this lane's codes for chunks 16*lane + 4*w .. + 3 (one byte each), w = 0..3
    unsigned cur4;                 // codes of the current group of four chunks (warp-uniform)
    template <int S>
    __device__ __forceinline__ void load(const unsigned* flagsA, const unsigned* flagsB, int stride, int rt, int ti, int lane) {
        uint4 wa = make_uint4(0u, 0u, 0u, 0u), wb = wa;
        if (flagsA != nullptr && 4 * lane < stride) {
            wa = __ldg(reinterpret_cast<const uint4*>(flagsA + (long long)rt * stride) + lane);
            wb = __ldg(reinterpret_cast<const uint4*>(flagsB + (long long)ti * stride) + lane);
        }
        const unsigned a[4] = {wa.x, wa.y, wa.z, wa.w}, b[4] = {wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            unsigned out = 0u;
            if (flagsA != nullptr) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const unsigned ma = (a[w] >> (8 * k)) & 0xffu, mb = (b[w] >> (8 * k)) & 0xffu;
                    const unsigned za = ma ? (unsigned)(__ffs(ma) - 1) : (unsigned)S, zb = mb ? (unsigned)(__ffs(mb) - 1) : (unsigned)S;
                    out |= (za * 8u + zb) << (8 * k);
                }
            }
            code[w] = out;
        }
        if (lane == 0) code[0] &= 0xffffff00u;                   // chunk 0 is always dense: it zero-initialises the accumulators
        cur4 = 0u;
    }
    // call for every chunk in order
    __device__ __forceinline__ void get(int c, int& za, int& zb) {
        if ((c & 3) == 0) {
            const int w = (c >> 2) & 3;
            const unsigned mine = (w == 0) ? code[0] : (w == 1) ? code[1] : (w == 2) ? code[2] : code[3];
            cur4 = __shfl_sync(0xffffffffu, mine, c >> 4);
        }
        const unsigned z = (cur4 >> ((c & 3) * 8)) & 0xffu;
        za = (int)(z >> 3);
        zb = (int)(z & 7u);
    }
};

// ------------------------------------------------------------------------------------------------------------
// MMA issue for one 64-byte k-chunk with the first ZA planes of A and the first ZB planes of B known to be all zero
// (ZA = ZB = 0: the dense case).  Everything but the two base descriptors is a compile-time constant, so a wide MMA costs
// the issuing thread a handful of instructions; a run-time (za, zb) version of this loop (descriptor arithmetic, loop
// control and predicates recomputed per MMA by ONE thread) was issue-bound: 7.3 instead of 5.5 ms per launch.
// For a fixed A plane a the partner planes b = ZB .. S-1-a accumulate into the diagonals a+b, i.e. into CONTIGUOUS TMEM
// columns [64(a+ZB), 64 S), and those B planes are contiguous rows in shared memory: the pair products are issued as ONE
// wide MMA (N = 64 (S-a-ZB); above 256 as two EQUAL halves, e.g. 320 = 160 + 160 rather than 256 + 64 -- a 64-wide MMA
// re-reads the 4 KB A tile for 32 cycles of work and is operand-port bound).  With per-pair N = 64 MMAs the operand fetch
// (6 KB per 32-cycle MMA = 192 B/clk) exceeded the 128 B/clk shared-memory port and capped the kernel at 2/3 of the
// tensor rate (ncu: sm__throughput 88 %, tensor pipe 42 %).
// ------------------------------------------------------------------------------------------------------------
// X = 1 adds the first DROPPED diagonal a + b = S (0-based planes) to the scheme: S + X diagonals, S(S+1)/2 + (S-1) products.  The CPU
// study (tools/plane_error_study.py, profiles/r02_plane_error_study.log) shows that with 8-bit planes the std error of the S = 5
// scheme is NOT the 40-bit operand truncation (5.5e-10 at N = 4096) but the dropped diagonal (9.0e-9); its four products recover
// the truncation level for 27 % more MMA work and no extra operand traffic, where a sixth plane costs 40 % and 20 %.
template <int S, int ZA, int ZB, int X = 0>
__device__ __forceinline__ void oz_issue_chunk(uint64_t adesc0, uint64_t bdesc0, uint32_t tmem_base, uint32_t idesc_base, uint32_t first) {
    constexpr int DMAX = S - 1 + X;                                     // last diagonal kept
    // B planes blo .. bhi (0-based plane numbers, >= ZB) against A plane a; init: the first MMA overwrites instead of accumulating
    auto issue = [&](int a, int blo, int bhi, bool init) {
        const int ncols = ON * (bhi - blo + 1);
        const int nhalf = (ncols > 256) ? 2 : 1;
        const int nw = ncols / nhalf;                                   // multiple of 32
        const uint64_t ad = adesc0 + (uint64_t)(((a - ZA) * OM * OKB) >> 4);      // the stage holds A planes ZA.. and B planes ZB.. from slot 0
#pragma unroll
        for (int hf = 0; hf < nhalf; ++hf) {
            const uint64_t bd = bdesc0 + (uint64_t)((((blo - ZB) * ON + hf * nw) * OKB) >> 4);
            const uint32_t dcol = tmem_base + (uint32_t)((a + blo) * ON + hf * nw);
            const uint32_t idn = idesc_base | ((uint32_t)(nw >> 3) << 17);
#pragma unroll
            for (int kk = 0; kk < OKB / 32; ++kk)
                umma_i8(dcol, ad + (uint64_t)(kk * 2), bd + (uint64_t)(kk * 2), idn, (init && kk == 0) ? 0u : 1u);
        }
    };
#pragma unroll
    for (int a = ZA; a < S; ++a) {
        const int bhi = (DMAX - a < S - 1) ? DMAX - a : S - 1;          // partner planes ZB .. bhi
        if (bhi < ZB) continue;
        if (first && X == 1 && a == 1) {
            // chunk 0 zero-initialises the accumulators: plane 0 covers diagonals 0 .. S-1, the extra diagonal S is first touched here
            issue(a, ZB, bhi - 1, false);
            issue(a, bhi, bhi, true);
        } else {
            issue(a, ZB, bhi, first && a == 0);
        }
    }
}
// (za, zb) -> specialised issue code through one indexed branch
template <int S, int X = 0>
__device__ __forceinline__ void oz_dispatch(int za, int zb, uint64_t adesc0, uint64_t bdesc0, uint32_t tmem_base, uint32_t idesc_base) {
#define OZ_CASE(ZA, ZB)                                                                                     \
    case (ZA) * 8 + (ZB):                                                                                   \
        if constexpr ((ZA) + (ZB) <= S - 1 + X && (ZA) < S && (ZB) < S) oz_issue_chunk<S, (ZA), (ZB), X>(adesc0, bdesc0, tmem_base, idesc_base, 0u); \
        break;
#define OZ_ROW(ZA) OZ_CASE(ZA, 0) OZ_CASE(ZA, 1) OZ_CASE(ZA, 2) OZ_CASE(ZA, 3) OZ_CASE(ZA, 4) OZ_CASE(ZA, 5) OZ_CASE(ZA, 6)
    switch (za * 8 + zb) {
        OZ_ROW(0) OZ_ROW(1) OZ_ROW(2) OZ_ROW(3) OZ_ROW(4) OZ_ROW(5) OZ_ROW(6)
        default: break;
    }
#undef OZ_ROW
#undef OZ_CASE
}
// planes of A / of B a chunk with za / zb leading zero planes needs, and the number of plane-pair products it carries (0: skip it)
template <int S, int X>
__device__ __forceinline__ void oz_chunk_planes(int za, int zb, int& nA, int& nB, int& npairs) {
    constexpr int DMAX = S - 1 + X;
    const int ahi = min(S - 1, DMAX - zb), bhi = min(S - 1, DMAX - za);
    nA = ahi - za + 1;
    nB = bhi - zb + 1;
    npairs = 0;
    if (nA <= 0 || nB <= 0) { nA = nB = 0; return; }
    for (int a = za; a <= ahi; ++a) npairs += min(S - 1, DMAX - a) - zb + 1;
}

__device__ __forceinline__ void oz_tile_decode(long long idx, int T64, int rowtiles, int& rt, int& ti) {
    constexpr int GR = 12, GI = 24;     // groups of 12 row tiles x 24 factor-row tiles (same span as the DMMA kernel's 12 x 12)
    const long long per_tib = (long long)rowtiles * GI;
    const int tib = (int)(idx / per_tib);
    const long long rem = idx - (long long)tib * per_tib;
    const int ti_cnt = min(GI, T64 - tib * GI);
    const int rb = (int)(rem / ((long long)GR * ti_cnt));
    const int rem2 = (int)(rem - (long long)rb * GR * ti_cnt);
    rt = rb * GR + rem2 / ti_cnt;
    ti = T64 - 1 - (tib * GI + rem2 % ti_cnt);
}

// SKIP = false: every plane product of every chunk (natural order, or spatial mode 2); SKIP = true: block masks consulted.
// The two variants keep separate producer / MMA loops on purpose.  Same-box A/B (profiles/r01_dense_vs_skip_loops.log): the
// skipping loops run the DENSE case 20 % slower than the original loops (N = 16384: 102 vs 85 ms per launch) although they
// issue the same TMA loads and the same MMAs -- and neither the templated issue code, nor the run-time tensor-map index,
// nor the converged-warp issue (which does remove the ELECT / R2UR waterfall around every UTCIMMA / UTMALDG) accounts for it.
// Unexplained at the end of round 1; the dense path therefore stays on the loops it was tuned with.
template <int S, bool SKIP, int X = 0>
__global__ void __launch_bounds__(SKIP ? OTHREADS_SKIP : OTHREADS, SKIP ? 1 : 2) ozaki_trmm_kernel(const __grid_constant__ PlaneMaps mapsA,
                                                                const __grid_constant__ PlaneMaps mapsB,
                                                                const double* __restrict__ scaleA, const double* __restrict__ scaleB,
                                                                int T64, int rowtiles, long long rows_total,
                                                                double* __restrict__ part, int* __restrict__ tile_counter, int digit_bits,
                                                                const unsigned* __restrict__ flagsA, const unsigned* __restrict__ flagsB,
                                                                int flags_stride, unsigned long long* __restrict__ exec_pairs, int whatif,
                                                                unsigned long long* __restrict__ prof) {
    (void)prof;                                               // per-role cycle counters: only the six-stage experiment of round 2 filled them (profiles/r02_role_cycles_*.log)
    // exec_pairs (may be null): number of (plane pair, 64-byte chunk) products this launch really issued, one atomicAdd per tile --
    // the executed-work figure of the roofline (the dense variant issues S(S+1)/2 per chunk by construction, counted on the host).
    // whatif: developer builds only (-DGPTB_OZ_WHATIF, tools/whatif.py): bit 0 = no TMA loads, bit 1 = no MMA issue, bit 2 = no
    // epilogue arithmetic -- which resource bounds the launch; compiled out of the product library.
#ifndef GPTB_OZ_WHATIF
    (void)whatif;
    constexpr int wi = 0;
#else
    const int wi = whatif;
#endif
    using C = Cfg<S>;
    constexpr int NST = C::NST;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int NBAR = SKIP ? C::NE : NST;
    __shared__ __align__(8) uint64_t full[NBAR], empty[NBAR], acc_full, acc_empty, slot_full[2], slot_empty[2], tile_go;
    constexpr int NMMA = SKIP ? 3 : 1;     // issuing warps (skipping variant: warps 1..3, chunk entry e belongs to warp 1 + e % 3)
    __shared__ uint32_t tmem_base_s;
    __shared__ int tile_slot[2];           // dynamic tile scheduler: the producer claims tiles, the other roles follow
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long ntiles = (long long)rowtiles * T64;

    if (tid == 0) {
        for (int i = 0; i < NBAR; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(&acc_full, NMMA);        // one tcgen05.commit per issuing warp
        mbar_init(&acc_empty, SKIP ? 8 : 4);   // one arrival per epilogue warp
        mbar_init(&tile_go, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&slot_full[i], 1); mbar_init(&slot_empty[i], NMMA + (SKIP ? 8 : 4)); }   // issuing warps + epilogue warps
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(OREG_LIGHT));
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(OREG_EPI));

    if (!SKIP && warp == 0) {
        // dense variant (no block masks): the single-lane loops of the kernel before the spatial mode, kept verbatim -- see the
        // note at the template head
        if (lane == 0) {
            int gs = 0;                                          // chunks issued so far (ring position)
            for (int lt = 0;; ++lt) {
                // claim the next tile in the global L2-blocked order (all SMs stay inside one window of ~#SM tiles, which is
                // what keeps the operand slabs L2-resident; a static stride let the CTAs drift apart and cost 25 %)
                if (lt >= 2) mbar_wait(&slot_empty[lt & 1], ((lt >> 1) - 1) & 1);
                int t = atomicAdd(tile_counter, 1);
                if (t >= ntiles) t = -1;
                tile_slot[lt & 1] = t;
                mbar_arrive(&slot_full[lt & 1]);
                if (t < 0) break;
                int rt, ti;
                oz_tile_decode(t, T64, rowtiles, rt, ti);
                for (int c = 0; c <= ti; ++c, ++gs) {
                    const int st = gs % NST;
                    if (gs >= NST) mbar_wait(&empty[st], ((gs / NST) - 1) & 1);
                    uint8_t* sA = smem + st * C::STAGE_BYTES;
                    uint8_t* sB = sA + S * OM * OKB;
                    mbar_expect_tx(&full[st], C::STAGE_BYTES);
                    tma_load_3d_u8(sA, &mapsA.m[S - 1], c * OKB, rt * OM, 0, &full[st]);
                    tma_load_3d_u8(sB, &mapsB.m[S - 1], c * OKB, ti * ON, 0, &full[st]);
                }
            }
        }
    } else if (!SKIP && warp == 1) {
        if (lane == 0) {
            // instruction descriptor: D = S32, A = B = INT8, both K-major, M = 128
            const uint32_t idesc_base = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OM >> 4) << 24);   // N is or-ed in per MMA
            int gs = 0;
            for (int lt = 0;; ++lt) {
                mbar_wait(&slot_full[lt & 1], (lt >> 1) & 1);
                const int t = tile_slot[lt & 1];
                mbar_arrive(&slot_empty[lt & 1]);
                if (t < 0) break;
                int rt, ti;
                oz_tile_decode(t, T64, rowtiles, rt, ti);
                if (lt > 0) {                                    // accumulators of the previous tile must be drained
                    mbar_wait(&acc_empty, (lt - 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                }
                for (int c = 0; c <= ti; ++c, ++gs) {
                    const int st = gs % NST;
                    mbar_wait(&full[st], (gs / NST) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                    const uint8_t* sA = smem + st * C::STAGE_BYTES;
                    const uint8_t* sB = sA + S * OM * OKB;
                    // For a fixed A plane a the partner planes b = 0 .. S-1-a accumulate into the diagonals a+b = a .. S-1,
                    // i.e. into CONTIGUOUS TMEM columns [64a, 64S); and those B planes are contiguous rows in shared memory.
                    // So the S-a pair products are issued as ONE wide MMA (N = 64(S-a), split at 256): the A tile is read
                    // from shared memory once per wide MMA instead of once per pair.  With per-pair N = 64 MMAs the operand
                    // fetch (6 KB per 32-cycle MMA = 192 B/clk) exceeded the 128 B/clk shared-memory port and capped the
                    // kernel at 2/3 of the tensor rate (ncu: sm__throughput 88 %, tensor pipe 42 %).
#pragma unroll
                    for (int a = 0; a < S; ++a) {
                        const uint64_t ad = smem_desc_sw64(sA + a * OM * OKB);
                        // N = 64 (S-a) columns; above 256 the product is issued as two EQUAL halves (e.g. 320 = 160 + 160, not
                        // 256 + 64: a 64-wide MMA re-reads the 4 KB A tile for 32 cycles of work and is operand-port bound)
                        const int ncols = ON * (S - a);
                        const int nhalf = (ncols > 256) ? 2 : 1;
                        const int nw = ncols / nhalf;                                   // multiple of 32
#pragma unroll
                        for (int hf = 0; hf < nhalf; ++hf) {
                            const uint64_t bd = smem_desc_sw64(sB + hf * nw * OKB);    // nw rows further down the stacked B planes
                            const uint32_t dcol = tmem_base + (uint32_t)(a * ON + hf * nw);
                            const uint32_t idn = idesc_base | ((uint32_t)(nw >> 3) << 17);
#pragma unroll
                            for (int kk = 0; kk < OKB / 32; ++kk)
                                umma_i8(dcol, ad + (uint64_t)(kk * 2), bd + (uint64_t)(kk * 2), idn, (a == 0 && c == 0 && kk == 0) ? 0u : 1u);
                        }
                    }
                    umma_commit(&empty[st]);        // the stage is free once these MMAs have read it
                }
                umma_commit(&acc_full);
            }
        }
    } else if (SKIP && warp == 0) {
        // ---------------- TMA producer: converged warp, one elected lane issues (see elect_one_sync) ----------------
        // Operand ring of NSLOT plane slots (slot = one A plane + one B plane of a 64-byte k-chunk): a chunk with n live planes takes n
        // CONSECUTIVE slots (the wide MMAs need its B planes stacked, the TMA box its A planes), so up to NE chunks are in flight
        // instead of three fixed 5-plane stages.  Measured with the what-if build (profiles/r02_whatif_b_*.log): with three stages
        // the launch was bound by the per-stage round trip (empty -> TMA issue -> TMA latency -> full -> MMA issue -> MMAs ->
        // commit, ~2000 cycles + the chunk's MMAs) divided by 3, far above the ~350 cycles of MMA work an average chunk carries.
        // Slots are released in FIFO order: the producer retires the oldest chunk (waits for its empty barrier) until the next
        // allocation fits.  The issuing warps replay the same allocation arithmetic, so no addresses travel between the roles.
        constexpr int NE = C::NE, NSLOT = C::NSLOT;
        int head = 0, free_slots = NSLOT;                    // next slot to hand out / slots not held by a chunk in flight
        int ej = 0, tj = 0, inflight = 0;                    // barrier index of the next chunk / of the oldest chunk in flight
        unsigned tpar = 0u;                                  // parity of empty[tj] that retires the oldest chunk
        unsigned long long used = 0ull;                      // 4 bits per barrier index: slots held by that chunk (incl. wrap padding)
        const uint32_t full0 = pinned_uniform_addr(&full[0]), empty0 = pinned_uniform_addr(&empty[0]);
        const uint32_t smemA = pinned_uniform_addr(smem), smemB = smemA + NSLOT * C::SLOT_A;
        for (int lt = 0;; ++lt) {
            // claim the next tile in the global L2-blocked order (all SMs stay inside one window of ~#SM tiles, which is
            // what keeps the operand slabs L2-resident; a static stride let the CTAs drift apart and cost 25 %)
            if (lt >= 2) mbar_wait(&slot_empty[lt & 1], ((lt >> 1) - 1) & 1);
            int t = 0;
            if (lane == 0) {
                t = atomicAdd(tile_counter, 1);
                if (t >= ntiles) t = -1;
                tile_slot[lt & 1] = t;
                mbar_arrive(&slot_full[lt & 1]);
            }
            t = warp_uniform(lane == 0 ? t : (int)0x80000000);
            if (t < 0) break;
            int rt, ti;
            oz_tile_decode(t, T64, rowtiles, rt, ti);
            TileMasks zr;
            zr.template load<S>(flagsA, flagsB, flags_stride, rt, ti, lane);
            for (int c = 0; c <= ti; ++c) {
                int za, zb;
                zr.get(c, za, zb);
                int nA, nB, npr;
                oz_chunk_planes<S, X>(za, zb, nA, nB, npr);  // planes za .. za+nA-1 of A meet planes zb .. zb+nB-1 of B
                if (npr == 0) continue;                     // nothing but zeros in this chunk: no slot, no load
                const int n = max(nA, nB);                  // slots: one A plane + one B plane each
                int base = head, need = n;
                if (head + n > NSLOT) { need += NSLOT - head; base = 0; }      // no wrap inside a chunk: the tail slots ride along as padding
                while (free_slots < need || inflight == NE) {                  // retire the oldest chunk(s)
                    mbar_wait_a(empty0 + 8u * tj, tpar);
                    free_slots += (int)((used >> (4 * tj)) & 15ull);
                    --inflight;
                    if (++tj == NE) { tj = 0; tpar ^= 1u; }
                }
                free_slots -= need;
                head = base + n;
                used = (used & ~(15ull << (4 * ej))) | ((unsigned long long)need << (4 * ej));
                const uint32_t fbar = full0 + 8u * ej;
                if (elect_one_sync()) {
                    if (wi & 1) {
                        mbar_expect_tx_a(fbar, 0);
                    } else {
                        mbar_expect_tx_a(fbar, (nA * OM + nB * ON) * OKB);
                        tma_load_3d_u8_a(smemA + base * C::SLOT_A, &mapsA.m[nA - 1], c * OKB, rt * OM, za, fbar);
                        tma_load_3d_u8_a(smemB + base * C::SLOT_B, &mapsB.m[nB - 1], c * OKB, ti * ON, zb, fbar);
                    }
                }
                ++inflight;
                if (++ej == NE) ej = 0;
            }
        }
    } else if (SKIP && warp >= 1 && warp <= 3) {
        // ---------------- MMA issue: three warps, chunk entry e belongs to warp 1 + e % 3 (converged warps, one elected lane issues) ----
        // Measured (tools/whatif.py, profiles/r02_whatif_a_*.log): with ONE issuing warp the launch time was the SUM of the control path
        // (barrier wait, mask decode, descriptor arithmetic, indexed branch: ~670 cycles per chunk) and the MMA time -- the tensor
        // pipe's instruction queue is shallow, the issuing warp blocks in UTCIMMA until earlier MMAs have drained, and its next
        // control path then runs with the pipe idle.  With three issuing warps the control path of chunks e+1, e+2 runs under the
        // MMAs of chunk e.  NE is a multiple of 3, so barrier j always belongs to warp 1 + j % 3 and every warp sees every phase of
        // its own barriers.  int32 accumulation is exact, so the order in which the warps' MMAs reach the pipe does not matter --
        // except that the chunk-0 products (accumulate = 0) must be issued first: the owner of a tile's first chunk signals tile_go
        // after issuing them, the others wait for it before their first issue of the tile.
        // instruction descriptor: D = S32, A = B = INT8, both K-major, M = 128
        constexpr int NE = C::NE, NSLOT = C::NSLOT;
        const uint32_t idesc_base = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OM >> 4) << 24);   // N is or-ed in per MMA
        const int my = warp - 1;
        int head = 0, ej = 0, e3 = 0;                        // the producer's allocation replayed: next slot, barrier index, e % 3
        unsigned epar = 0u;                                  // parity of full[ej] for the current lap of the barrier ring
        const uint32_t full0 = pinned_uniform_addr(&full[0]), empty0 = pinned_uniform_addr(&empty[0]);
        const uint32_t smemA = pinned_uniform_addr(smem), smemB = smemA + NSLOT * C::SLOT_A;
        const uint64_t adescA = smem_desc_sw64_a(smemA), bdescB = smem_desc_sw64_a(smemB);
        const uint32_t accf = pinned_uniform_addr(&acc_full), tgo = pinned_uniform_addr(&tile_go);
        for (int lt = 0;; ++lt) {
            mbar_wait(&slot_full[lt & 1], (lt >> 1) & 1);
            const int t = warp_uniform((int)tile_slot[lt & 1]);
            __syncwarp();
            if (lane == 0) mbar_arrive(&slot_empty[lt & 1]);
            if (t < 0) break;
            int rt, ti;
            oz_tile_decode(t, T64, rowtiles, rt, ti);
            if (lt > 0) {                                    // accumulators of the previous tile must be drained
                mbar_wait(&acc_empty, (lt - 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            }
            bool go_seen = (e3 == my);                       // chunk 0 of this tile is mine: I open the tile
            // spatial mode: per (row tile, chunk) / (factor-row tile, chunk) masks of the non-zero digit planes.  Leading zero
            // planes (small values: far-away training points, far-off-diagonal entries of L^-1) are neither loaded nor
            // multiplied -- exact, the skipped products are sums of zeros.
            TileMasks zr;
            zr.template load<S>(flagsA, flagsB, flags_stride, rt, ti, lane);
            unsigned npairs = 0;                             // plane-pair products issued by this warp for this tile
            for (int c = 0; c <= ti; ++c) {
                int za, zb;
                zr.get(c, za, zb);
                int nA, nB, npr;
                oz_chunk_planes<S, X>(za, zb, nA, nB, npr);
                if (npr == 0) continue;
                const int n = max(nA, nB);
                const int base = (head + n > NSLOT) ? 0 : head;
                head = base + n;
                const bool mine = (e3 == my);
                const uint32_t fbar = full0 + 8u * ej, ebar = empty0 + 8u * ej;
                const unsigned par = epar;
                if (++e3 == 3) e3 = 0;
                if (++ej == NE) { ej = 0; epar ^= 1u; }
                if (!mine) continue;
                npairs += (unsigned)npr;
                const uint64_t adesc0 = adescA + (uint64_t)((base * C::SLOT_A) >> 4), bdesc0 = bdescB + (uint64_t)((base * C::SLOT_B) >> 4);
                mbar_wait_a(fbar, par);
                if (!go_seen) {
                    mbar_wait_a(tgo, (unsigned)(lt & 1));
                    go_seen = true;
                }
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                if (elect_one_sync()) {
                    if (!(wi & 2)) {
                        if (c == 0) oz_issue_chunk<S, 0, 0, X>(adesc0, bdesc0, tmem_base, idesc_base, 1u);   // zero-initialises every column
                        else oz_dispatch<S, X>(za, zb, adesc0, bdesc0, tmem_base, idesc_base);
                    }
                    umma_commit_a(ebar);                    // the slots are free once these MMAs have read them
                    if (c == 0) mbar_arrive_a(tgo);         // the zero-initialising products are in the pipe: the other warps may issue
                }
            }
            if (!go_seen) mbar_wait_a(tgo, (unsigned)(lt & 1));     // keep this warp's phase count of tile_go in step
            if (elect_one_sync()) {
                umma_commit_a(accf);                         // acc_full completes when all three warps' products of the tile are done
                if (exec_pairs != nullptr && npairs != 0u) atomicAdd(exec_pairs, (unsigned long long)npairs);
            }
        }
    } else if (warp >= 4) {
        // ---------------- epilogue warps 4..7: TMEM lane quarter = warp % 4 ----------------
        const int quarter = warp & 3;
        const int q = quarter * 32 + lane;                      // query row within the tile = TMEM lane
        // digits are 1-based: diagonal d = a+b-2 carries 2^-bits(d+2).  The diagonals are recombined three at a time in 64-bit INTEGER
        // arithmetic (|acc| < 2^31, so acc_d 2^2b + acc_d+1 2^b + acc_d+2 < 2^48: exact), each group is converted once through the
        // mantissa trick (exact below 2^51) and only the NG = ceil(S / 3) group values meet in FP64: 2 NG + 1 FP64 instructions per
        // output instead of 2 S -- the epilogue's FP64 work is what the next tile's first MMAs wait for (acc_empty).
        constexpr int ND = S + (SKIP ? X : 0);                  // diagonals held in TMEM
        constexpr int NG = (ND + 2) / 3;
        double gw[NG];                                          // weight of the LAST diagonal of each group
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const int dlast = (3 * g + 2 < ND) ? 3 * g + 2 : ND - 1;
            gw[g] = ldexp(1.0, -digit_bits * (dlast + 2));
        }
        for (int lt = 0;; ++lt) {
            mbar_wait(&slot_full[lt & 1], (lt >> 1) & 1);
            const int t = tile_slot[lt & 1];
            __syncwarp();
            if (lane == 0) mbar_arrive(&slot_empty[lt & 1]);
            if (t < 0) break;
            int rt, ti;
            oz_tile_decode(t, T64, rowtiles, rt, ti);
            mbar_wait(&acc_full, lt & 1);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            double ssum = 0.0;
            if (wi & 4) {
                asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty);
            }
            // skipping variant: warps 4..7 take columns 0..31, warps 8..11 columns 32..63 of the tile (two partial sums per row and tile)
            const int cb0 = (SKIP && warp >= 8) ? ON / 32 : 0, cb1 = SKIP ? cb0 + ON / 32 : ON / 16;
#pragma unroll 1
            for (int cb = (wi & 4) ? cb1 : cb0; cb < cb1; ++cb) {
                double v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = 0.0;
                uint32_t r[ND][16];
#pragma unroll
                for (int d = 0; d < ND; ++d) {         // all diagonals of this column block in flight, one wait
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(d * ON + cb * 16);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                        : "=r"(r[d][0]), "=r"(r[d][1]), "=r"(r[d][2]), "=r"(r[d][3]), "=r"(r[d][4]), "=r"(r[d][5]), "=r"(r[d][6]), "=r"(r[d][7]),
                          "=r"(r[d][8]), "=r"(r[d][9]), "=r"(r[d][10]), "=r"(r[d][11]), "=r"(r[d][12]), "=r"(r[d][13]), "=r"(r[d][14]), "=r"(r[d][15])
                        : "r"(taddr));
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const int d0 = 3 * g, cnt = (d0 + 3 <= ND) ? 3 : ND - d0;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        long long acc = (long long)(int)r[d0][j];
#pragma unroll
                        for (int e = 1; e < cnt; ++e) acc = acc * (1LL << digit_bits) + (long long)(int)r[d0 + e][j];
                        // int64 -> double without the conversion pipe: 1.5 * 2^52 + acc sits exactly in the mantissa (|acc| < 2^51)
                        const double da = __longlong_as_double(acc + 0x4338000000000000LL) - 6755399441055744.0;
                        v[j] = fma(da, gw[g], v[j]);
                    }
                }
                if (cb == cb1 - 1) {
                    // every accumulator column of this tile has been read: let the MMA thread start the next tile
                    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty);
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const double vv = v[j] * __ldg(scaleB + (long long)ti * ON + cb * 16 + j);
                    ssum = fma(vv, vv, ssum);
                }
            }
            const long long grow = (long long)rt * OM + q;
            const double sa = scaleA[grow];
            part[(long long)(ti + ((SKIP && warp >= 8) ? T64 : 0)) * rows_total + grow] = ssum * sa * sa;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
}

}  // namespace oz
}  // namespace gptb
