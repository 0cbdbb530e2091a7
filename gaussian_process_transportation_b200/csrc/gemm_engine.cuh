// FP64 tile engine for sm_100a: C(128x128) += A(128xK) * B(128xK)^T with both operands K-contiguous ("NT").
//
// Every dense contraction on the GP path is phrased in this form (DESIGN.md "one engine"):
//   Cholesky panel/trailing updates, triangular inversion, K^-1, and the variance triangular multiply.
// FP64 has no tcgen05 path on Blackwell (tcgen05.mma kinds are f16/tf32/f8f6f4/i8/mx*), so the tensor work is
// mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) fed from shared memory.  Measured on B200: DMMA and DFMA both peak at
// 37.0 TFLOP/s (profiles/r01_fp64_peaks.json), cuBLAS DGEMM reaches 35.4.
//
// Operand feed: one producer warp issues TMA bulk-tensor copies (cp.async.bulk.tensor.3d, SASS UTMALDG) into a 3-deep
// ring of 64 KB stages guarded by full/empty mbarriers; eight consumer warps (2 x 4, warp tile 64 x 32) only execute
// LDS + DMMA.  The all-thread cp.async (LDGSTS) feed this replaced capped the kernel at 28 TFLOP/s (75 % of the pipe):
// the LDGSTS issue burst and its address arithmetic share the LSU/MIO path with the fragment loads
// (tools/feed_bench.cu: cp.async 28.1, TMA 36.7 TFLOP/s with the identical inner loop; profiles/r01_feed_bench.log).
//
// Shared-memory layout per operand slab (128 rows x 32 k):  [k/4][row][k%4]  -> a warp's fragment load
// (row = g = lane/4, k = t = lane%4) touches 32 consecutive doubles (bank-conflict free).  TMA produces that layout
// directly from a row-major matrix through a 3-D tensor map (4, rows, K/4) with strides (8 B, ld*8 B, 32 B) and box
// (4, 128, 8): the dimension order is what transposes the k-groups outward.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gptb {

constexpr int TS = 128;               // tile edge (rows of A, rows of B, and the k super-block)
constexpr int BK = 32;                // k-slab per pipeline stage
constexpr int SLABS_PER_TILE = TS / BK;
constexpr int NSTAGE = 3;             // 3 x 64 KB: the producer runs up to two slabs (~8 us of DMMA work) ahead
constexpr int CONSUMER_THREADS = 256; // 8 warps: 2 (m) x 4 (n), warp tile 64 x 32
constexpr int GEMM_THREADS = CONSUMER_THREADS + 32;   // + 1 TMA producer warp
constexpr int SLAB_DOUBLES = TS * BK; // one operand slab
constexpr int GEMM_SMEM_BYTES = NSTAGE * 2 * SLAB_DOUBLES * 8;   // 192 KB

enum { MASK_NONE = 0, MASK_LOWER = 1 /* keep k <= r */, MASK_UPPER = 2 /* keep k >= r */ };

struct Operand {
    const CUtensorMap* map;   // 3-D view (4, rows, K/4) of a row-major matrix, box (4, 128, BK/4)
    int row0;                 // first row of the 128-row operand tile
    int k0;                   // element offset added to the k index (lets a kernel address a sub-block as k-tile 0)
    int mask;                 // MASK_* applied on the k-tile `diag_kt` (the operand tile sits on its matrix' diagonal there)
    int diag_kt;              // k-tile index where the mask applies (-1: never)
};

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// ---- mbarrier / TMA primitives -------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(b)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(smem_u32(dst)),
        "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// barrier over the 8 consumer warps only (the producer warp never joins)
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;\n" ::"n"(CONSUMER_THREADS) : "memory"); }

struct PipeBarriers {
    uint64_t full[NSTAGE];
    uint64_t empty[NSTAGE];
};

// Called once at kernel entry by all GEMM_THREADS threads.
__device__ __forceinline__ void pipe_init(PipeBarriers* pb) {
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NSTAGE; ++i) {
            mbar_init(&pb->full[i], 1);                    // producer's expect_tx arrival; TMA completes the bytes
            mbar_init(&pb->empty[i], CONSUMER_THREADS / 32);   // one arrival per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
}

// One k-slab of DMMA work for a consumer warp.  s = absolute slab index (k = s*BK).
__device__ __forceinline__ void consume_slab(const Operand& A, const Operand& B, const double* sA, const double* sB, int s,
                                             double (&acc)[8][4][2], int wm, int wn, int g, int t) {
    const int kt = s / SLABS_PER_TILE, sl = s % SLABS_PER_TILE;
    const bool mA = (A.mask != MASK_NONE) && (kt == A.diag_kt);
    const bool mB = (B.mask != MASK_NONE) && (kt == B.diag_kt);
    const double* pa = sA + ((wm * 64 + g) << 2) + t;
    const double* pbf = sB + ((wn * 32 + g) << 2) + t;
#pragma unroll
    for (int kg = 0; kg < BK / 4; ++kg) {
        double a[8], b[4];
#pragma unroll
        for (int mi = 0; mi < 8; ++mi) a[mi] = pa[(kg * TS + mi * 8) << 2];
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) b[ni] = pbf[(kg * TS + ni * 8) << 2];
        if (mA | mB) {          // block-uniform branch: only on slabs that cross a diagonal tile
            const int kl = sl * BK + kg * 4 + t;
            if (mA) {
#pragma unroll
                for (int mi = 0; mi < 8; ++mi) {
                    int r = wm * 64 + mi * 8 + g;
                    bool keep = (A.mask == MASK_LOWER) ? (kl <= r) : (kl >= r);
                    a[mi] = keep ? a[mi] : 0.0;
                }
            }
            if (mB) {
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                    int r = wn * 32 + ni * 8 + g;
                    bool keep = (B.mask == MASK_LOWER) ? (kl <= r) : (kl >= r);
                    b[ni] = keep ? b[ni] : 0.0;
                }
            }
        }
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
    }
}

__device__ __forceinline__ void acc_clear(double (&acc)[8][4][2]) {
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) { acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0; }
}

// Persistent job loop: this CTA runs jobs job0, job0+stride, ... < njobs.  jobfn(job, A, B, kt_begin, kt_end) describes a
// 128x128 output tile; epifn(job, A, B, acc) is called by the consumer threads with the finished accumulators
// (acc[mi][ni][0..1] <-> C[wm*64 + mi*8 + g][wn*32 + ni*8 + 2t + {0,1}], g = lane>>2, t = lane&3, wm = warp>>2,
// wn = warp&3).  The producer warp keeps streaming the next jobs' slabs through the ring while the consumers run
// their epilogue, so per-tile prologue/epilogue latency is hidden.  Must be called by all GEMM_THREADS threads, once.
//
// CPREF = true additionally streams the job's 128x128 OUTPUT tile (rows A.row0, columns B.row0.. of the matrix behind
// cmap) through the same ring as two extra entries right behind the job's operand slabs, so a read-modify-write
// epilogue finds the old values in shared memory instead of stalling on global loads.  Entry e holds columns
// [64e, 64e+64) as two operand-style boxes of 32 columns ([c/4][row][c%4]); consumer warp (wm, wn) owns box wn and is
// handed its box pointer: epifn(job, A, B, acc, box).
template <bool CPREF, typename JobFn, typename EpiFn>
__device__ __forceinline__ void gemm_nt_jobs(int job0, int stride, int njobs, JobFn jobfn, EpiFn epifn, double* smem,
                                             PipeBarriers* pb, const CUtensorMap* cmap = nullptr) {
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    int gs = 0;   // slabs issued / consumed so far by this CTA (identical sequence on both sides)
    if (warp == CONSUMER_THREADS / 32) {
        // ---------------- TMA producer: one elected lane ----------------
        if (lane == 0) {
            for (int job = job0; job < njobs; job += stride) {
                Operand A, B;
                int kb, ke;
                jobfn(job, A, B, kb, ke);
                const int nslab = (ke - kb) * SLABS_PER_TILE, s0 = kb * SLABS_PER_TILE;
                for (int it = 0; it < nslab; ++it, ++gs) {
                    const int st = gs % NSTAGE;
                    if (gs >= NSTAGE) mbar_wait(&pb->empty[st], ((gs / NSTAGE) - 1) & 1);
                    double* sA = smem + st * 2 * SLAB_DOUBLES;
                    double* sB = sA + SLAB_DOUBLES;
                    const int kel = (s0 + it) * BK;
                    mbar_expect_tx(&pb->full[st], 2 * SLAB_DOUBLES * 8);
                    tma_load_3d(sA, A.map, 0, A.row0, (A.k0 + kel) >> 2, &pb->full[st]);
                    tma_load_3d(sB, B.map, 0, B.row0, (B.k0 + kel) >> 2, &pb->full[st]);
                }
                if constexpr (CPREF) {
                    for (int e = 0; e < 2; ++e, ++gs) {
                        const int st = gs % NSTAGE;
                        if (gs >= NSTAGE) mbar_wait(&pb->empty[st], ((gs / NSTAGE) - 1) & 1);
                        double* sA = smem + st * 2 * SLAB_DOUBLES;
                        mbar_expect_tx(&pb->full[st], 2 * SLAB_DOUBLES * 8);
                        tma_load_3d(sA, cmap, 0, A.row0, (B.row0 + 64 * e) >> 2, &pb->full[st]);
                        tma_load_3d(sA + SLAB_DOUBLES, cmap, 0, A.row0, (B.row0 + 64 * e + 32) >> 2, &pb->full[st]);
                    }
                }
            }
        }
        return;
    }
    // ---------------- consumers: LDS + DMMA ----------------
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;
    for (int job = job0; job < njobs; job += stride) {
        Operand A, B;
        int kb, ke;
        jobfn(job, A, B, kb, ke);
        const int nslab = (ke - kb) * SLABS_PER_TILE, s0 = kb * SLABS_PER_TILE;
        double acc[8][4][2];
        acc_clear(acc);
        for (int it = 0; it < nslab; ++it, ++gs) {
            const int st = gs % NSTAGE;
            mbar_wait(&pb->full[st], (gs / NSTAGE) & 1);
            const double* sA = smem + st * 2 * SLAB_DOUBLES;
            consume_slab(A, B, sA, sA + SLAB_DOUBLES, s0 + it, acc, wm, wn, g, t);
            __syncwarp();
            if (lane == 0) mbar_arrive(&pb->empty[st]);      // this warp is done reading stage st
        }
        if constexpr (CPREF) {
            for (int e = 0; e < 2; ++e, ++gs) {
                const int st = gs % NSTAGE;
                mbar_wait(&pb->full[st], (gs / NSTAGE) & 1);
                if ((wn >> 1) == e) epifn(job, A, B, acc, smem + st * 2 * SLAB_DOUBLES + (wn & 1) * SLAB_DOUBLES);
                __syncwarp();
                if (lane == 0) mbar_arrive(&pb->empty[st]);
            }
        } else {
            epifn(job, A, B, acc, (const double*)nullptr);
        }
    }
}

// Single-tile form: accumulates (does not clear) into acc.  kt range is in 128-wide k-tiles.  Must be called by all
// GEMM_THREADS threads, at most once per kernel (barrier phases start at 0).  On return the consumer warps have passed a
// consumer_sync(): shared memory may be reused by the epilogue.
__device__ __forceinline__ void gemm_nt_tile(const Operand& A, const Operand& B, int kt_begin, int kt_end,
                                             double (&acc)[8][4][2], double* smem, PipeBarriers* pb) {
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int nslab = (kt_end - kt_begin) * SLABS_PER_TILE;
    const int s0 = kt_begin * SLABS_PER_TILE;
    if (nslab <= 0) return;
    if (warp == CONSUMER_THREADS / 32) {
        if (lane == 0) {
            for (int it = 0; it < nslab; ++it) {
                const int st = it % NSTAGE;
                if (it >= NSTAGE) mbar_wait(&pb->empty[st], ((it / NSTAGE) - 1) & 1);
                double* sA = smem + st * 2 * SLAB_DOUBLES;
                double* sB = sA + SLAB_DOUBLES;
                const int kel = (s0 + it) * BK;
                mbar_expect_tx(&pb->full[st], 2 * SLAB_DOUBLES * 8);
                tma_load_3d(sA, A.map, 0, A.row0, (A.k0 + kel) >> 2, &pb->full[st]);
                tma_load_3d(sB, B.map, 0, B.row0, (B.k0 + kel) >> 2, &pb->full[st]);
            }
        }
        return;
    }
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;
    for (int it = 0; it < nslab; ++it) {
        const int st = it % NSTAGE;
        mbar_wait(&pb->full[st], (it / NSTAGE) & 1);
        const double* sA = smem + st * 2 * SLAB_DOUBLES;
        consume_slab(A, B, sA, sA + SLAB_DOUBLES, s0 + it, acc, wm, wn, g, t);
        __syncwarp();
        if (lane == 0) mbar_arrive(&pb->empty[st]);
    }
    consumer_sync();
}

__device__ __forceinline__ bool is_consumer() { return threadIdx.x < CONSUMER_THREADS; }

// visit every accumulator pair: f(row, col, v0 /*col*/, v1 /*col+1*/) with row/col local to the 128x128 tile
template <typename F>
__device__ __forceinline__ void acc_foreach(double (&acc)[8][4][2], F f) {
    if (threadIdx.x >= CONSUMER_THREADS) return;      // the producer warp holds no accumulators
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) f(wm * 64 + mi * 8 + g, wn * 32 + ni * 8 + 2 * t, acc[mi][ni][0], acc[mi][ni][1]);
}

// linear index -> (i, j) with 0 <= j <= i, row-major over the lower triangle
__device__ __forceinline__ void tri_decode(long long idx, int& i, int& j) {
    int r = (int)((sqrt(8.0 * (double)idx + 1.0) - 1.0) * 0.5);
    while ((long long)(r + 1) * (r + 2) / 2 <= idx) ++r;
    while ((long long)r * (r + 1) / 2 > idx) --r;
    i = r;
    j = (int)(idx - (long long)r * (r + 1) / 2);
}

}  // namespace gptb
