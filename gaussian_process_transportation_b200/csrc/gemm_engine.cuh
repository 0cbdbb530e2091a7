// FP64 tile engine for sm_100a: C(128x128) += A(128xK) * B(128xK)^T with both operands K-contiguous ("NT").
//
// Every dense contraction on the GP path is phrased in this form (DESIGN.md "one engine"):
//   Cholesky panel/trailing updates, triangular inversion, K^-1, and the variance triangular multiply.
// FP64 has no tcgen05 path on Blackwell (tcgen05.mma kinds are f16/tf32/f8f6f4/i8/mx*), so the tensor work is
// mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) fed from shared memory.  Measured on B200: DMMA and DFMA both peak at
// 37.0 TFLOP/s (profiles/r01_fp64_peaks.json), cuBLAS DGEMM reaches 35.4; the math pipe, not the memory system,
// is the limiter (a 128x128x16 slab is 4096 cycles of DMMA against 32 KB of operands), so operands are staged with a
// 3-deep cp.async ring and the fragment loads are laid out to be bank-conflict free.
//
// Shared-memory layout per operand slab (128 rows x 16 k):  [k/4][row][k%4]  -> a warp's fragment load
// (row = g = lane/4, k = t = lane%4) touches 32 consecutive doubles.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gptb {

constexpr int TS = 128;               // tile edge (rows of A, rows of B, and the k super-block)
constexpr int BK = 16;                // k-slab
constexpr int SLABS_PER_TILE = TS / BK;
constexpr int NSTAGE = 3;
constexpr int GEMM_THREADS = 256;     // 8 warps: 2 (m) x 4 (n), warp tile 64 x 32
constexpr int SLAB_DOUBLES = TS * BK; // one operand slab
constexpr int GEMM_SMEM_BYTES = NSTAGE * 2 * SLAB_DOUBLES * 8;   // 96 KB

enum { MASK_NONE = 0, MASK_LOWER = 1 /* keep k <= r */, MASK_UPPER = 2 /* keep k >= r */ };

struct Operand {
    const double* base;   // pointer to (first row of the 128-row operand tile, column 0 of k-space)
    long long ld;         // leading dimension (doubles), multiple of 16
    int mask;             // MASK_* applied on the k-tile `diag_kt` (the operand tile sits on its matrix' diagonal there)
    int diag_kt;          // k-tile index where the mask applies (-1: never)
};

__device__ __forceinline__ void cp_async16(void* smem_ptr, const void* gmem_ptr) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_ptr));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_ptr));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// issue the cp.async copies of one (A,B) slab pair; k0 = absolute k column of the slab
__device__ __forceinline__ void load_slab(double* sA, double* sB, const Operand& A, const Operand& B, long long k0, int tid) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int c = tid + GEMM_THREADS * i;      // 1024 16-byte chunks per operand slab
        int row = c >> 3, kc = c & 7;
        int soff = (((kc >> 1) * TS + row) << 2) + ((kc & 1) << 1);
        cp_async16(sA + soff, A.base + (long long)row * A.ld + k0 + (kc << 1));
        cp_async16(sB + soff, B.base + (long long)row * B.ld + k0 + (kc << 1));
    }
}

// acc[mi][ni][0..1] <-> C[wm*64 + mi*8 + g][wn*32 + ni*8 + 2t + {0,1}],  g = lane>>2, t = lane&3,
// wm = warp>>2, wn = warp&3.  Accumulates (does not clear) into acc.  kt range is in 128-wide k-tiles.
__device__ __forceinline__ void gemm_nt_tile(const Operand& A, const Operand& B, int kt_begin, int kt_end,
                                             double (&acc)[8][4][2], double* smem) {
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;
    const int nslab = (kt_end - kt_begin) * SLABS_PER_TILE;
    const int s0 = kt_begin * SLABS_PER_TILE;
    if (nslab <= 0) return;

#pragma unroll
    for (int st = 0; st < NSTAGE - 1; ++st) {
        if (st < nslab) load_slab(smem + st * 2 * SLAB_DOUBLES, smem + st * 2 * SLAB_DOUBLES + SLAB_DOUBLES, A, B,
                                  (long long)(s0 + st) * BK, tid);
        cp_async_commit();
    }
    for (int it = 0; it < nslab; ++it) {
        cp_async_wait<NSTAGE - 2>();
        __syncthreads();
        {
            int nx = it + NSTAGE - 1;
            if (nx < nslab) {
                int st = nx % NSTAGE;
                load_slab(smem + st * 2 * SLAB_DOUBLES, smem + st * 2 * SLAB_DOUBLES + SLAB_DOUBLES, A, B,
                          (long long)(s0 + nx) * BK, tid);
            }
            cp_async_commit();
        }
        const double* sA = smem + (it % NSTAGE) * 2 * SLAB_DOUBLES;
        const double* sB = sA + SLAB_DOUBLES;
        const int s = s0 + it;
        const int kt = s >> 3, sl = s & 7;
        const bool mA = (A.mask != MASK_NONE) && (kt == A.diag_kt);
        const bool mB = (B.mask != MASK_NONE) && (kt == B.diag_kt);
        const double* pa = sA + ((wm * 64 + g) << 2) + t;
        const double* pb = sB + ((wn * 32 + g) << 2) + t;
#pragma unroll
        for (int kg = 0; kg < 4; ++kg) {
            double a[8], b[4];
#pragma unroll
            for (int mi = 0; mi < 8; ++mi) a[mi] = pa[(kg * TS + mi * 8) << 2];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) b[ni] = pb[(kg * TS + ni * 8) << 2];
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
    cp_async_wait<0>();
    __syncthreads();       // smem may be reused by the caller's epilogue / next tile
}

__device__ __forceinline__ void acc_clear(double (&acc)[8][4][2]) {
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) { acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0; }
}

// visit every accumulator pair: f(row, col, v0 /*col*/, v1 /*col+1*/) with row/col local to the 128x128 tile
template <typename F>
__device__ __forceinline__ void acc_foreach(double (&acc)[8][4][2], F f) {
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
