// Fit-side kernels: Gram build (RBF+White), blocked FP64 Cholesky, triangular solves for alpha, triangular inverse,
// K^-1 and the fused log-marginal-likelihood gradient pass.
//
// Storage convention for every N x N factor buffer (row-major, ld = Npad = roundup(N,128)):
//   lower triangle (incl. diagonal) holds the matrix M, the strict upper triangle holds M^T ("mirror"),
// so both M and M^T are available K-contiguous for the NT tile engine without extra memory.  On the 128x128
// diagonal tiles the two halves meet; the engine masks the wrong half in registers (gemm_engine.cuh).
// Rows/cols >= N are padding: the padded Gram block is the identity, so L, L^-1 are the identity there too.
#pragma once
#include "gemm_engine.cuh"

namespace gptb {

constexpr int MAXD = 4;
constexpr int MAXP = 4;

// Radial profile of the stationary part: RBF (sklearn:kernels.py:1558-1587) or Matern nu = 1.5 / 2.5
// (sklearn:kernels.py Matern.__call__), all as functions of s = sum_a ((x_a - y_a)/ell_a)^2.
enum { KIND_RBF = 0, KIND_MATERN15 = 1, KIND_MATERN25 = 2, KIND_MATERN05 = 3 };

struct KParams {
    double c, s2, jitter;
    double ell[MAXD];
    double inv_ell[MAXD];
    int kind;
};

// profile(s) with k = c * profile; EXPF is the exponential to use for a non-positive argument
template <typename EXPF>
__device__ __forceinline__ double kernel_profile(double s, int kind, EXPF expf_neg) {
    if (kind == KIND_RBF) return expf_neg(-0.5 * s);
    const double r = sqrt(s);
    if (kind == KIND_MATERN05) return expf_neg(-r);            // exponential kernel (sklearn Matern nu=0.5)
    if (kind == KIND_MATERN15) {
        const double t = 1.7320508075688772 * r;          // sqrt(3) * dist
        return (1.0 + t) * expf_neg(-t);
    }
    const double t = 2.23606797749979 * r;                // sqrt(5) * dist
    return (1.0 + t + t * t / 3.0) * expf_neg(-t);
}

// d profile / d log ell_a = grad_factor(s) * d2_a  with d2_a = ((x_a - y_a)/ell_a)^2   (sklearn Matern/RBF eval_gradient)
__device__ __forceinline__ double kernel_grad_factor(double s, int kind) {
    if (kind == KIND_RBF) return exp(-0.5 * s);
    if (kind == KIND_MATERN05) return s > 0.0 ? exp(-sqrt(s)) / sqrt(s) : 0.0;   // sklearn zeroes the 0/0 entries (kernels.py Matern nu=0.5)
    if (kind == KIND_MATERN15) return 3.0 * exp(-sqrt(3.0 * s));
    const double t = sqrt(5.0 * s);
    return 5.0 / 3.0 * (t + 1.0) * exp(-t);
}

// ------------------------------------------------------------------------------------------------------------
// Xs[a][n] = X[a][n] / ell[a]   (sklearn divides the inputs by the length-scale before differencing,
// sklearn:kernels.py:1561-1565); SoA, padded entries stay 0.
// ------------------------------------------------------------------------------------------------------------
__global__ void scale_inputs_kernel(const double* __restrict__ X, double* __restrict__ Xs, int N, int Npad, int d, KParams kp) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= Npad) return;
    for (int a = 0; a < d; ++a) Xs[(long long)a * Npad + n] = (n < N) ? X[(long long)a * Npad + n] / kp.ell[a] : 0.0;
}

// ------------------------------------------------------------------------------------------------------------
// Gram build, lower tiles only (diagonal tiles are written in full).  K = c*exp(-r2/2) + (s2+jitter) on the diagonal.
// One CTA per 128x128 tile; a warp writes 1 KB of a row per store instruction (4 doubles per lane).
// Roofline: 8*N^2/2 bytes written once vs N^2/2 FP64 exp (~21 DFMA slots each, profiles/r01_fp64_peaks.json) ->
// exp-bound on B200 (37 TF FP64 vs 6.4 TB/s HBM).
// ------------------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) gram_lower_kernel(const double* __restrict__ Xs, double* __restrict__ K, int N, int Npad,
                                                         KParams kp) {
    int ti, tj;
    tri_decode(blockIdx.x, ti, tj);
    __shared__ double xi[D][TS];
    const int tid = threadIdx.x;
    for (int e = tid; e < D * TS; e += 256) xi[e / TS][e % TS] = Xs[(long long)(e / TS) * Npad + ti * TS + (e % TS)];
    __syncthreads();
    const int cg = (tid & 31) * 4;   // 4 consecutive columns
    const int r0 = tid >> 5;         // rows r0, r0+8, ...
    double xj[D][4];
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
        for (int q = 0; q < 4; ++q) xj[a][q] = Xs[(long long)a * Npad + tj * TS + cg + q];
    const double diag_add = kp.s2 + kp.jitter;
#pragma unroll 2
    for (int rr = 0; rr < 16; ++rr) {
        int r = r0 + rr * 8;
        int gi = ti * TS + r;
        double out[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int gj = tj * TS + cg + q;
            double s = 0.0;
#pragma unroll
            for (int a = 0; a < D; ++a) {
                double df = xi[a][r] - xj[a][q];
                s += df * df;
            }
            double v = kp.c * kernel_profile(s, kp.kind, [](double z) { return exp(z); });
            if (gi == gj) v += diag_add;
            if (gi >= N || gj >= N) v = (gi == gj) ? 1.0 : 0.0;
            out[q] = v;
        }
        double4 o = make_double4(out[0], out[1], out[2], out[3]);
        *reinterpret_cast<double4*>(K + (long long)gi * Npad + tj * TS + cg) = o;
    }
}

// ------------------------------------------------------------------------------------------------------------
// Diagonal tile (128 x 128): Cholesky factor and its explicit inverse, entirely in one CTA's shared memory.
// This kernel is the serial spine of the blocked factorisation, so it is built for latency:
//   * 32 x 32 diagonal blocks are factored and inverted by ONE warp in registers (lane i owns row i; pivots,
//     column scalings and rank-1 updates travel by warp shuffles -- no barriers inside a block);
//   * everything between blocks (panel solve as a product with the block inverse, trailing update, assembly of the
//     off-diagonal blocks of the inverse) is a small DMMA product on shared-memory operands by all 8 warps.
// Writes L (lower) + L^T (upper) back into the factor buffer's diagonal tile and Dinv = L_kk^-1 (dense, zeros above
// the diagonal) into dinv.  info: first non-positive pivot (1-based global order), LAPACK dpotrf convention.
// Shared layout: S[128][132] (leading dimension = 4 mod 16 makes both [row][k] and [k][row] DMMA fragment reads
// conflict-free), Zd[4][32][36] block inverses, Tm[32][132] scratch.
// ------------------------------------------------------------------------------------------------------------
constexpr int DLD = TS + 4;          // 132
constexpr int ZLD = 36;
constexpr int DB = 32;               // inner block
constexpr int DIAG_SMEM_BYTES = (TS * DLD + 4 * DB * ZLD + DB * DLD) * 8;   // 205,824 B

// C[M x N] = beta*C + alpha * A * B^T on shared-memory operands, 8 warps.  Element (m,k) of A is A[m*lda+k] (or
// A[k*lda+m] when A_COL); element (n,k) of B is B[n*ldb+k] (or B[k*ldb+n] when B_COL).  M, N multiples of 8, K of 4.
// Work unit: 8 rows x up to 32 columns (4 accumulator fragments) per warp.  lower_only skips 8x8 blocks above the diagonal.
// A warp reads every A fragment of its row strip before it stores, so C may alias A's rows (in-place panel solve).
template <bool A_COL, bool B_COL>
__device__ __forceinline__ void smem_gemm(double* C, int ldc, const double* A, int lda, const double* B, int ldb, int M, int N,
                                          int K, double alpha, double beta, bool lower_only) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int rbs = M >> 3, cgs = (N + 31) >> 5;
    for (int task = warp; task < rbs * cgs; task += 8) {
        const int rb = task / cgs, cg = task % cgs;
        const int r0 = rb * 8, c0 = cg * 32;
        int nb = min(4, (N - c0) >> 3);
        if (lower_only) nb = min(nb, rb - cg * 4 + 1);
        if (nb <= 0) continue;
        double acc[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
        for (int k0 = 0; k0 < K; k0 += 4) {
            const double a = A_COL ? A[(k0 + t) * lda + r0 + g] : A[(r0 + g) * lda + k0 + t];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (j < nb) {
                    const int n = c0 + j * 8 + g;
                    const double b = B_COL ? B[(k0 + t) * ldb + n] : B[n * ldb + k0 + t];
                    dmma884(acc[j][0], acc[j][1], a, b);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j < nb) {
                double* dst = C + (r0 + g) * ldc + c0 + j * 8 + 2 * t;
                double o0 = alpha * acc[j][0], o1 = alpha * acc[j][1];
                if (beta != 0.0) { o0 += dst[0]; o1 += dst[1]; }
                dst[0] = o0;
                dst[1] = o1;
            }
        }
    }
}

// One warp: Cholesky of a 32x32 block (lane i holds row i in a[0..31], lower part meaningful; dg = a[lane], the lane's own diagonal
// entry) and, in the same 32 column steps, the inverse of the factor.  On return a[] holds row i of L (entries above the diagonal are
// unspecified), x[] holds COLUMN `lane` of L^-1 (x[r] = Linv[r][lane], zero for r < lane).  Returns the 1-based index of the first
// non-positive pivot in the block, 0 if none.
//   * column j of L is broadcast through shared memory (`scratch`: 64 doubles, warp-private, double-buffered).  The first version
//     broadcast every L[k][j] with a 64-bit shuffle (62 SHFL per column step, ~1000 for the inverse) and spent 63 % of the
//     diagonal-tile kernel here (tools/diag_prof.py: 19.7 k cycles per block);
//   * the inverse is the COLUMN-oriented forward substitution (lane c solves L x = e_c): once x_j is known, the running sums of the
//     rows below take  s_k += L[k][j] x_j  -- the same broadcast column the factor's rank-1 update reads, so the inverse costs no
//     extra loads and its short chain (s_j -> x_j) hides under the pivot latency instead of following the factor (17 k -> 7 k cycles);
//   * the pivot chain never goes through shared memory: every lane keeps its own diagonal entry up to date with its own l
//     (dg -= l^2, the same value the general update would produce), so a step's critical path is SHFL -> rsqrt -> mul -> DFMA.
constexpr int WARP_BLOCK_SCRATCH = 2 * DB;                       // doubles
constexpr double DBL_MIN_POS = 2.2250738585072014e-308;
// 1/sqrt(d) for a normal positive d, branch-free: MUFU.RSQ64H seed (relative error 2^-22) and one cubic step, the sequence CUDA's
// rsqrt() uses on its fast path -- without its slow-path branch, which keeps the pivot chain and the rank-1 updates around it in one
// basic block, so the scheduler overlaps them (with sqrt() + rsqrt() calls the warp sat through ~13 dependent FP64 operations per
// column before it issued anything else).
__device__ __forceinline__ double rsqrt_normal(double d) {
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d));
    const double e = fma(d, -(y0 * y0), 1.0);
    const double pc = fma(e, 0.375, 0.5);
    return fma(pc, y0 * e, y0);
}
__device__ __forceinline__ int warp_potrf_trtri32(double (&a)[DB], double (&x)[DB], double dg, int lane, double* scratch) {
    const unsigned full = 0xffffffffu;
    double* col = scratch;                                       // [2][32]: current column of L, double-buffered
    int bad = 0;
#pragma unroll
    for (int c = 0; c < DB; ++c) x[c] = 0.0;                     // x[k], k > j: running sum of row k; becomes x_k at step k
    // software-pipelined by hand: the pivot of step j+1 (SHFL -> rsqrt) is started as soon as the lane's diagonal entry is up to
    // date, before the rank-1 updates of step j are issued (a warp issues in order)
    double d = __shfl_sync(full, dg, 0);
    double rinv = rsqrt_normal(d >= DBL_MIN_POS ? d : 1.0);
#pragma unroll
    for (int j = 0; j < DB; ++j) {
        const bool ok = d >= DBL_MIN_POS;                     // non-positive (or subnormal) pivot: LAPACK's info, the block carries on with 1
        if (!ok && bad == 0) bad = j + 1;
        const double l = a[j] * rinv;                         // column j of L (lanes > j); LAPACK dpotf2 scales by 1/ajj too
        const double sd = d * rinv;                           // the pivot itself: d * d^-1/2 with one correction step (off the chain)
        const double lref = fma(fma(-sd, sd, d), 0.5 * rinv, sd);
        a[j] = (lane == j) ? (ok ? lref : 1.0) : l;
        dg = fma(-l, l, dg);                                  // own diagonal entry (meaningful on lanes > j)
        double* cj = col + (j & 1) * DB;
        cj[lane] = l;
        const double xj = (((lane == j) ? 1.0 : 0.0) - x[j]) * rinv;     // lanes > j: exactly zero
        x[j] = xj;
        if (j + 1 < DB) {
            d = __shfl_sync(full, dg, j + 1);
            rinv = rsqrt_normal(d >= DBL_MIN_POS ? d : 1.0);
        }
        __syncwarp();
#pragma unroll
        for (int k = j + 1; k < DB; ++k) {
            const double ck = cj[k];
            a[k] = fma(-l, ck, a[k]);                         // rank-1 update (only i >= k is used later)
            x[k] = fma(ck, xj, x[k]);
        }
    }
    __syncwarp();
    return bad;
}

// Forward substitution is fused here: when rhs != nullptr the kernel also emits z_k = L_kk^-1 y_k (y_k = rows of the
// running right-hand side, already updated by the panels of the previous steps) into sol.
__global__ void __launch_bounds__(256, 1) potrf_diag_kernel(double* __restrict__ Lbuf, long long ld, int kt, double* __restrict__ dinv,
                                                            int* info, const double* __restrict__ rhs, double* __restrict__ sol,
                                                            int Npad, int p, long long* __restrict__ prof = nullptr) {
    extern __shared__ __align__(16) double sm[];
    __shared__ double ys[8][TS];             // right-hand sides, rows >= p zero (the 8 columns of a DMMA B fragment)
    // prof (developer hook, gptb_test_potrf_tile): clock64 at the phase boundaries, written by thread 0
    int pslot = 0;
    auto stamp = [&]() {
        if (prof != nullptr && threadIdx.x == 0) {
            prof[pslot] = clock64();
            if (pslot == 0 || pslot == 12) {              // wall clock (ns) beside the first and last cycle stamp: the SM clock the tile ran at
                unsigned long long ns;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
                prof[pslot == 0 ? 20 : 21] = (long long)ns;
            }
        }
        ++pslot;
    };
    stamp();
    if (rhs != nullptr)
        for (int e = threadIdx.x; e < 8 * TS; e += 256) ys[e / TS][e % TS] = (e / TS < p) ? rhs[(long long)(e / TS) * Npad + kt * TS + (e % TS)] : 0.0;
    double* S = sm;                         // [128][DLD]
    double* Zd = sm + TS * DLD;             // [4][32][ZLD]
    double* Tm = Zd + 4 * DB * ZLD;         // [32][DLD]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* tile = Lbuf + (long long)kt * TS * ld + kt * TS;
    // 32 double2 loads per thread, 16 in flight at a time (the tile comes from L2: latency, not bandwidth)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        double2 v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int e = tid + (half * 16 + u) * 256;
            v[u] = *reinterpret_cast<const double2*>(tile + (long long)(e >> 6) * ld + ((e & 63) << 1));
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int e = tid + (half * 16 + u) * 256;
            const int r = e >> 6, c2 = (e & 63) << 1;
            S[r * DLD + c2] = v[u].x;
            S[r * DLD + c2 + 1] = v[u].y;
        }
    }
    __syncthreads();
    stamp();                                 // 1: tile loaded
    // ---- blocked Cholesky, 32-wide block columns -----------------------------------------------------------------
    for (int b = 0; b < TS / DB; ++b) {
        const int j0 = b * DB;
        if (warp == 0) {
            double a[DB], x[DB];
#pragma unroll
            for (int c = 0; c < DB; ++c) a[c] = S[(j0 + lane) * DLD + j0 + c];
            const int bad = warp_potrf_trtri32(a, x, S[(j0 + lane) * DLD + j0 + lane], lane, Tm);   // Tm is free until the off-diagonal inverse phase
            if (bad && lane == 0) atomicCAS(info, 0, kt * TS + j0 + bad);
#pragma unroll
            for (int c = 0; c < DB; ++c) {
                if (c <= lane) S[(j0 + lane) * DLD + j0 + c] = a[c];
                Zd[(b * DB + c) * ZLD + lane] = x[c];          // x[r] = Linv[r][lane] -> row r, column lane
            }
        }
        __syncthreads();
        stamp();                             // 2, 4, 6, 8: 32 x 32 block factorised + inverted by warp 0
        const int rem = TS - j0 - DB;
        if (rem > 0) {
            // panel: rows below the block times the block inverse (in place)
            smem_gemm<false, false>(S + (j0 + DB) * DLD + j0, DLD, S + (j0 + DB) * DLD + j0, DLD, Zd + b * DB * ZLD, ZLD, rem, DB, DB,
                                    1.0, 0.0, false);
            __syncthreads();
            // trailing update of the remaining lower triangle
            smem_gemm<false, false>(S + (j0 + DB) * DLD + j0 + DB, DLD, S + (j0 + DB) * DLD + j0, DLD, S + (j0 + DB) * DLD + j0, DLD,
                                    rem, rem, DB, -1.0, 1.0, true);
            __syncthreads();
        }
        stamp();                             // 3, 5, 7, 9: panel + trailing update inside the tile
    }
    // ---- write L (lower) and its mirror (upper) back ---------------------------------------------------------------
    for (int e = tid; e < TS * TS; e += 256) {
        int r = e >> 7, c = e & 127;
        tile[(long long)r * ld + c] = (c <= r) ? S[r * DLD + c] : S[c * DLD + r];
    }
    __syncthreads();
    stamp();                                 // 10: L written back
    // ---- inverse: off-diagonal 32-blocks, block row by block row.  (Linv_ij)^T is kept in S's upper block (j,i). ----
    for (int i = 1; i < TS / DB; ++i) {
        // Tm[:, 32j..32j+31] = L_ij Linv_jj + sum_{k=j+1}^{i-1} L_ik Linv_kj        for every j < i
        for (int j = 0; j < i; ++j) {
            smem_gemm<false, true>(Tm + j * DB, DLD, S + (i * DB) * DLD + j * DB, DLD, Zd + j * DB * ZLD, ZLD, DB, DB, DB, 1.0, 0.0, false);
        }
        __syncthreads();
        for (int j = 0; j + 1 < i; ++j) {
            smem_gemm<false, false>(Tm + j * DB, DLD, S + (i * DB) * DLD + (j + 1) * DB, DLD, S + (j * DB) * DLD + (j + 1) * DB, DLD, DB, DB,
                                    (i - j - 1) * DB, 1.0, 1.0, false);
        }
        __syncthreads();
        // (Linv_ij)^T[n][m] = - sum_kk Tm[kk][32j+n] * Linv_ii[m][kk]   for all j < i at once (rows n = 0 .. 32i-1)
        smem_gemm<true, false>(S + i * DB, DLD, Tm, DLD, Zd + i * DB * ZLD, ZLD, i * DB, DB, DB, -1.0, 0.0, false);
        __syncthreads();
    }
    stamp();                                 // 11: off-diagonal blocks of the inverse
    double* dk = dinv + (long long)kt * TS * TS;
    for (int e = tid; e < TS * TS; e += 256) {
        int r = e >> 7, c = e & 127;
        int bi = r >> 5, bj = c >> 5;
        double v = 0.0;
        if (bi == bj) v = (c <= r) ? Zd[(bi * DB + (r & 31)) * ZLD + (c & 31)] : 0.0;
        else if (bi > bj) v = S[c * DLD + r];
        dk[r * TS + c] = v;
    }
    if (rhs != nullptr) {
        // z = Linv y as a DMMA product with the right-hand sides as the 8 (padded) columns of B: 16 strips of 8 rows, two per warp
        // (w and 15 - w: the work per strip grows with the row).  Off-diagonal blocks are read as (Linv_ij)^T from S's upper half,
        // the diagonal block from Zd.  (The first version ran 128 threads through a scalar loop with the accumulators indexed by a
        // run-time p -- local memory -- and took 35 k of the kernel's 114 k cycles.)
        const int g = lane >> 2, t = lane & 3;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            const int strip = half == 0 ? warp : 15 - warp;
            const int r0 = strip * 8, bi = r0 >> 5;
            double z0 = 0.0, z1 = 0.0;
            for (int k0 = 0; k0 < bi * DB; k0 += 4) dmma884(z0, z1, S[(k0 + t) * DLD + r0 + g], ys[g][k0 + t]);
            const double* zd = Zd + (bi * DB + (r0 & 31) + g) * ZLD;
            for (int k0 = 0; k0 < DB; k0 += 4) dmma884(z0, z1, zd[k0 + t], ys[g][bi * DB + k0 + t]);
            if (2 * t < p) sol[(long long)(2 * t) * Npad + kt * TS + r0 + g] = z0;
            if (2 * t + 1 < p) sol[(long long)(2 * t + 1) * Npad + kt * TS + r0 + g] = z1;
        }
    }
    __syncthreads();
    stamp();                                 // 12: inverse written, forward substitution done
}

// ------------------------------------------------------------------------------------------------------------
// Panel: L[i,k] = A[i,k] * Dinv_k^T for the row tiles i > k (in place), plus the mirror L[i,k]^T into the upper half.
// Fused forward substitution: y_i -= L[i,k] z_k with the freshly computed tile still in registers (rhs may be null).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEMM_THREADS, 1) potrf_panel_kernel(const __grid_constant__ CUtensorMap mapL,
                                                                     const __grid_constant__ CUtensorMap mapD,
                                                                     double* __restrict__ Lbuf, long long ld, int kt,
                                                                     double* __restrict__ rhs, const double* __restrict__ sol, int Npad,
                                                                     int p, int first_tile) {
    extern __shared__ __align__(128) double smem[];
    __shared__ PipeBarriers pipe;
    __shared__ double zs[MAXP][TS];
    __shared__ double red[MAXP][4][TS];
    if (rhs != nullptr)
        for (int e = threadIdx.x; e < p * TS; e += GEMM_THREADS) zs[e / TS][e % TS] = sol[(long long)(e / TS) * Npad + kt * TS + (e % TS)];
    pipe_init(&pipe);
    const int ti = first_tile + blockIdx.x;            // kt + 1, or kt + 2 when the spine kernel makes tile (kt+1, kt)
    double acc[8][4][2];
    acc_clear(acc);
    // k-space of this product is the 128 columns of block column kt: offset k so that k-tile 0 is that block
    Operand A{&mapL, ti * TS, kt * TS, MASK_NONE, -1};
    Operand B{&mapD, kt * TS, 0, MASK_NONE, -1};   // Dinv rows n, k <= n (zeros stored above)
    gemm_nt_tile(A, B, 0, 1, acc, smem, &pipe);
    if (!is_consumer()) return;
    double* out = Lbuf + (long long)ti * TS * ld + (long long)kt * TS;          // lower: rows ti, cols kt
    double* outT = Lbuf + (long long)kt * TS * ld + (long long)ti * TS;         // mirror: rows kt, cols ti
    acc_foreach(acc, [&](int r, int c, double v0, double v1) {
        *reinterpret_cast<double2*>(out + (long long)r * ld + c) = make_double2(v0, v1);
        outT[(long long)c * ld + r] = v0;
        outT[(long long)(c + 1) * ld + r] = v1;
    });
    if (rhs == nullptr) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3, wm = warp >> 2, wn = warp & 3;
    for (int q = 0; q < p; ++q) {
#pragma unroll
        for (int mi = 0; mi < 8; ++mi) {
            double sacc = 0.0;
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int c = wn * 32 + ni * 8 + 2 * t;
                sacc = fma(acc[mi][ni][0], zs[q][c], sacc);
                sacc = fma(acc[mi][ni][1], zs[q][c + 1], sacc);
            }
            sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
            sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
            if (t == 0) red[q][wn][wm * 64 + mi * 8 + g] = sacc;
        }
    }
    consumer_sync();
    if (threadIdx.x < TS) {
        const int r = threadIdx.x;
        for (int q = 0; q < p; ++q)
            rhs[(long long)q * Npad + ti * TS + r] -= (red[q][0][r] + red[q][1][r]) + (red[q][2][r] + red[q][3][r]);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Trailing update: A[i,j] -= sum_{kk in [k0,k1)} L[i,kk] L[j,kk]^T (a rank-(128*(k1-k0)) SYRK on the DMMA engine), as a
// persistent job loop so that the read-modify-write of tile t overlaps the operand stream of tile t+1.
//   mode 0: the lower triangle of tiles (i,j), base <= j <= i < T          (njobs = r(r+1)/2, r = T - base)
//   mode 1: the single tile column j = base, rows base <= i < T            (njobs = T - base)   -- the look-ahead column
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEMM_THREADS, 1) potrf_trailing_kernel(const __grid_constant__ CUtensorMap mapL,
                                                                        double* __restrict__ Lbuf, long long ld, int k0, int k1,
                                                                        int base, int mode, int njobs, int job_skip) {
    extern __shared__ __align__(128) double smem[];
    __shared__ PipeBarriers pipe;
    pipe_init(&pipe);
    auto tile_of = [&](int job, int& ti, int& tj) {
        job += job_skip;                               // 1: the triangle without its first tile (base, base) -- the spine kernel's
        if (mode == 0) {
            int a, b;
            tri_decode(job, a, b);
            ti = base + a;
            tj = base + b;
        } else {
            ti = base + job;
            tj = base;
        }
    };
    auto jobfn = [&](int job, Operand& A, Operand& B, int& kb, int& ke) {
        int ti, tj;
        tile_of(job, ti, tj);
        A = Operand{&mapL, ti * TS, 0, MASK_NONE, -1};
        B = Operand{&mapL, tj * TS, 0, MASK_NONE, -1};
        kb = k0;
        ke = k1;
    };
    // epilogue: old tile values arrive through the ring (box = this warp's 32 columns as [c/4][row][c%4]); the new
    // values leave as fire-and-forget global stores, so the DMMA pipe restarts on the next tile without a load stall
    auto epifn = [&](int job, const Operand& A, const Operand& B, double (&acc)[8][4][2], const double* box) {
        double* out = Lbuf + (long long)A.row0 * ld + (long long)B.row0;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int g = lane >> 2, t = lane & 3, wm = warp >> 2, wn = warp & 3;
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int r = wm * 64 + mi * 8 + g;
                const int kg = ni * 2 + (t >> 1), off = (2 * t) & 3;
                const double2 old = *reinterpret_cast<const double2*>(box + (((kg * TS) + r) << 2) + off);
                *reinterpret_cast<double2*>(out + (long long)r * ld + wn * 32 + ni * 8 + 2 * t) =
                    make_double2(old.x - acc[mi][ni][0], old.y - acc[mi][ni][1]);
            }
    };
    gemm_nt_jobs<true>(blockIdx.x, gridDim.x, njobs, jobfn, epifn, smem, &pipe, &mapL);
}

// ------------------------------------------------------------------------------------------------------------
// Trailing update, half-tile variant: the same rank-128k SYRK with 128 x 64 output tiles, 4 consumer warps + 1 TMA
// producer warp per CTA, a 2-stage 48 KB ring -- so TWO CTAs fit on an SM and one CTA's read-modify-write epilogue
// runs underneath the other CTA's DMMA stream (with one 128 x 128 CTA per SM the pipe idled ~19 % of the time during
// the epilogues; profiles/r01_trailing_v1_ncu_key_metrics.txt).  Jobs come from a global atomic queue; a CTA that
// lands on the reserved SM exits at once, which keeps that SM free for the look-ahead diagonal-tile kernel.
//   mode 0: lower triangle of tiles base <= j <= i < T, two column halves each;  mode 1: tile column j = base.
// ------------------------------------------------------------------------------------------------------------
constexpr int H_BN = 64;
constexpr int H_CONS = 128;                       // 4 consumer warps: 2 (m) x 2 (n), warp tile 64 x 32
constexpr int H_THREADS = H_CONS + 32;
constexpr int H_NSTAGE = 2;
constexpr int H_STAGE_DOUBLES = (TS + H_BN) * BK;
constexpr int H_SMEM_BYTES = H_NSTAGE * H_STAGE_DOUBLES * 8;   // 96 KB

__device__ __forceinline__ unsigned smid() {
    unsigned r;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
    return r;
}

__global__ void __launch_bounds__(H_THREADS, 2) potrf_trailing64_kernel(const __grid_constant__ CUtensorMap mapL,
                                                                       const __grid_constant__ CUtensorMap mapL64,
                                                                       double* __restrict__ Lbuf, long long ld, int k0, int k1, int base,
                                                                       int mode, int njobs, int* __restrict__ job_counter,
                                                                       int reserved_sm, int job_skip) {
    extern __shared__ __align__(128) double smem[];
    __shared__ __align__(8) uint64_t full[H_NSTAGE], empty[H_NSTAGE], slot_full[2], slot_empty[2];
    __shared__ int job_slot[2];
    // SMs reserved_sm .. #SM-1 stay free for the spine stream: a CTA that lands there leaves at once and the queue hands its jobs to the
    // others -- unless it is the last one that could (every other CTA of the grid left the same way: the rest of the GPU was busy with
    // other streams' or other handles' kernels); job_counter[1] counts the leavers.
    if (reserved_sm >= 0 && (int)smid() >= reserved_sm) {
        __shared__ int leave;
        if (threadIdx.x == 0) leave = atomicAdd(job_counter + 1, 1) < (int)gridDim.x - 1;
        __syncthreads();
        if (leave) return;
    }
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < H_NSTAGE; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], H_CONS / 32); }
        for (int i = 0; i < 2; ++i) { mbar_init(&slot_full[i], 1); mbar_init(&slot_empty[i], H_CONS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    auto decode = [&](int job, int& ti, int& tj, int& half) {
        job += job_skip;                               // 2: both halves of tile (base, base) belong to the spine kernel
        half = job & 1;
        const int tjob = job >> 1;
        if (mode == 0) {
            int a, b;
            tri_decode(tjob, a, b);
            ti = base + a;
            tj = base + b;
        } else {
            ti = base + tjob;
            tj = base;
        }
    };
    const int nslab = (k1 - k0) * SLABS_PER_TILE, s0 = k0 * SLABS_PER_TILE;
    if (warp == H_CONS / 32) {
        // ---------------- producer ----------------
        if (lane == 0) {
            int gs = 0;
            for (int lt = 0;; ++lt) {
                if (lt >= 2) mbar_wait(&slot_empty[lt & 1], ((lt >> 1) - 1) & 1);
                int job = atomicAdd(job_counter, 1);
                if (job >= njobs) job = -1;
                job_slot[lt & 1] = job;
                mbar_arrive(&slot_full[lt & 1]);
                if (job < 0) break;
                int ti, tj, half;
                decode(job, ti, tj, half);
                const int brow = tj * TS + half * H_BN;
                for (int it = 0; it < nslab + 2; ++it, ++gs) {
                    const int st = gs % H_NSTAGE;
                    if (gs >= H_NSTAGE) mbar_wait(&empty[st], ((gs / H_NSTAGE) - 1) & 1);
                    double* sA = smem + st * H_STAGE_DOUBLES;
                    double* sB = sA + SLAB_DOUBLES;
                    if (it < nslab) {
                        const int kel = (s0 + it) * BK;
                        mbar_expect_tx(&full[st], H_STAGE_DOUBLES * 8);
                        tma_load_3d(sA, &mapL, 0, ti * TS, kel >> 2, &full[st]);
                        tma_load_3d(sB, &mapL64, 0, brow, kel >> 2, &full[st]);
                    } else {
                        // old values of the output half tile: 32 columns per ring entry, staged like an A slab
                        mbar_expect_tx(&full[st], SLAB_DOUBLES * 8);
                        tma_load_3d(sA, &mapL, 0, ti * TS, (brow + 32 * (it - nslab)) >> 2, &full[st]);
                    }
                }
            }
        }
        return;
    }
    // ---------------- consumers ----------------
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
    int gs = 0;
    for (int lt = 0;; ++lt) {
        mbar_wait(&slot_full[lt & 1], (lt >> 1) & 1);
        const int job = job_slot[lt & 1];
        __syncwarp();
        if (lane == 0) mbar_arrive(&slot_empty[lt & 1]);
        if (job < 0) break;
        int ti, tj, half;
        decode(job, ti, tj, half);
        double acc[8][4][2];
        acc_clear(acc);
        for (int it = 0; it < nslab; ++it, ++gs) {
            const int st = gs % H_NSTAGE;
            mbar_wait(&full[st], (gs / H_NSTAGE) & 1);
            const double* sA = smem + st * H_STAGE_DOUBLES;
            const double* sB = sA + SLAB_DOUBLES;
            const double* pa = sA + ((wm * 64 + g) << 2) + t;
            const double* pbf = sB + ((wn * 32 + g) << 2) + t;
#pragma unroll
            for (int kg = 0; kg < BK / 4; ++kg) {
                double a[8], b[4];
#pragma unroll
                for (int mi = 0; mi < 8; ++mi) a[mi] = pa[(kg * TS + mi * 8) << 2];
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) b[ni] = pbf[(kg * H_BN + ni * 8) << 2];
#pragma unroll
                for (int mi = 0; mi < 8; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
        }
        double* out = Lbuf + (long long)ti * TS * ld + (long long)tj * TS + half * H_BN;
        for (int e = 0; e < 2; ++e, ++gs) {
            const int st = gs % H_NSTAGE;
            mbar_wait(&full[st], (gs / H_NSTAGE) & 1);
            if (wn == e) {
                const double* box = smem + st * H_STAGE_DOUBLES;      // [c/4][row][c%4], 32 columns of this warp
#pragma unroll
                for (int mi = 0; mi < 8; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) {
                        const int r = wm * 64 + mi * 8 + g;
                        const int kg = ni * 2 + (t >> 1), off = (2 * t) & 3;
                        const double2 old = *reinterpret_cast<const double2*>(box + (((kg * TS) + r) << 2) + off);
                        *reinterpret_cast<double2*>(out + (long long)r * ld + wn * 32 + ni * 8 + 2 * t) =
                            make_double2(old.x - acc[mi][ni][0], old.y - acc[mi][ni][1]);
                    }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Spine step: the two tiles between one diagonal tile and the next,
//     L[k+1,k] = A[k+1,k] Dinv_k^T          (+ its mirror, + the forward-substitution update y_{k+1} -= L[k+1,k] z_k)
//     A[k+1,k+1] -= L[k+1,k] L[k+1,k]^T     (lower 32-blocks)
// as ONE launch of eight CTAs.  As single 128^3 tiles of the panel and look-ahead launches these two products cost
// 2 x 17 us of DMMA time on one SM each (plus two launches) on the chain  diag(k) -> panel -> column -> diag(k+1), i.e. more than the
// diagonal tile itself; here each CTA owns a 64 x 32 block of the output (rank = 4 * row half + column quarter), stages its
// 64 + 32 operand rows in shared memory and runs the small DMMA product of the diagonal-tile kernel (smem_gemm), the second
// product reading the first one's result back from L2 after a barrier over the eight CTAs (a counter in global memory: a cluster
// would have to find its eight SMs inside one GPC, and the SMs the trailing kernel leaves free are not chosen by GPC).  The spine stream is then
// diag(k) -> spine(k) -> diag(k+1); the wide panel / trailing kernels only feed it (factorize_device).
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int ld_acquire_gpu_s32(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_s32(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

constexpr int SPINE_CTAS = 8;
constexpr int SP_LDC = 33, SP_LDY = 9;
constexpr int SPINE_SMEM_BYTES = (64 * DLD + 32 * DLD + 64 * SP_LDC + 8 * DLD + 64 * SP_LDY) * 8;   // 131,328 B

// rows x 128 doubles from global memory (leading dimension ld, read through L2) into shared memory rows of DLD doubles
__device__ __forceinline__ void spine_stage_rows(double* dst, const double* src, long long ld, int rows) {
    const int n2 = rows * 64;
    for (int e0 = threadIdx.x; e0 < n2; e0 += 256 * 8) {
        double2 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int e = e0 + u * 256;
            if (e < n2) v[u] = __ldcg(reinterpret_cast<const double2*>(src + (long long)(e >> 6) * ld + ((e & 63) << 1)));
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int e = e0 + u * 256;
            if (e < n2) *reinterpret_cast<double2*>(dst + (e >> 6) * DLD + ((e & 63) << 1)) = v[u];
        }
    }
}

__global__ void __launch_bounds__(256, 1)
    potrf_spine_kernel(double* Lbuf, long long ld, int kt, const double* __restrict__ dinv, double* rhs, const double* __restrict__ sol, int Npad,
                       int p, int* arrivals, int arrivals_target) {
    extern __shared__ __align__(16) double sp[];
    double* As = sp;                          // [64][DLD]
    double* Bs = As + 64 * DLD;               // [32][DLD]
    double* Cs = Bs + 32 * DLD;               // [64][SP_LDC]
    double* Zs = Cs + 64 * SP_LDC;            // [8][DLD]
    double* Cy = Zs + 8 * DLD;                // [64][SP_LDY]
    const int tid = threadIdx.x;
    const int rank = blockIdx.x % SPINE_CTAS, h = rank >> 2, q = rank & 3;
    double* panel = Lbuf + ((long long)(kt + 1) * TS) * ld + (long long)kt * TS;            // tile (k+1, k)
    double* mirror = Lbuf + ((long long)kt * TS) * ld + (long long)(kt + 1) * TS;           // tile (k, k+1)
    double* next = Lbuf + ((long long)(kt + 1) * TS) * ld + (long long)(kt + 1) * TS;       // tile (k+1, k+1)
    // ---- product 1: P[64h.., 32q..] = A[64h.., :] Dinv[32q.., :]^T; Dinv row c is zero beyond column c -------------------------
    spine_stage_rows(As, panel + (long long)(64 * h) * ld, ld, 64);
    spine_stage_rows(Bs, dinv + (long long)kt * TS * TS + (long long)(32 * q) * TS, TS, 32);
    __syncthreads();
    if (tid == 0) atomicAdd(arrivals, 1);           // "my rows of A are staged": the product is in place, nobody may store before all have read
    smem_gemm<false, false>(Cs, SP_LDC, As, DLD, Bs, DLD, 64, 32, 32 * (q + 1), 1.0, 0.0, false);
    if (tid == 0)
        while (ld_acquire_gpu_s32(arrivals) < arrivals_target - SPINE_CTAS) {}
    __syncthreads();
    for (int e = tid; e < 64 * 32; e += 256) {
        const int r = e >> 5, c = e & 31;
        panel[(long long)(64 * h + r) * ld + 32 * q + c] = Cs[r * SP_LDC + c];
    }
    for (int e = tid; e < 64 * 32; e += 256) {
        const int r = e & 63, c = e >> 6;
        mirror[(long long)(32 * q + c) * ld + 64 * h + r] = Cs[r * SP_LDC + c];
    }
    // all eight CTAs meet again (a cooperative launch: they are on the machine together; the trailing kernel keeps SMs free for them)
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        atomicAdd(arrivals, 1);
        while (ld_acquire_gpu_s32(arrivals) < arrivals_target) {}
    }
    __syncthreads();
    // ---- product 2: A'[64h.., 32q..] -= P[64h.., :] P[32q.., :]^T (blocks entirely above the diagonal are never read) -----------
    const bool lower = !(h == 0 && q >= 2);
    const bool fwd = (q == 0) && (rhs != nullptr);
    if (!lower && !fwd) return;
    spine_stage_rows(As, panel + (long long)(64 * h) * ld, ld, 64);
    if (lower) {
        spine_stage_rows(Bs, panel + (long long)(32 * q) * ld, ld, 32);
        for (int e = tid; e < 64 * 32; e += 256) {
            const int r = e >> 5, c = e & 31;
            Cs[r * SP_LDC + c] = __ldcg(next + (long long)(64 * h + r) * ld + 32 * q + c);
        }
    }
    if (fwd)
        for (int e = tid; e < 8 * TS; e += 256) Zs[(e >> 7) * DLD + (e & 127)] = ((e >> 7) < p) ? sol[(long long)(e >> 7) * Npad + kt * TS + (e & 127)] : 0.0;
    __syncthreads();
    if (lower) smem_gemm<false, false>(Cs, SP_LDC, As, DLD, Bs, DLD, 64, 32, TS, -1.0, 1.0, false);
    if (fwd) smem_gemm<false, false>(Cy, SP_LDY, As, DLD, Zs, DLD, 64, 8, TS, 1.0, 0.0, false);
    __syncthreads();
    if (lower)
        for (int e = tid; e < 64 * 32; e += 256) {
            const int r = e >> 5, c = e & 31;
            next[(long long)(64 * h + r) * ld + 32 * q + c] = Cs[r * SP_LDC + c];
        }
    if (fwd && tid < 64 * p) {
        const int r = tid & 63, qq = tid >> 6;
        rhs[(long long)qq * Npad + (kt + 1) * TS + 64 * h + r] -= Cy[r * SP_LDY + qq];
    }
}

// ------------------------------------------------------------------------------------------------------------
// Back substitution alpha = L^-T z (the forward half z = L^-1 y is fused into the factorisation above), block by block
// from the last tile:  step k:  alpha_k = Dinv_k^T z_k
//                               z_j -= L[k,j]^T alpha_k  for j < k   (one CTA per tile j; reads the 128x128 tile of the
//                               upper mirror = rows of tile j, columns of tile k, K-contiguous).
// Few right-hand sides: memory-bound (each step touches k tiles of 128 KB once).
// ------------------------------------------------------------------------------------------------------------
// One launch per step: every CTA first forms alpha_k = Dinv_k^T z_k (redundantly; the 128 KB block is L2-resident and
// the 256 threads issue all of its loads up front), CTA 0 publishes it, then CTA j applies its tile's update.
__global__ void __launch_bounds__(256) trsv_back_step_kernel(const double* __restrict__ Lbuf, long long ld, const double* __restrict__ dinv,
                                                             double* __restrict__ rhs, double* __restrict__ sol, int Npad, int p, int kt) {
    __shared__ double zs[MAXP][TS];
    __shared__ double part[8][MAXP][TS];
    __shared__ double al[MAXP][TS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int e = tid; e < p * TS; e += 256) zs[e / TS][e % TS] = rhs[(long long)(e / TS) * Npad + kt * TS + (e % TS)];
    const double* dk = dinv + (long long)kt * TS * TS;
    // alpha[c] = sum_r Dinv[r][c] z[r]: warp w takes rows 16w..16w+15, lane takes columns 4*lane..4*lane+3
    double2 d0[16], d1[16];
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) {
        const double* rp = dk + (warp * 16 + rr) * TS + lane * 4;
        d0[rr] = *reinterpret_cast<const double2*>(rp);
        d1[rr] = *reinterpret_cast<const double2*>(rp + 2);
    }
    __syncthreads();
    for (int q = 0; q < p; ++q) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
        for (int rr = 0; rr < 16; ++rr) {
            const double zv = zs[q][warp * 16 + rr];
            a0 = fma(d0[rr].x, zv, a0);
            a1 = fma(d0[rr].y, zv, a1);
            a2 = fma(d1[rr].x, zv, a2);
            a3 = fma(d1[rr].y, zv, a3);
        }
        part[warp][q][lane * 4 + 0] = a0;
        part[warp][q][lane * 4 + 1] = a1;
        part[warp][q][lane * 4 + 2] = a2;
        part[warp][q][lane * 4 + 3] = a3;
    }
    __syncthreads();
    for (int e = tid; e < p * TS; e += 256) {
        const int q = e / TS, c = e % TS;
        double sacc = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) sacc += part[w][q][c];
        al[q][c] = sacc;
        if (blockIdx.x == 0) sol[(long long)q * Npad + kt * TS + c] = sacc;
    }
    __syncthreads();
    const int tj = blockIdx.x;                        // tiles above the diagonal block row: j < kt
    if (tj >= kt) return;
    double a[MAXP][4];
    for (int q = 0; q < MAXP; ++q)
        for (int e = 0; e < 4; ++e) a[q][e] = (q < p) ? al[q][lane * 4 + e] : 0.0;
    for (int rr = 0; rr < 16; ++rr) {
        const int r = warp * 16 + rr;
        const double* rowp = Lbuf + ((long long)tj * TS + r) * ld + (long long)kt * TS + lane * 4;   // mirror: L[k,j]^T rows
        const double2 u0 = *reinterpret_cast<const double2*>(rowp);
        const double2 u1 = *reinterpret_cast<const double2*>(rowp + 2);
        for (int q = 0; q < p; ++q) {
            double sacc = u0.x * a[q][0];
            sacc = fma(u0.y, a[q][1], sacc);
            sacc = fma(u1.x, a[q][2], sacc);
            sacc = fma(u1.y, a[q][3], sacc);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, off);
            if (lane == 0) rhs[(long long)q * Npad + tj * TS + r] -= sacc;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// The same back substitution as ONE launch: CTA b owns block row j = T-1-b, keeps z_j and Dinv_j in shared memory and follows the
// chain  x_{T-1} -> x_{T-2} -> ...  through per-block flags in global memory (flag[k] == epoch: x_k is published).  For every k > j
// it prefetches the tile L[k,j] (lower triangle: rows of block k, contiguous in the columns of block j) into registers BEFORE
// spinning on flag[k], so on the chain's critical path a step costs: flag -> 4 KB of x_k -> 192 DFMA per thread -> reduce ->
// Dinv_j^T z_j from shared memory -> publish.  That is 3-4 us per block against 22-24 us for one launch per block (N = 4096:
// 0.7 ms of a 5.1 ms fit).  Producers have the LOWER block indices, so with in-order CTA dispatch a spinning CTA never keeps its
// producer off the machine (the decoupled-look-back argument), whatever T is.  Summation order is fixed: deterministic.
// ------------------------------------------------------------------------------------------------------------
constexpr int BACKCHAIN_SMEM_BYTES = TS * TS * (int)sizeof(double);

__global__ void __launch_bounds__(256, 1) trsv_back_chain_kernel(const double* __restrict__ Lbuf, long long ld, const double* __restrict__ dinv,
                                                                 const double* __restrict__ rhs, double* sol, int Npad, int p, int T, int* flags,
                                                                 int epoch) {
    extern __shared__ __align__(16) double dsm[];            // Dinv_j, row-major [r][c]
    __shared__ double zs[MAXP][TS];
    __shared__ double xs[MAXP][TS];
    __shared__ double part[8][MAXP][TS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int j = T - 1 - (int)blockIdx.x;
    {
        const double2* src = reinterpret_cast<const double2*>(dinv + (long long)j * TS * TS);
        double2* dst = reinterpret_cast<double2*>(dsm);
#pragma unroll 8
        for (int e = tid; e < TS * TS / 2; e += 256) dst[e] = src[e];
        for (int e = tid; e < MAXP * TS; e += 256) zs[e / TS][e % TS] = (e / TS < p) ? rhs[(long long)(e / TS) * Npad + j * TS + (e % TS)] : 0.0;
    }
    __syncthreads();
    for (int k = T - 1; k > j; --k) {
        // thread (warp, lane): rows c = 16 warp .. 16 warp + 15 of block k, columns r = 4 lane .. 4 lane + 3 of block j
        double2 u0[16], u1[16];
#pragma unroll
        for (int rr = 0; rr < 16; ++rr) {
            const double* rowp = Lbuf + ((long long)k * TS + warp * 16 + rr) * ld + (long long)j * TS + lane * 4;
            u0[rr] = *reinterpret_cast<const double2*>(rowp);
            u1[rr] = *reinterpret_cast<const double2*>(rowp + 2);
        }
        if (tid == 0)
            while (ld_acquire_gpu_s32(flags + k) != epoch) {}
        __syncthreads();
        for (int e = tid; e < p * TS; e += 256) xs[e / TS][e % TS] = __ldcg(sol + (long long)(e / TS) * Npad + k * TS + (e % TS));
        __syncthreads();
        for (int q = 0; q < p; ++q) {
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
            for (int rr = 0; rr < 16; ++rr) {
                const double xv = xs[q][warp * 16 + rr];
                a0 = fma(u0[rr].x, xv, a0);
                a1 = fma(u0[rr].y, xv, a1);
                a2 = fma(u1[rr].x, xv, a2);
                a3 = fma(u1[rr].y, xv, a3);
            }
            part[warp][q][lane * 4 + 0] = a0;
            part[warp][q][lane * 4 + 1] = a1;
            part[warp][q][lane * 4 + 2] = a2;
            part[warp][q][lane * 4 + 3] = a3;
        }
        __syncthreads();
        for (int e = tid; e < p * TS; e += 256) {
            const int q = e / TS, c = e % TS;
            double sacc = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) sacc += part[w][q][c];
            zs[q][c] -= sacc;
        }
        // the two barriers of the next round (flag, x_k) separate these reads of part[] from its next writes
    }
    __syncthreads();
    // x_j[c] = sum_r Dinv_j[r][c] z_j[r]
    for (int q = 0; q < p; ++q) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
        for (int rr = 0; rr < 16; ++rr) {
            const int r = warp * 16 + rr;
            const double zv = zs[q][r];
            const double2 d0 = *reinterpret_cast<const double2*>(dsm + r * TS + lane * 4);
            const double2 d1 = *reinterpret_cast<const double2*>(dsm + r * TS + lane * 4 + 2);
            a0 = fma(d0.x, zv, a0);
            a1 = fma(d0.y, zv, a1);
            a2 = fma(d1.x, zv, a2);
            a3 = fma(d1.y, zv, a3);
        }
        part[warp][q][lane * 4 + 0] = a0;
        part[warp][q][lane * 4 + 1] = a1;
        part[warp][q][lane * 4 + 2] = a2;
        part[warp][q][lane * 4 + 3] = a3;
    }
    __syncthreads();
    for (int e = tid; e < p * TS; e += 256) {
        const int q = e / TS, c = e % TS;
        double sacc = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) sacc += part[w][q][c];
        sol[(long long)q * Npad + j * TS + c] = sacc;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) st_release_gpu_s32(flags + j, epoch);
}

// ------------------------------------------------------------------------------------------------------------
// LML scalar terms (sklearn:_gpr.py:613-617): out[0] = sum_q y_q . alpha_q ; out[1] = sum_i log L_ii.  One CTA,
// fixed summation order (deterministic).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) lml_terms_kernel(const double* __restrict__ Y, const double* __restrict__ alpha,
                                                         const double* __restrict__ Lbuf, long long ld, int N, int Npad, int p,
                                                         double* __restrict__ out) {
    __shared__ double red[2][1024];
    double s0 = 0.0, s1 = 0.0;
    for (int n = threadIdx.x; n < N; n += 1024) {
        for (int q = 0; q < p; ++q) s0 = fma(Y[(long long)q * Npad + n], alpha[(long long)q * Npad + n], s0);
        s1 += log(Lbuf[(long long)n * ld + n]);
    }
    red[0][threadIdx.x] = s0;
    red[1][threadIdx.x] = s1;
    __syncthreads();
    for (int w = 512; w > 0; w >>= 1) {
        if (threadIdx.x < w) {
            red[0][threadIdx.x] += red[0][threadIdx.x + w];
            red[1][threadIdx.x] += red[1][threadIdx.x + w];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = red[0][0]; out[1] = red[1][0]; }
}

// ------------------------------------------------------------------------------------------------------------
// Triangular inverse by recursive doubling on the tile engine.  Minv: lower = L^-1, upper = mirror.
// init: copy the diagonal-tile inverses (Dinv) into Minv's diagonal tiles (lower = Dinv, upper = Dinv^T).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) trtri_init_kernel(double* __restrict__ Minv, long long ld, const double* __restrict__ dinv) {
    const int kt = blockIdx.x;
    const double* dk = dinv + (long long)kt * TS * TS;
    double* tile = Minv + (long long)kt * TS * ld + kt * TS;
    for (int e = threadIdx.x; e < TS * TS; e += 256) {
        int r = e >> 7, c = e & 127;
        tile[(long long)r * ld + c] = (c <= r) ? dk[r * TS + c] : dk[c * TS + r];
    }
}

// Level with half-size s tiles: diagonal super-blocks [2qs, 2qs+s) = "A" and [2qs+s, min(2qs+2s,T)) = "C".
// Product 1 (transposed temp):  W[n][m] = sum_{k in A, k >= n} Minv^T[n][k] * L[m][k],   n in A (rows), m in C (cols)
__global__ void __launch_bounds__(GEMM_THREADS, 1) trtri_level_p1_kernel(const __grid_constant__ CUtensorMap mapL,
                                                                        const __grid_constant__ CUtensorMap mapM,
                                                                        double* __restrict__ W, long long ld, int T, int s) {
    extern __shared__ __align__(128) double smem[];
    __shared__ PipeBarriers pipe;
    pipe_init(&pipe);
    // decode blockIdx.x -> (pair q, tile n in A, tile m in C); enumerate with per-pair stride s*s (skip out-of-range)
    const int per = s * s;
    const int q = blockIdx.x / per, rem = blockIdx.x % per;
    const int tn = 2 * q * s + rem / s;          // A tile (row of W)
    const int tm = 2 * q * s + s + rem % s;      // C tile (col of W)
    if (tm >= T) return;
    const int a_end = 2 * q * s + s;             // k-tiles tn .. a_end-1
    double acc[8][4][2];
    acc_clear(acc);
    Operand A{&mapM, tn * TS, 0, MASK_UPPER, tn};   // mirror rows n: Minv^T[n][k], k >= n
    Operand B{&mapL, tm * TS, 0, MASK_NONE, -1};    // L rows m (below the diagonal for k in A)
    gemm_nt_tile(A, B, tn, a_end, acc, smem, &pipe);
    double* out = W + (long long)tn * TS * ld + (long long)tm * TS;
    acc_foreach(acc, [&](int r, int c, double v0, double v1) {
        *reinterpret_cast<double2*>(out + (long long)r * ld + c) = make_double2(v0, v1);
    });
}

// Product 2:  Minv[m][n] = - sum_{k in C, k <= m} Minv[m][k] * W[n][k],  m in C (rows), n in A (cols); plus mirror.
__global__ void __launch_bounds__(GEMM_THREADS, 1) trtri_level_p2_kernel(const __grid_constant__ CUtensorMap mapM,
                                                                        const __grid_constant__ CUtensorMap mapW,
                                                                        double* __restrict__ Minv, long long ld, int T, int s) {
    extern __shared__ __align__(128) double smem[];
    __shared__ PipeBarriers pipe;
    pipe_init(&pipe);
    const int per = s * s;
    const int q = blockIdx.x / per, rem = blockIdx.x % per;
    const int tm = 2 * q * s + s + rem / s;      // C tile (row of Minv)
    const int tn = 2 * q * s + rem % s;          // A tile (col of Minv)
    if (tm >= T) return;
    const int c_begin = 2 * q * s + s;           // k-tiles c_begin .. tm
    double acc[8][4][2];
    acc_clear(acc);
    Operand A{&mapM, tm * TS, 0, MASK_LOWER, tm};   // Minv rows m, k <= m
    Operand B{&mapW, tn * TS, 0, MASK_NONE, -1};    // W rows n, cols k in C
    gemm_nt_tile(A, B, c_begin, tm + 1, acc, smem, &pipe);
    double* out = Minv + (long long)tm * TS * ld + (long long)tn * TS;
    double* outT = Minv + (long long)tn * TS * ld + (long long)tm * TS;
    acc_foreach(acc, [&](int r, int c, double v0, double v1) {
        *reinterpret_cast<double2*>(out + (long long)r * ld + c) = make_double2(-v0, -v1);
        outT[(long long)c * ld + r] = -v0;
        outT[(long long)(c + 1) * ld + r] = -v1;
    });
}

// K^-1 = L^-T L^-1:  Kinv[i][j] = sum_{k >= max(i,j)} Minv^T[i][k] Minv^T[j][k]; lower tiles (ti >= tj) + mirror.
__global__ void __launch_bounds__(GEMM_THREADS, 1) kinv_kernel(const __grid_constant__ CUtensorMap mapM, double* __restrict__ Kinv,
                                                              long long ld, int T) {
    extern __shared__ __align__(128) double smem[];
    __shared__ PipeBarriers pipe;
    pipe_init(&pipe);
    int ti, tj;
    tri_decode(blockIdx.x, ti, tj);
    double acc[8][4][2];
    acc_clear(acc);
    Operand A{&mapM, ti * TS, 0, MASK_UPPER, ti};
    Operand B{&mapM, tj * TS, 0, MASK_UPPER, tj};   // only bites when tj == ti
    gemm_nt_tile(A, B, ti, T, acc, smem, &pipe);
    double* out = Kinv + (long long)ti * TS * ld + (long long)tj * TS;
    double* outT = Kinv + (long long)tj * TS * ld + (long long)ti * TS;
    const bool diag = (ti == tj);
    acc_foreach(acc, [&](int r, int c, double v0, double v1) {
        *reinterpret_cast<double2*>(out + (long long)r * ld + c) = make_double2(v0, v1);
        if (!diag) {
            outT[(long long)c * ld + r] = v0;
            outT[(long long)(c + 1) * ld + r] = v1;
        }
    });
}

// ------------------------------------------------------------------------------------------------------------
// Fused LML-gradient pass (sklearn:_gpr.py:629-651): one sweep over the lower tiles of K^-1 that regenerates
// K_ij and the per-dimension scaled squared distances and reduces
//   g_c    = 1/2 sum_ij W_ij cR_ij,   g_ell_a = 1/2 sum_ij W_ij cR_ij (dx_a/ell_a)^2,   g_s2 = 1/2 s2 sum_i W_ii
// with W_ij = sum_q alpha_iq alpha_jq - p Kinv_ij.  Off-diagonal tiles count twice (symmetry).
// Per-CTA partials are written to part[blockIdx.x][2+D]; lml_grad_reduce sums them in a fixed order.
// Roofline: reads 8*N^2/2 bytes of K^-1 once (HBM) against N^2/2 exps -> exp-bound on B200.
// ------------------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) lml_grad_kernel(const double* __restrict__ Xs, const double* __restrict__ alpha,
                                                       const double* __restrict__ Kinv, int N, int Npad, int p, KParams kp,
                                                       double* __restrict__ part) {
    int ti, tj;
    tri_decode(blockIdx.x, ti, tj);
    __shared__ double xi[D][TS];
    __shared__ double ai[MAXP][TS];
    __shared__ double red[8][2 + D];
    const int tid = threadIdx.x;
    for (int e = tid; e < D * TS; e += 256) xi[e / TS][e % TS] = Xs[(long long)(e / TS) * Npad + ti * TS + (e % TS)];
    for (int e = tid; e < p * TS; e += 256) ai[e / TS][e % TS] = alpha[(long long)(e / TS) * Npad + ti * TS + (e % TS)];
    __syncthreads();
    const int cg = (tid & 31) * 4, r0 = tid >> 5;
    double xj[D][4], aj[MAXP][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int a = 0; a < D; ++a) xj[a][q] = Xs[(long long)a * Npad + tj * TS + cg + q];
        for (int o = 0; o < MAXP; ++o) aj[o][q] = (o < p) ? alpha[(long long)o * Npad + tj * TS + cg + q] : 0.0;
    }
    double gc = 0.0, gs = 0.0, gl[D];
#pragma unroll
    for (int a = 0; a < D; ++a) gl[a] = 0.0;
    const double wsym = (ti == tj) ? 1.0 : 2.0;
    for (int rr = 0; rr < 16; ++rr) {
        int r = r0 + rr * 8;
        int gi = ti * TS + r;
        double4 kv = *reinterpret_cast<const double4*>(Kinv + (long long)gi * Npad + tj * TS + cg);
        double kin[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int gj = tj * TS + cg + q;
            if (gi < N && gj < N) {
                double aa = 0.0;
                for (int o = 0; o < p; ++o) aa = fma(ai[o][r], aj[o][q], aa);
                double w = aa - (double)p * kin[q];
                double s = 0.0, d2[D];
#pragma unroll
                for (int a = 0; a < D; ++a) {
                    double df = xi[a][r] - xj[a][q];
                    d2[a] = df * df;
                    s += d2[a];
                }
                gc += kp.c * kernel_profile(s, kp.kind, [](double z) { return exp(z); }) * w;
                const double kg = kp.c * kernel_grad_factor(s, kp.kind) * w;
#pragma unroll
                for (int a = 0; a < D; ++a) gl[a] = fma(kg, d2[a], gl[a]);
                if (gi == gj) gs += w;
            }
        }
    }
    // block reduction: warp shuffle then fixed-order sum over the 8 warps
    double vals[2 + D];
    vals[0] = gc; vals[1] = gs;
#pragma unroll
    for (int a = 0; a < D; ++a) vals[2 + a] = gl[a];
#pragma unroll
    for (int v = 0; v < 2 + D; ++v) {
        double x = vals[v];
        for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
        if ((tid & 31) == 0) red[tid >> 5][v] = x;
    }
    __syncthreads();
    if (tid < 2 + D) {
        double x = 0.0;
        for (int w = 0; w < 8; ++w) x += red[w][tid];
        // gs is a diagonal quantity: no symmetry factor
        part[(long long)blockIdx.x * (2 + MAXD) + tid] = (tid == 1) ? x : wsym * x;
    }
}

__global__ void __launch_bounds__(256) lml_grad_reduce_kernel(const double* __restrict__ part, long long nparts, int d, KParams kp,
                                                              double* __restrict__ grad) {
    // grad layout: [g_c, g_ell_0..g_ell_{d-1}, g_s2]
    __shared__ double red[256];
    for (int v = 0; v < 2 + d; ++v) {
        double s = 0.0;
        for (long long i = threadIdx.x; i < nparts; i += 256) s += part[i * (2 + MAXD) + v];
        red[threadIdx.x] = s;
        __syncthreads();
        for (int w = 128; w > 0; w >>= 1) {
            if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            double tot = red[0];
            if (v == 0) grad[0] = 0.5 * tot;
            else if (v == 1) grad[1 + d] = 0.5 * kp.s2 * tot;
            else grad[1 + (v - 2)] = 0.5 * tot;
        }
        __syncthreads();
    }
}


// ------------------------------------------------------------------------------------------------------------
// Rank-1 append of one training point at fixed hyper-parameters (the greedy loop of gaussian_process_al.py:26-55 adds one point per
// iteration; with fixed hyper-parameters its re-fit is exactly this update).  With M = L^-1 (lower + mirror) and the new point's
// kernel vector k:   l = M k,  dd = sqrt(k** - l.l),  new row of L = [l, dd],  new row of M = [r, 1/dd] with r = -(l^T M)/dd,
// alpha <- [alpha + r^T z_n ; z_n/dd] with z_n = r.Y + y_n/dd.   O(N^2) memory-bound work instead of the O(N^3) re-factorisation.
// ------------------------------------------------------------------------------------------------------------
struct NewPoint {
    double x[MAXD], xs[MAXD], y[MAXP];
};

__global__ void __launch_bounds__(256) append_kvec_kernel(const double* __restrict__ Xs, int N, int Npad, int d, NewPoint np, KParams kp,
                                                          double* __restrict__ kv) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Npad) return;
    double v = 0.0;
    if (j < N) {
        double sq = 0.0;
        for (int a = 0; a < d; ++a) {
            const double df = np.xs[a] - Xs[(long long)a * Npad + j];
            sq = fma(df, df, sq);
        }
        v = kp.c * kernel_profile(sq, kp.kind, [](double z) { return exp(z); });
    }
    kv[j] = v;
}

// out[i] = sum_{j <= i} M[i][j] v[j], i < N (one warp per row of the lower triangle)
__global__ void __launch_bounds__(256) gemv_lower_rows_kernel(const double* __restrict__ M, long long ld, int N, const double* __restrict__ v,
                                                              double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= N) return;
    const double* row = M + (long long)i * ld;
    double s = 0.0;
    for (int j = lane; j <= i; j += 32) s = fma(row[j], v[j], s);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) out[i] = s;
}

// scal[0] = dd = sqrt(kss - l.l) (0 when not positive), fixed-order reduction
__global__ void __launch_bounds__(1024) append_pivot_kernel(const double* __restrict__ l, int N, double kss, double* __restrict__ scal) {
    __shared__ double red[1024];
    double s = 0.0;
    for (int i = threadIdx.x; i < N; i += 1024) s = fma(l[i], l[i], s);
    red[threadIdx.x] = s;
    __syncthreads();
    for (int w = 512; w > 0; w >>= 1) {
        if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double d2 = kss - red[0];
        scal[0] = d2 > 0.0 ? sqrt(d2) : 0.0;
    }
}

// r[j] = -(1/dd) (l[j] M[j][j] + sum_{j < i < N} l[i] M[i][j]); M[i][j] for i > j is read from the mirror (row j, column i): contiguous
__global__ void __launch_bounds__(256) gemv_mirror_rows_kernel(const double* __restrict__ M, long long ld, int N, const double* __restrict__ l,
                                                               const double* __restrict__ scal, double* __restrict__ r) {
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (j >= N) return;
    const double* row = M + (long long)j * ld;
    double s = 0.0;
    for (int i = j + lane; i < N; i += 32) s = fma(row[i], l[i], s);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) r[j] = -s / scal[0];
}

// write the new row / mirror column of L and M, the last diagonal tile's explicit inverse, the new training point and alpha
__global__ void __launch_bounds__(1024) append_write_kernel(double* __restrict__ Lbuf, double* __restrict__ Minv, double* __restrict__ dinv, long long ld,
                                                            int N, int Npad, int d, int p, const double* __restrict__ l, const double* __restrict__ r,
                                                            const double* __restrict__ scal, NewPoint np, double* __restrict__ X, double* __restrict__ Xs,
                                                            double* __restrict__ Y, double* __restrict__ alpha) {
    __shared__ double red[MAXP][1024];
    const int n = N;                                            // index of the new point
    const double dd = scal[0], idd = 1.0 / dd;
    double zacc[MAXP] = {0.0, 0.0, 0.0, 0.0};
    for (int j = threadIdx.x; j < N; j += 1024) {
        const double lj = l[j], rj = r[j];
        Lbuf[(long long)n * ld + j] = lj;
        Lbuf[(long long)j * ld + n] = lj;
        Minv[(long long)n * ld + j] = rj;
        Minv[(long long)j * ld + n] = rj;
        for (int q = 0; q < p; ++q) zacc[q] = fma(rj, Y[(long long)q * Npad + j], zacc[q]);
    }
    for (int q = 0; q < MAXP; ++q) red[q][threadIdx.x] = zacc[q];
    __syncthreads();
    for (int w = 512; w > 0; w >>= 1) {
        if (threadIdx.x < w)
            for (int q = 0; q < MAXP; ++q) red[q][threadIdx.x] += red[q][threadIdx.x + w];
        __syncthreads();
    }
    double zn[MAXP];
    for (int q = 0; q < MAXP; ++q) zn[q] = (q < p) ? red[q][0] + np.y[q] * idd : 0.0;
    for (int j = threadIdx.x; j < N; j += 1024)
        for (int q = 0; q < p; ++q) alpha[(long long)q * Npad + j] += r[j] * zn[q];
    const int tile0 = (n / TS) * TS, nl = n - tile0;
    double* dk = dinv + (long long)(n / TS) * TS * TS;
    for (int c = threadIdx.x; c < TS; c += 1024) dk[nl * TS + c] = (c < nl) ? r[tile0 + c] : (c == nl ? idd : 0.0);
    if (threadIdx.x == 0) {
        Lbuf[(long long)n * ld + n] = dd;
        Minv[(long long)n * ld + n] = idd;
        for (int a = 0; a < d; ++a) { X[(long long)a * Npad + n] = np.x[a]; Xs[(long long)a * Npad + n] = np.xs[a]; }
        for (int q = 0; q < p; ++q) { Y[(long long)q * Npad + n] = np.y[q]; alpha[(long long)q * Npad + n] = zn[q] * idd; }
    }
}

}  // namespace gptb
