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

struct KParams {
    double c, s2, jitter;
    double ell[MAXD];
    double inv_ell[MAXD];
};

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
            double v = kp.c * exp(-0.5 * s);
            if (gi == gj) v += diag_add;
            if (gi >= N || gj >= N) v = (gi == gj) ? 1.0 : 0.0;
            out[q] = v;
        }
        double4 o = make_double4(out[0], out[1], out[2], out[3]);
        *reinterpret_cast<double4*>(K + (long long)gi * Npad + tj * TS + cg) = o;
    }
}

// ------------------------------------------------------------------------------------------------------------
// Diagonal tile: in-shared-memory blocked Cholesky (16-wide left-looking panels) and triangular inverse.
// Writes L (lower) + L^T (upper) back into the factor buffer's diagonal tile and Dinv = L_kk^-1 (dense, zeros above
// the diagonal) into dinv.  info: first non-positive pivot (1-based global order), LAPACK dpotrf convention.
// ------------------------------------------------------------------------------------------------------------
constexpr int DLD = TS + 1;
constexpr int DIAG_SMEM_BYTES = (TS * DLD + TS) * 8;

__device__ __forceinline__ void chol_tile_smem(double* S, int tid, int nthreads, int gbase, int* info) {
    // S[r*DLD + c], lower triangle in/out
    for (int j0 = 0; j0 < TS; j0 += 16) {
        // (1) left-looking update of panel columns j0..j0+15 (rows j0..127) with columns 0..j0-1
        if (j0 > 0) {
            int cnt = (TS - j0) * 16;
            for (int e = tid; e < cnt; e += nthreads) {
                int r = j0 + (e >> 4), c = j0 + (e & 15);
                if (c <= r) {
                    const double* pr = S + r * DLD;
                    const double* pc = S + c * DLD;
                    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                    for (int m = 0; m < j0; m += 4) {
                        s0 = fma(pr[m], pc[m], s0);
                        s1 = fma(pr[m + 1], pc[m + 1], s1);
                        s2 = fma(pr[m + 2], pc[m + 2], s2);
                        s3 = fma(pr[m + 3], pc[m + 3], s3);
                    }
                    S[r * DLD + c] -= (s0 + s1) + (s2 + s3);
                }
            }
        }
        __syncthreads();
        // (2) warp 0 factors the 16x16 diagonal block
        if (tid < 32) {
            for (int j = 0; j < 16; ++j) {
                int jj = j0 + j;
                double dgl = S[jj * DLD + jj];
                if (tid == 0) {
                    if (!(dgl > 0.0)) {
                        if (atomicCAS(info, 0, gbase + jj + 1) == 0) {}
                    }
                }
                double piv = (dgl > 0.0) ? sqrt(dgl) : 1.0;
                __syncwarp();
                if (tid == 0) S[jj * DLD + jj] = piv;
                if (tid > j && tid < 16) S[(j0 + tid) * DLD + jj] /= piv;
                __syncwarp();
                // update remaining lower part of the 16x16 block: pairs (i,c), j < c <= i <= 15
                for (int e = tid; e < 256; e += 32) {
                    int i = e >> 4, c = e & 15;
                    if (c > j && c <= i) S[(j0 + i) * DLD + j0 + c] -= S[(j0 + i) * DLD + jj] * S[(j0 + c) * DLD + jj];
                }
                __syncwarp();
            }
        }
        __syncthreads();
        // (3) rows below the block: X = A * L11^-T, one thread per row
        {
            int r = j0 + 16 + tid;
            if (r < TS) {
                double x[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    double s = S[r * DLD + j0 + c];
#pragma unroll
                    for (int m = 0; m < 16; ++m)
                        if (m < c) s = fma(-x[m], S[(j0 + c) * DLD + j0 + m], s);
                    x[c] = s / S[(j0 + c) * DLD + j0 + c];
                }
#pragma unroll
                for (int c = 0; c < 16; ++c) S[r * DLD + j0 + c] = x[c];
            }
        }
        __syncthreads();
    }
}

// Z = L^-1 for the lower-triangular L in S (lower part). Z is built transposed in the strict upper part of S:
// S[c*DLD + r] = Z[r][c] for r > c; zd[r] = Z[r][r].  Thread c (< 128) owns column c.
__device__ __forceinline__ void trinv_tile_smem(double* S, double* zd, int tid) {
    if (tid < TS) zd[tid] = 1.0 / S[tid * DLD + tid];
    __syncthreads();
    if (tid < TS) {
        const int c = tid;
        double* zrow = S + c * DLD;   // Z[k][c] lives at zrow[k] for k > c
        const double zcc = zd[c];
        for (int r = 1; r < TS; ++r) {
            // all threads walk the same (r,k) so the L reads broadcast; threads with c >= r idle
            if (c < r) {
                const double* lr = S + r * DLD;
                double s0 = lr[c] * zcc, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                int k = c + 1;
                for (; k + 3 < r; k += 4) {
                    s0 = fma(lr[k], zrow[k], s0);
                    s1 = fma(lr[k + 1], zrow[k + 1], s1);
                    s2 = fma(lr[k + 2], zrow[k + 2], s2);
                    s3 = fma(lr[k + 3], zrow[k + 3], s3);
                }
                for (; k < r; ++k) s0 = fma(lr[k], zrow[k], s0);
                zrow[r] = -((s0 + s1) + (s2 + s3)) * zd[r];
            }
            // zrow[r] is only read by the same thread later: no barrier needed
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) potrf_diag_kernel(double* __restrict__ Lbuf, long long ld, int kt, double* __restrict__ dinv,
                                                         int* info) {
    extern __shared__ double sm[];
    double* S = sm;
    double* zd = sm + TS * DLD;
    const int tid = threadIdx.x;
    double* tile = Lbuf + (long long)kt * TS * ld + kt * TS;
    for (int e = tid; e < TS * TS; e += 256) {
        int r = e >> 7, c = e & 127;
        S[r * DLD + c] = tile[(long long)r * ld + c];
    }
    __syncthreads();
    chol_tile_smem(S, tid, 256, kt * TS, info);
    // write L (lower) and L^T (upper) back
    for (int e = tid; e < TS * TS; e += 256) {
        int r = e >> 7, c = e & 127;
        double v = (c <= r) ? S[r * DLD + c] : S[c * DLD + r];
        tile[(long long)r * ld + c] = v;
    }
    __syncthreads();
    trinv_tile_smem(S, zd, tid);
    double* dk = dinv + (long long)kt * TS * TS;
    for (int e = tid; e < TS * TS; e += 256) {
        int r = e >> 7, c = e & 127;
        double v = (c < r) ? S[c * DLD + r] : ((c == r) ? zd[r] : 0.0);
        dk[r * TS + c] = v;
    }
}

// ------------------------------------------------------------------------------------------------------------
// Panel: L[i,k] = A[i,k] * Dinv_k^T for the row tiles i > k (in place), plus the mirror L[i,k]^T into the upper half.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEMM_THREADS, 1) potrf_panel_kernel(const __grid_constant__ CUtensorMap mapL,
                                                                     const __grid_constant__ CUtensorMap mapD,
                                                                     double* __restrict__ Lbuf, long long ld, int kt) {
    extern __shared__ __align__(128) double smem[];
    __shared__ PipeBarriers pipe;
    pipe_init(&pipe);
    const int ti = kt + 1 + blockIdx.x;
    double acc[8][4][2];
    acc_clear(acc);
    // k-space of this product is the 128 columns of block column kt: offset k so that k-tile 0 is that block
    Operand A{&mapL, ti * TS, kt * TS, MASK_NONE, -1};
    Operand B{&mapD, kt * TS, 0, MASK_NONE, -1};   // Dinv rows n, k <= n (zeros stored above)
    gemm_nt_tile(A, B, 0, 1, acc, smem, &pipe);
    double* out = Lbuf + (long long)ti * TS * ld + (long long)kt * TS;          // lower: rows ti, cols kt
    double* outT = Lbuf + (long long)kt * TS * ld + (long long)ti * TS;         // mirror: rows kt, cols ti
    acc_foreach(acc, [&](int r, int c, double v0, double v1) {
        *reinterpret_cast<double2*>(out + (long long)r * ld + c) = make_double2(v0, v1);
        outT[(long long)c * ld + r] = v0;
        outT[(long long)(c + 1) * ld + r] = v1;
    });
}

// ------------------------------------------------------------------------------------------------------------
// Trailing update: A[i,j] -= sum_{kk in [k0,k1)} L[i,kk] L[j,kk]^T for k1 <= j <= i < T (lower tiles), i.e. a rank-(128*(k1-k0))
// SYRK on the DMMA engine.  blockIdx.x enumerates the lower triangle of the trailing tile grid, heaviest rows first
// is irrelevant here (all tiles cost the same).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEMM_THREADS, 1) potrf_trailing_kernel(const __grid_constant__ CUtensorMap mapL,
                                                                        double* __restrict__ Lbuf, long long ld, int k0, int k1,
                                                                        int first_tile) {
    extern __shared__ __align__(128) double smem[];
    __shared__ PipeBarriers pipe;
    pipe_init(&pipe);
    int a, b;
    tri_decode(blockIdx.x, a, b);
    const int ti = first_tile + a, tj = first_tile + b;
    double acc[8][4][2];
    acc_clear(acc);
    Operand A{&mapL, ti * TS, 0, MASK_NONE, -1};
    Operand B{&mapL, tj * TS, 0, MASK_NONE, -1};
    gemm_nt_tile(A, B, k0, k1, acc, smem, &pipe);
    double* out = Lbuf + (long long)ti * TS * ld + (long long)tj * TS;
    acc_foreach(acc, [&](int r, int c, double v0, double v1) {
        double2* p = reinterpret_cast<double2*>(out + (long long)r * ld + c);
        double2 o = *p;
        o.x -= v0;
        o.y -= v1;
        *p = o;
    });
}

// ------------------------------------------------------------------------------------------------------------
// Blocked triangular solve steps for alpha = L^-T L^-1 Y (few right-hand sides, memory-bound).
// Forward  (transposed = 0), step k: z_k = Dinv_k y_k ; y_i -= L[i,k] z_k for i > k     (rows of the lower half)
// Backward (transposed = 1), step k: a_k = Dinv_k^T z_k ; z_j -= L[k,j]^T a_k for j < k (rows of the upper mirror)
// rhs is SoA [p][Npad], updated in place; sol receives the solved block.  CTA b handles one other row tile
// (plus, for b == 0, publishing the solved block).  Every CTA recomputes the 128x128 block product (L2-resident).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) trsv_step_kernel(const double* __restrict__ Lbuf, long long ld, const double* __restrict__ dinv,
                                                        double* __restrict__ rhs, double* __restrict__ sol, int Npad, int p, int kt,
                                                        int transposed) {
    __shared__ double yk[MAXP][TS];
    __shared__ double zk[MAXP][TS];
    const int tid = threadIdx.x;
    for (int q = 0; q < p; ++q) yk[q][tid] = rhs[(long long)q * Npad + kt * TS + tid];
    __syncthreads();
    const double* dk = dinv + (long long)kt * TS * TS;
    double z[MAXP];
    for (int q = 0; q < MAXP; ++q) z[q] = 0.0;
    if (!transposed) {
        // z[r] = sum_{c<=r} Dinv[r][c] y[c]
        for (int c = 0; c <= tid; ++c) {
            double dv = dk[tid * TS + c];
            for (int q = 0; q < p; ++q) z[q] = fma(dv, yk[q][c], z[q]);
        }
    } else {
        // z[c] = sum_{r>=c} Dinv[r][c] y[r]   (coalesced over tid = c)
        for (int r = tid; r < TS; ++r) {
            double dv = dk[r * TS + tid];
            for (int q = 0; q < p; ++q) z[q] = fma(dv, yk[q][r], z[q]);
        }
    }
    for (int q = 0; q < p; ++q) zk[q][tid] = z[q];
    __syncthreads();
    if (blockIdx.x == 0)
        for (int q = 0; q < p; ++q) sol[(long long)q * Npad + kt * TS + tid] = z[q];
    // update one other row tile
    int other;
    if (!transposed) {
        other = kt + 1 + blockIdx.x;
        if (other >= Npad / TS) return;
    } else {
        other = kt - 1 - blockIdx.x;
        if (other < 0) return;
    }
    const double* rowp = Lbuf + ((long long)other * TS + tid) * ld + (long long)kt * TS;
    double s[MAXP];
    for (int q = 0; q < MAXP; ++q) s[q] = 0.0;
    for (int c = 0; c < TS; c += 2) {
        double2 lv = *reinterpret_cast<const double2*>(rowp + c);
        for (int q = 0; q < p; ++q) {
            s[q] = fma(lv.x, zk[q][c], s[q]);
            s[q] = fma(lv.y, zk[q][c + 1], s[q]);
        }
    }
    for (int q = 0; q < p; ++q) rhs[(long long)q * Npad + other * TS + tid] -= s[q];
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
                double kr = kp.c * exp(-0.5 * s) * w;
                gc += kr;
#pragma unroll
                for (int a = 0; a < D; ++a) gl[a] = fma(kr, d2[a], gl[a]);
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

}  // namespace gptb
