// Query-side kernels: fused k(x*,X) generator + posterior mean/Jacobian reduction, the variance triangular
// multiply (|L^-1 k*|^2 and |L^-1 dk*/dx_a|^2 on the DMMA tile engine) and the transport epilogue.
#pragma once
#include "factor.cuh"

namespace gptb {

struct Affine {
    int on;                 // apply gamma() to the inputs
    double s;
    double R[MAXD][MAXD];
    double Sbar[MAXD];
    double Tbar[MAXD];
};

constexpr int QPW = 1;          // queries per warp
constexpr int QPB = 8;          // queries per CTA (8 warps)
constexpr int TRMM_GR = 12;     // L2 blocking of the variance triangular multiply: row tiles per group
constexpr int TRMM_GI = 12;     //                                                   inverse-factor row tiles per group

// ------------------------------------------------------------------------------------------------------------
// exp(x) for x <= 0 (the RBF profile), FP64, ~1 ulp: x = m*(ln2/16) + r with m = 16 n + j, |r| <= ln2/32, so that
// exp(x) = 2^n * 2^(j/16) * exp(r) needs only a degree-7 polynomial (remainder 1.2e-18) evaluated in Estrin form
// (dependency depth 3) plus one table multiply.  12 FP64 issue slots instead of libdevice's ~21
// (profiles/r01_fp64_peaks.json), and no special-case branches: arguments below -700 return 0.
// The 16-entry table lives in shared memory (lanes with different j hit different banks).
// ------------------------------------------------------------------------------------------------------------
__constant__ double c_exp2_16[16] = {
    0x1.0000000000000p+0, 0x1.0b5586cf9890fp+0, 0x1.172b83c7d517bp+0, 0x1.2387a6e756238p+0, 0x1.306fe0a31b715p+0, 0x1.3dea64c123422p+0,
    0x1.4bfdad5362a27p+0, 0x1.5ab07dd485429p+0, 0x1.6a09e667f3bcdp+0, 0x1.7a11473eb0187p+0, 0x1.8ace5422aa0dbp+0, 0x1.9c49182a3f090p+0,
    0x1.ae89f995ad3adp+0, 0x1.c199bdd85529cp+0, 0x1.d5818dcfba487p+0, 0x1.ea4afa2a490dap+0};

__device__ __forceinline__ double exp_neg(double x, const double* tab) {
    const double MAGIC = 6755399441055744.0;               // 1.5 * 2^52: the integer lands in the low mantissa word
    const double t = fma(x, 0x1.71547652b82fep+4, MAGIC);   // x * 16/ln2
    const int m = __double2loint(t);
    const double mf = t - MAGIC;
    double r = fma(mf, -0x1.62e42fee00000p-5, x);           // ln2/16 high part (32 significant bits: m*hi is exact)
    r = fma(mf, -0x1.a39ef35793c76p-37, r);
    const double r2 = r * r;
    const double p01 = 1.0 + r;
    const double p23 = fma(r, 1.0 / 6.0, 0.5);
    const double p45 = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    const double p67 = fma(r, 1.0 / 5040.0, 1.0 / 720.0);
    const double r4 = r2 * r2;
    const double lo = fma(r2, p23, p01);
    const double hi = fma(r2, p67, p45);
    double pv = fma(r4, hi, lo) * tab[m & 15];
    // scale by 2^n through the exponent field (n <= 0 here; result stays normal for x >= -700)
    const int n = m >> 4;
    pv = __hiloint2double(__double2hiint(pv) + (n << 20), __double2loint(pv));
    return (x < -700.0) ? 0.0 : pv;
}

// ------------------------------------------------------------------------------------------------------------
// Generator: for a batch of queries, regenerate k(x*, X) on the fly (never read from HBM), reduce it against alpha
// for the mean and the analytic Jacobian, and -- only when a variance is requested -- emit the rows of the
// right-hand-side matrix for the triangular multiply:
//    rhs row (0*Bpad + q)       = k*            (GPTB_STD)
//    rhs row ((1+a)*Bpad + q)   = dk*/dx_a      (GPTB_JACVAR)       = k* (X_a - x_a)/ell_a^2   gaussian_process.py:82-87
//    rhs row ((1+d+a)*Bpad + q) = k* + dk*/dx_a (GPTB_DVAR, polarisation for the cross term) gaussian_process.py:104-126
// Lanes run over training points (coalesced X/alpha loads and rhs stores), each warp keeps QPW queries in
// registers; per-query accumulators are warp-reduced at the end.  grid = (Bpad/QPB, nsplit): the training range is
// split when the batch alone cannot fill 148 SMs, partial sums are reduced in a fixed order by finalize.
// Roofline: FP64 pipe (DFMA issue) -- ~21 DFMA slots for exp + 3D+2 for the distance + P(1+D) accumulate per pair.
// ------------------------------------------------------------------------------------------------------------
template <int D, int P, bool STORE>
__global__ void __launch_bounds__(256) kstar_kernel(const double* __restrict__ xq, const double* __restrict__ Xs,
                                                    const double* __restrict__ alpha, int N, int Npad, int B, int Bpad,
                                                    KParams kp, Affine af, unsigned flags, double* __restrict__ rhs,
                                                    double* __restrict__ xr, double* __restrict__ macc, int nsplit) {
    constexpr int NACC = P + P * D;
    __shared__ double etab[16];
    if (threadIdx.x < 16) etab[threadIdx.x] = c_exp2_16[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q0 = blockIdx.x * QPB + warp * QPW;
    double xs[QPW][D];
    bool valid[QPW];
#pragma unroll
    for (int qq = 0; qq < QPW; ++qq) {
        int q = q0 + qq;
        valid[qq] = q < B;
        double xin[D], xe[D];
#pragma unroll
        for (int a = 0; a < D; ++a) xin[a] = valid[qq] ? xq[(long long)q * D + a] : 0.0;
        if (af.on) {
#pragma unroll
            for (int a = 0; a < D; ++a) {
                double s = 0.0;
#pragma unroll
                for (int b = 0; b < D; ++b) s += af.R[a][b] * (xin[b] - af.Sbar[b]);
                xe[a] = af.s * s + af.Tbar[a];
            }
        } else {
#pragma unroll
            for (int a = 0; a < D; ++a) xe[a] = xin[a];
        }
#pragma unroll
        for (int a = 0; a < D; ++a) xs[qq][a] = xe[a] / kp.ell[a];
        if (blockIdx.y == 0 && lane == 0 && valid[qq])
#pragma unroll
            for (int a = 0; a < D; ++a) xr[(long long)q * D + a] = xe[a];
    }
    double acc[QPW][NACC];
#pragma unroll
    for (int qq = 0; qq < QPW; ++qq)
#pragma unroll
        for (int v = 0; v < NACC; ++v) acc[qq][v] = 0.0;

    const bool st_k = STORE && (flags & 1u), st_g = STORE && (flags & 2u), st_kg = STORE && (flags & 4u);
    const int per = ((Npad / 32 + nsplit - 1) / nsplit) * 32;
    const int nbeg = blockIdx.y * per;
    const int nend = min(Npad, nbeg + per);
    for (int n = nbeg + lane; n < nend; n += 32) {
        double xn[D], al[P];
#pragma unroll
        for (int a = 0; a < D; ++a) xn[a] = Xs[(long long)a * Npad + n];
#pragma unroll
        for (int o = 0; o < P; ++o) al[o] = alpha[(long long)o * Npad + n];
        const bool inb = n < N;
#pragma unroll
        for (int qq = 0; qq < QPW; ++qq) {
            double df[D], s = 0.0;
#pragma unroll
            for (int a = 0; a < D; ++a) {
                df[a] = xs[qq][a] - xn[a];
                s += df[a] * df[a];
            }
            double k = (inb && valid[qq]) ? kp.c * kernel_profile(s, kp.kind, [&](double z) { return exp_neg(z, etab); }) : 0.0;
            const long long row = (long long)(q0 + qq) * Npad + n;
            if (STORE && st_k) rhs[row] = k;
#pragma unroll
            for (int o = 0; o < P; ++o) acc[qq][o] = fma(k, al[o], acc[qq][o]);
#pragma unroll
            for (int a = 0; a < D; ++a) {
                double u = -k * df[a];                 // k * (X_a - x_a)/ell_a
                if (STORE && (st_g | st_kg)) {
                    double gval = u * kp.inv_ell[a];
                    if (st_g) rhs[(long long)(1 + a) * Bpad * Npad + row] = gval;
                    if (st_kg) rhs[(long long)(1 + D + a) * Bpad * Npad + row] = k + gval;
                }
#pragma unroll
                for (int o = 0; o < P; ++o) acc[qq][P + o * D + a] = fma(u, al[o], acc[qq][P + o * D + a]);
            }
        }
    }
#pragma unroll
    for (int qq = 0; qq < QPW; ++qq) {
#pragma unroll
        for (int v = 0; v < NACC; ++v) {
            double x = acc[qq][v];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
            acc[qq][v] = x;
        }
        if (lane == 0) {
            double* dst = macc + ((long long)blockIdx.y * Bpad + q0 + qq) * NACC;
#pragma unroll
            for (int o = 0; o < P; ++o) dst[o] = acc[qq][o];
#pragma unroll
            for (int o = 0; o < P; ++o)
#pragma unroll
                for (int a = 0; a < D; ++a) dst[P + o * D + a] = acc[qq][P + o * D + a] * kp.inv_ell[a];
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Variance triangular multiply:  W = RHS * Linv^T restricted to k <= i (Linv lower), reduced on the fly to row sums of
// squares:  part[ti][row] = sum_{i in tile ti} ( sum_{k<=i} RHS[row][k] * Linv[i][k] )^2.
// This is the batched TRSM of the reference (solve_triangular(L, K*^T), sklearn:_gpr.py:460) re-expressed as a TRMM with the
// explicit inverse factor so every flop is a DMMA tile; W is never written to memory.
// Algorithmic work per 128x128 output tile with ti: 2*128*128*128*(ti+1) flops.  Heaviest tiles are scheduled first.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEMM_THREADS, 1) trmm_sumsq_kernel(const __grid_constant__ CUtensorMap mapR,
                                                                    const __grid_constant__ CUtensorMap mapM, int T, int rowtiles,
                                                                    long long rows_total, double* __restrict__ part) {
    extern __shared__ __align__(128) double smem[];
    __shared__ PipeBarriers pipe;
    pipe_init(&pipe);
    // Tile order = L2 blocking: CTAs that run concurrently (consecutive blockIdx) form groups of TRMM_GR row tiles x
    // TRMM_GI inverse-factor row tiles, so every operand slab fetched from HBM is reused ~12x out of L2 instead of
    // being re-read per tile (35.7 GB -> a few GB per launch at N=4096, profiles/r01_trmm_*).  ti-blocks run from the
    // heaviest (longest k-range) to the lightest.
    int ti, rt;
    {
        const long long idx = blockIdx.x;
        const long long per_tib = (long long)rowtiles * TRMM_GI;
        const int tib = (int)(idx / per_tib);
        const long long rem = idx - (long long)tib * per_tib;
        const int ti_cnt = min(TRMM_GI, T - tib * TRMM_GI);
        const int rb = (int)(rem / ((long long)TRMM_GR * ti_cnt));
        const int rem2 = (int)(rem - (long long)rb * TRMM_GR * ti_cnt);
        rt = rb * TRMM_GR + rem2 / ti_cnt;
        ti = T - 1 - (tib * TRMM_GI + rem2 % ti_cnt);
    }
    double acc[8][4][2];
    acc_clear(acc);
    Operand A{&mapR, rt * TS, 0, MASK_NONE, -1};
    Operand B{&mapM, ti * TS, 0, MASK_LOWER, ti};
    gemm_nt_tile(A, B, 0, ti + 1, acc, smem, &pipe);
    if (!is_consumer()) return;
    // row sums of squares: thread owns rows wm*64+mi*8+g; reduce over its 8 columns, the 4 lanes t, then the 4 wn warps
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3, wm = warp >> 2, wn = warp & 3;
    double* red = smem;   // [4][128]
#pragma unroll
    for (int mi = 0; mi < 8; ++mi) {
        double s = 0.0;
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            s = fma(acc[mi][ni][0], acc[mi][ni][0], s);
            s = fma(acc[mi][ni][1], acc[mi][ni][1], s);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (t == 0) red[wn * TS + wm * 64 + mi * 8 + g] = s;
    }
    consumer_sync();
    if (threadIdx.x < TS) {
        int r = threadIdx.x;
        double s = (red[r] + red[TS + r]) + (red[2 * TS + r] + red[3 * TS + r]);
        part[(long long)ti * rows_total + (long long)rt * TS + r] = s;
    }
}

// ------------------------------------------------------------------------------------------------------------
// Joint posterior covariance (GaussianProcess.predict(return_cov=True) / samples(); gaussian_process.py:50-60,
// sklearn:_gpr.py:470-475): W = RHS * Linv^T is materialised once (same tile engine, store epilogue), then
//   cov[a][b] = k(x_a, x_b) + s2*[a==b] - sum_k W[a][k] W[b][k].
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEMM_THREADS, 1) trmm_store_kernel(const __grid_constant__ CUtensorMap mapR,
                                                                    const __grid_constant__ CUtensorMap mapM, int T, int rowtiles,
                                                                    double* __restrict__ W, long long ld) {
    extern __shared__ __align__(128) double smem[];
    __shared__ PipeBarriers pipe;
    pipe_init(&pipe);
    const int ti = T - 1 - (int)(blockIdx.x / rowtiles);
    const int rt = (int)(blockIdx.x % rowtiles);
    double acc[8][4][2];
    acc_clear(acc);
    Operand A{&mapR, rt * TS, 0, MASK_NONE, -1};
    Operand B{&mapM, ti * TS, 0, MASK_LOWER, ti};
    gemm_nt_tile(A, B, 0, ti + 1, acc, smem, &pipe);
    double* out = W + (long long)rt * TS * ld + (long long)ti * TS;
    acc_foreach(acc, [&](int r, int c, double v0, double v1) {
        *reinterpret_cast<double2*>(out + (long long)r * ld + c) = make_double2(v0, v1);
    });
}

template <int D>
__global__ void __launch_bounds__(GEMM_THREADS, 1) cov_kernel(const __grid_constant__ CUtensorMap mapW, int T, int mtiles,
                                                             const double* __restrict__ xr, KParams kp, int M,
                                                             double* __restrict__ cov) {
    extern __shared__ __align__(128) double smem[];
    __shared__ PipeBarriers pipe;
    __shared__ double xa[D][TS], xb[D][TS];
    const int ta = blockIdx.x / mtiles, tb = blockIdx.x % mtiles;
    for (int e = threadIdx.x; e < D * TS; e += GEMM_THREADS) {
        const int a = e / TS, r = e % TS;
        const int qa = ta * TS + r, qb = tb * TS + r;
        xa[a][r] = (qa < M) ? xr[(long long)qa * D + a] / kp.ell[a] : 0.0;
        xb[a][r] = (qb < M) ? xr[(long long)qb * D + a] / kp.ell[a] : 0.0;
    }
    pipe_init(&pipe);
    double acc[8][4][2];
    acc_clear(acc);
    Operand A{&mapW, ta * TS, 0, MASK_NONE, -1};
    Operand B{&mapW, tb * TS, 0, MASK_NONE, -1};
    gemm_nt_tile(A, B, 0, T, acc, smem, &pipe);
    acc_foreach(acc, [&](int r, int c, double v0, double v1) {
        const int qa = ta * TS + r;
        if (qa >= M) return;
        double v[2] = {v0, v1};
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int qb = tb * TS + c + j;
            if (qb < M) {
                double s = 0.0;
#pragma unroll
                for (int a = 0; a < D; ++a) {
                    const double df = xa[a][r] - xb[a][c + j];
                    s += df * df;
                }
                double kss = kp.c * kernel_profile(s, kp.kind, [](double z) { return exp(z); });
                if (qa == qb) kss += kp.s2;
                cov[(long long)qa * M + qb] = kss - v[j];
            }
        }
    });
}

// ------------------------------------------------------------------------------------------------------------
// Epilogue: fixed-order reduction of the partials, then every reference-layout output of the transport flow.
// ------------------------------------------------------------------------------------------------------------
struct QueryOut {
    double* mean;    // (M,p)
    double* std;     // (M,p)
    double* jac;     // (M,p,d)
    double* jacvar;  // (M,p,d)
    double* xhat;    // (M,d)
    double* vhat;    // (M,d)
    double* vvar;    // (M,p)
    double* jphi;    // (M,d,d)
    double* dvar;    // (d,M)  -- leading dimension Mtot
};

template <int D, int P>
__global__ void __launch_bounds__(128) finalize_kernel(const double* __restrict__ macc, int nsplit, const double* __restrict__ part,
                                                       int T, int B, int Bpad, long long rows_total, const double* __restrict__ xr,
                                                       const double* __restrict__ vel, KParams kp, Affine af, unsigned qflags,
                                                       QueryOut out, long long q_off, long long Mtot) {
    constexpr int NACC = P + P * D;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= B) return;
    const long long gq = q_off + q;
    double m[NACC];
#pragma unroll
    for (int v = 0; v < NACC; ++v) m[v] = 0.0;
    for (int sidx = 0; sidx < nsplit; ++sidx) {
        const double* src = macc + ((long long)sidx * Bpad + q) * NACC;
#pragma unroll
        for (int v = 0; v < NACC; ++v) m[v] += src[v];
    }
    if (qflags & 0x001u)
#pragma unroll
        for (int o = 0; o < P; ++o) out.mean[gq * P + o] = m[o];
    if (qflags & 0x004u)
#pragma unroll
        for (int v = 0; v < P * D; ++v) out.jac[gq * P * D + v] = m[P + v];
    if (qflags & 0x002u) {
        double ss = 0.0;
        for (int ti = 0; ti < T; ++ti) ss += part[(long long)ti * rows_total + q];
        double var = (kp.c + kp.s2) - ss;
        var = var < 0.0 ? 0.0 : var;                                    // sklearn:_gpr.py:485-491
        double sd = sqrt(var) - sqrt(kp.s2);                            // gaussian_process.py:49
#pragma unroll
        for (int o = 0; o < P; ++o) out.std[gq * P + o] = sd;
    }
    double jv[D];
    if (qflags & (0x008u | 0x100u)) {
#pragma unroll
        for (int a = 0; a < D; ++a) {
            double ss = 0.0;
            for (int ti = 0; ti < T; ++ti) ss += part[(long long)ti * rows_total + (long long)(1 + a) * Bpad + q];
            jv[a] = ss;
        }
    }
    if (qflags & 0x100u) {
        double s0 = 0.0;
        for (int ti = 0; ti < T; ++ti) s0 += part[(long long)ti * rows_total + q];
#pragma unroll
        for (int a = 0; a < D; ++a) {
            double sk = 0.0;
            for (int ti = 0; ti < T; ++ti) sk += part[(long long)ti * rows_total + (long long)(1 + D + a) * Bpad + q];
            out.dvar[(long long)a * Mtot + gq] = -(sk - jv[a] - s0);     // -2 g^T K^-1 k*
        }
    }
    if (qflags & 0x008u) {
#pragma unroll
        for (int a = 0; a < D; ++a) {
            jv[a] = kp.c / (kp.ell[a] * kp.ell[a]) - jv[a];             // gaussian_process.py:98
#pragma unroll
            for (int o = 0; o < P; ++o) out.jacvar[(gq * P + o) * D + a] = jv[a];
        }
    }
    if (qflags & 0x020u)
#pragma unroll
        for (int a = 0; a < D; ++a) out.xhat[gq * D + a] = xr[(long long)q * D + a] + m[a < P ? a : 0];
    if constexpr (D == P) {
        if (qflags & (0x040u | 0x080u)) {
            double jp[D][D];                                             // Jphi = R + Jpsi R   policy_transportation.py:45
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < D; ++k) s += m[P + i * D + k] * af.R[k][j];
                    jp[i][j] = af.R[i][j] + s;
                }
            if (qflags & 0x080u)
#pragma unroll
                for (int i = 0; i < D; ++i)
#pragma unroll
                    for (int j = 0; j < D; ++j) out.jphi[(gq * D + i) * D + j] = jp[i][j];
            if (qflags & 0x040u) {
                double v[D], rv[D];
#pragma unroll
                for (int a = 0; a < D; ++a) v[a] = vel[(long long)q * D + a];
#pragma unroll
                for (int i = 0; i < D; ++i) {
                    double s = 0.0, s2 = 0.0;
#pragma unroll
                    for (int j = 0; j < D; ++j) {
                        s += jp[i][j] * v[j];
                        s2 += af.R[i][j] * v[j];
                    }
                    out.vhat[gq * D + i] = s;
                    rv[i] = s2;
                }
                if (qflags & 0x008u) {
                    double s = 0.0;
#pragma unroll
                    for (int a = 0; a < D; ++a) s += jv[a] * (rv[a] * rv[a]);   // policy_transportation.py:51-52
#pragma unroll
                    for (int o = 0; o < P; ++o) out.vvar[gq * P + o] = s;
                }
            }
        }
    }
}

}  // namespace gptb
