// Query-side kernels: fused k(x*,X) generator + posterior mean/Jacobian reduction, the variance triangular
// multiply (|L^-1 k*|^2 and |L^-1 dk*/dx_a|^2 on the DMMA tile engine) and the transport epilogue.
#pragma once
#include "factor.cuh"
#include "digits.cuh"

namespace gptb {

struct Affine {
    int on;                 // apply gamma() to the inputs
    double s;
    double R[MAXD][MAXD];
    double Sbar[MAXD];
    double Tbar[MAXD];
};

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
// MODE 0: no right-hand-side rows (mean/Jacobian only); 1: FP64 rows into `rhs`; 2: int8 digit planes into `planes`
// (INT8-sliced variance path, ozaki.cuh) -- the digits are produced here, straight from the freshly generated values,
// with one power-of-two scale per row TYPE (k* <= c, |dk*/dx_a| <= c/ell_a, ...), so the FP64 rows never touch HBM.
struct DigitScales {
    double down[1 + 2 * MAXD];     // 2^-e per row type: 0 = k*, 1+a = dk*/dx_a, 1+D+a = k* + dk*/dx_a
    double scale[1 + 2 * MAXD];    // 2^e
    // spatial mode (gptb_set_spatial): queries are processed in Morton order (qperm[sorted position] = position in the batch) and
    // the generator records, per (128-query row tile, 64-point chunk), which digit planes hold a non-zero digit (one byte per
    // block, bit t = plane t; rows of `flags` are flags_stride 32-bit words long) so the product kernel can skip zero planes.
    const unsigned* qperm;
    unsigned* flags;
    int flags_stride;
};
constexpr int FLAG_WORDS = 128;    // 512 chunks of 64 training points = N <= 32768

// Morton code of the (affine-transformed) query positions relative to the bounding box of the training inputs; keys are sorted
// with a radix sort, vals carries the position in the batch.
struct MortonBox {
    double lo[MAXD], scale[MAXD];
    int bits;
};
template <int D>
__global__ void __launch_bounds__(256) morton_keys_kernel(const double* __restrict__ xq, int B, Affine af, MortonBox mb, unsigned* __restrict__ keys,
                                                          unsigned* __restrict__ vals) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= B) return;
    double xin[D], xe[D];
#pragma unroll
    for (int a = 0; a < D; ++a) xin[a] = xq[(long long)q * D + a];
    if (af.on) {
#pragma unroll
        for (int a = 0; a < D; ++a) {
            double sacc = 0.0;
#pragma unroll
            for (int b = 0; b < D; ++b) sacc += af.R[a][b] * (xin[b] - af.Sbar[b]);
            xe[a] = af.s * sacc + af.Tbar[a];
        }
    } else {
#pragma unroll
        for (int a = 0; a < D; ++a) xe[a] = xin[a];
    }
    const double top = (double)((1u << mb.bits) - 1u);
    unsigned cell[D];
#pragma unroll
    for (int a = 0; a < D; ++a) {
        double t = (xe[a] - mb.lo[a]) * mb.scale[a];
        t = (t > 0.0) ? ((t < top) ? t : top) : 0.0;              // clamps NaN to 0 as well
        cell[a] = (unsigned)t;
    }
    unsigned code = 0;
    for (int b = 0; b < mb.bits; ++b)
#pragma unroll
        for (int a = 0; a < D; ++a) code |= ((cell[a] >> b) & 1u) << (b * D + a);
    keys[q] = code;
    vals[q] = (unsigned)q;
}


template <int S, int BITS>
__device__ __forceinline__ void emit_digits(const double (&val)[4], double down, int8_t* __restrict__ planes, long long plane_stride,
                                            long long off, unsigned* smask_row = nullptr, int chunk = 0) {
    unsigned packed[S];
    if constexpr (BITS == 8) {
        digits8_pack4<S>(val, down, packed);          // here `down` already carries the 256^S factor (DigitScales)
    } else {
        // same cascade as oz::slice_rows_kernel (error-free: scaling by powers of two, rint and the subtraction are exact)
#pragma unroll
        for (int t = 0; t < S; ++t) packed[t] = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double y = val[j] * down;
#pragma unroll
            for (int t = 0; t < S; ++t) {
                const double dgt = digit_step<BITS>(y);
                packed[t] |= ((unsigned)(__double2int_rn(dgt)) & 0xffu) << (8 * j);
            }
        }
    }
#pragma unroll
    for (int t = 0; t < S; ++t) *reinterpret_cast<unsigned*>(planes + (long long)t * plane_stride + off) = packed[t];
    if (smask_row != nullptr) {
        // non-zero planes of this lane's four digits, OR-ed over the 16 lanes that share a 64-point chunk, into the CTA's mask
        unsigned m = 0;
#pragma unroll
        for (int t = 0; t < S; ++t) m |= (packed[t] != 0u ? 1u : 0u) << t;
        m |= __shfl_xor_sync(0xffffffffu, m, 1);
        m |= __shfl_xor_sync(0xffffffffu, m, 2);
        m |= __shfl_xor_sync(0xffffffffu, m, 4);
        m |= __shfl_xor_sync(0xffffffffu, m, 8);
        if ((threadIdx.x & 15) == 0 && m != 0u) atomicOr(smask_row + (chunk >> 2), m << ((chunk & 3) * 8));
    }
}

template <int D, int P, int MODE, int S, int BITS = 7>
__global__ void __launch_bounds__(256) kstar_kernel(const double* __restrict__ xq, const double* __restrict__ Xs,
                                                    const double* __restrict__ alpha, int N, int Npad, int B, int Bpad,
                                                    KParams kp, Affine af, unsigned flags, double* __restrict__ rhs,
                                                    double* __restrict__ xr, double* __restrict__ macc, int nsplit,
                                                    int8_t* __restrict__ planes, long long plane_stride, double* __restrict__ scaleA,
                                                    DigitScales ds) {
    constexpr int NACC = P + P * D;
    __shared__ double etab[16];
    __shared__ unsigned smask[(MODE == 2 && BITS == 8) ? (1 + 2 * D) : 1][(MODE == 2 && BITS == 8) ? FLAG_WORDS : 1];
    const bool use_flags = (MODE == 2 && BITS == 8) && ds.flags != nullptr;
    if (threadIdx.x < 16) etab[threadIdx.x] = c_exp2_16[threadIdx.x];
    if (use_flags)
        for (int e = threadIdx.x; e < (1 + 2 * D) * FLAG_WORDS; e += 256) (&smask[0][0])[e] = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = blockIdx.x * QPB + warp;             // one query per warp; lanes run over training points (4 each)
    const bool valid = q < B;
    const long long qsrc = (MODE == 2 && ds.qperm != nullptr && valid) ? (long long)ds.qperm[q] : (long long)q;   // spatial mode: sorted order
    double xs[D];
    {
        double xin[D], xe[D];
#pragma unroll
        for (int a = 0; a < D; ++a) xin[a] = valid ? xq[qsrc * D + a] : 0.0;
        if (af.on) {
#pragma unroll
            for (int a = 0; a < D; ++a) {
                double sacc = 0.0;
#pragma unroll
                for (int b = 0; b < D; ++b) sacc += af.R[a][b] * (xin[b] - af.Sbar[b]);
                xe[a] = af.s * sacc + af.Tbar[a];
            }
        } else {
#pragma unroll
            for (int a = 0; a < D; ++a) xe[a] = xin[a];
        }
#pragma unroll
        for (int a = 0; a < D; ++a) xs[a] = xe[a] / kp.ell[a];
        if (blockIdx.y == 0 && lane == 0 && valid)
#pragma unroll
            for (int a = 0; a < D; ++a) xr[(long long)q * D + a] = xe[a];
    }
    double acc[NACC];
#pragma unroll
    for (int v = 0; v < NACC; ++v) acc[v] = 0.0;

    const bool st_k = MODE != 0 && (flags & 1u), st_g = MODE != 0 && (flags & 2u), st_kg = MODE != 0 && (flags & 4u);
    if (MODE == 2 && blockIdx.y == 0 && lane == 0) {
        if (st_k) scaleA[q] = ds.scale[0];
#pragma unroll
        for (int a = 0; a < D; ++a) {
            if (st_g) scaleA[(long long)(1 + a) * Bpad + q] = ds.scale[1 + a];
            if (st_kg) scaleA[(long long)(1 + D + a) * Bpad + q] = ds.scale[1 + D + a];
        }
    }
    const int per = ((Npad / 128 + nsplit - 1) / nsplit) * 128;
    const int nbeg = blockIdx.y * per;
    const int nend = min(Npad, nbeg + per);
    for (int n0 = nbeg + 4 * lane; n0 < nend; n0 += 128) {
        double xn[D][4], al[P][4];
#pragma unroll
        for (int a = 0; a < D; ++a) {
            const double2 u0 = *reinterpret_cast<const double2*>(Xs + (long long)a * Npad + n0);
            const double2 u1 = *reinterpret_cast<const double2*>(Xs + (long long)a * Npad + n0 + 2);
            xn[a][0] = u0.x; xn[a][1] = u0.y; xn[a][2] = u1.x; xn[a][3] = u1.y;
        }
#pragma unroll
        for (int o = 0; o < P; ++o) {
            const double2 u0 = *reinterpret_cast<const double2*>(alpha + (long long)o * Npad + n0);
            const double2 u1 = *reinterpret_cast<const double2*>(alpha + (long long)o * Npad + n0 + 2);
            al[o][0] = u0.x; al[o][1] = u0.y; al[o][2] = u1.x; al[o][3] = u1.y;
        }
        double kv[4], gv[D][4];
        double dfj[D][4], sj[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            sj[j] = 0.0;
#pragma unroll
            for (int a = 0; a < D; ++a) {
                dfj[a][j] = xs[a] - xn[a][j];
                sj[j] += dfj[a][j] * dfj[a][j];
            }
        }
        // the kernel family is a run-time parameter: branch ONCE per four points (not per point) so that the four exp chains of
        // the RBF path sit in one basic block and interleave on the FP64 pipe
        if (kp.kind == KIND_RBF) {
#pragma unroll
            for (int j = 0; j < 4; ++j) kv[j] = exp_neg(-0.5 * sj[j], etab);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) kv[j] = kernel_profile(sj[j], kp.kind, [&](double z) { return exp_neg(z, etab); });
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool inb = (n0 + j) < N;
            const double k = (inb && valid) ? kp.c * kv[j] : 0.0;
            kv[j] = k;
#pragma unroll
            for (int o = 0; o < P; ++o) acc[o] = fma(k, al[o][j], acc[o]);
#pragma unroll
            for (int a = 0; a < D; ++a) {
                const double u = -k * dfj[a][j];             // k * (X_a - x_a)/ell_a
                gv[a][j] = u * kp.inv_ell[a];                // dk*/dx_a
#pragma unroll
                for (int o = 0; o < P; ++o) acc[P + o * D + a] = fma(u, al[o][j], acc[P + o * D + a]);
            }
        }
        if (MODE == 1) {
            const long long row = (long long)q * Npad + n0;
            if (st_k) {
                *reinterpret_cast<double2*>(rhs + row) = make_double2(kv[0], kv[1]);
                *reinterpret_cast<double2*>(rhs + row + 2) = make_double2(kv[2], kv[3]);
            }
#pragma unroll
            for (int a = 0; a < D; ++a) {
                if (st_g) {
                    double* dst = rhs + (long long)(1 + a) * Bpad * Npad + row;
                    *reinterpret_cast<double2*>(dst) = make_double2(gv[a][0], gv[a][1]);
                    *reinterpret_cast<double2*>(dst + 2) = make_double2(gv[a][2], gv[a][3]);
                }
                if (st_kg) {
                    double* dst = rhs + (long long)(1 + D + a) * Bpad * Npad + row;
                    *reinterpret_cast<double2*>(dst) = make_double2(kv[0] + gv[a][0], kv[1] + gv[a][1]);
                    *reinterpret_cast<double2*>(dst + 2) = make_double2(kv[2] + gv[a][2], kv[3] + gv[a][3]);
                }
            }
        } else if (MODE == 2) {
            const long long off = (long long)q * Npad + n0;
            const int chunk = n0 >> 6;
            constexpr bool FL = (BITS == 8);
            if (st_k) emit_digits<S, BITS>(kv, ds.down[0], planes, plane_stride, off, (FL && use_flags) ? smask[0] : nullptr, chunk);
#pragma unroll
            for (int a = 0; a < D; ++a) {
                if (st_g)
                    emit_digits<S, BITS>(gv[a], ds.down[1 + a], planes, plane_stride, (long long)(1 + a) * Bpad * Npad + off,
                                         (FL && use_flags) ? smask[FL ? 1 + a : 0] : nullptr, chunk);
                if (st_kg) {
                    const double kg[4] = {kv[0] + gv[a][0], kv[1] + gv[a][1], kv[2] + gv[a][2], kv[3] + gv[a][3]};
                    emit_digits<S, BITS>(kg, ds.down[1 + D + a], planes, plane_stride, (long long)(1 + D + a) * Bpad * Npad + off,
                                         (FL && use_flags) ? smask[FL ? 1 + D + a : 0] : nullptr, chunk);
                }
            }
        }
    }
#pragma unroll
    for (int v = 0; v < NACC; ++v) {
        double x = acc[v];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
        acc[v] = x;
    }
    if (use_flags) {
        // flush the CTA's block masks: its 8 queries sit in one 128-query row tile per row type
        __syncthreads();
        const int words = (Npad / 64 + 3) / 4;
        for (int e = threadIdx.x; e < (1 + 2 * D) * words; e += 256) {
            const int r = e / words, w = e % words;
            const unsigned m = smask[r][w];
            if (m != 0u) {
                const long long rowtile = ((long long)r * Bpad + (long long)blockIdx.x * QPB) / 128;
                atomicOr(ds.flags + rowtile * ds.flags_stride + w, m);
            }
        }
    }
    if (lane == 0) {
        double* dst = macc + ((long long)blockIdx.y * Bpad + q) * NACC;
#pragma unroll
        for (int o = 0; o < P; ++o) dst[o] = acc[o];
#pragma unroll
        for (int o = 0; o < P; ++o)
#pragma unroll
            for (int a = 0; a < D; ++a) dst[P + o * D + a] = acc[P + o * D + a] * kp.inv_ell[a];
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
// The same product for SMALL batches (rowtiles * T below the SM count: a transport of ~100 points, a rollout step, a probe): with one
// CTA per output tile the launch lasts as long as its longest k-loop, T tiles of 17 us each on ONE SM -- 0.54 ms at N = 4096 for a
// single query -- while the rest of the machine idles.  Split-k: job (rt, ti, kc) multiplies k-tiles [kc*KC, min((kc+1)*KC, ti+1))
// and stores its 128 x 128 partial of W; a second kernel adds the partials of one output tile in k order, squares and sums the rows.
// KC is the smallest chunk for which all jobs fit the machine in one wave (host: trmm_splitk_plan).  Fixed summation order:
// deterministic, but not bit-identical to the unsplit kernel (the k-sum is associated differently).
//   job = rt * J + off(ti) + kc,  off(ti) = sum_{t < ti} ceil((t+1)/KC),  J = off(T).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEMM_THREADS, 1) trmm_splitk_kernel(const __grid_constant__ CUtensorMap mapR,
                                                                     const __grid_constant__ CUtensorMap mapM, int T, int KC, int J,
                                                                     double* __restrict__ Wpart) {
    extern __shared__ __align__(128) double smem[];
    __shared__ PipeBarriers pipe;
    pipe_init(&pipe);
    const int job = blockIdx.x, rt = job / J;
    int rem = job - rt * J, ti = 0;
    for (;; ++ti) {
        const int n = (ti + KC) / KC;
        if (rem < n) break;
        rem -= n;
    }
    const int kb = rem * KC, ke = min(kb + KC, ti + 1);
    double acc[8][4][2];
    acc_clear(acc);
    Operand A{&mapR, rt * TS, 0, MASK_NONE, -1};
    Operand B{&mapM, ti * TS, 0, MASK_LOWER, ti};
    gemm_nt_tile(A, B, kb, ke, acc, smem, &pipe);
    double* out = Wpart + (long long)job * TS * TS;
    acc_foreach(acc, [&](int r, int c, double v0, double v1) { *reinterpret_cast<double2*>(out + r * TS + c) = make_double2(v0, v1); });
}

__global__ void __launch_bounds__(256) trmm_splitk_reduce_kernel(const double* __restrict__ Wpart, int T, int KC, int J, long long rows_total,
                                                                 double* __restrict__ part) {
    const int ti = blockIdx.x % T, rt = blockIdx.x / T;
    int off = 0;
    for (int t = 0; t < ti; ++t) off += (t + KC) / KC;
    const int n = (ti + KC) / KC;
    const double* base = Wpart + ((long long)rt * J + off) * TS * TS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll 4
    for (int rr = 0; rr < 16; ++rr) {
        const int r = warp * 16 + rr;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        for (int kc = 0; kc < n; ++kc) {
            const double* p = base + (long long)kc * TS * TS + r * TS + lane * 4;
            const double2 a = *reinterpret_cast<const double2*>(p);
            const double2 b = *reinterpret_cast<const double2*>(p + 2);
            s0 += a.x; s1 += a.y; s2 += b.x; s3 += b.y;
        }
        double q = s0 * s0;
        q = fma(s1, s1, q);
        q = fma(s2, s2, q);
        q = fma(s3, s3, q);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        if (lane == 0) part[(long long)ti * rows_total + (long long)rt * TS + r] = q;
    }
}

// ------------------------------------------------------------------------------------------------------------
// ... and for at most EIGHT right-hand-side rows (one query with its polarisation rows: a rollout step, a control-loop query): the
// tile engine would pad every row to a 128-row tile -- seven tiles' worth of products for one point with derivative_of_variance --
// so this is the matrix-vector form instead: L^-1 is streamed ONCE (it stays in L2 between the steps of a rollout), the right-hand
// sides are the eight columns of the DMMA B fragment.  A CTA works on one 128-row tile of L^-1 with 16 warps of 8 rows; a lane reads 32 contiguous bytes of its row per step and feeds them to four DMMAs (the k-slot <-> lane map is the same
// for both operands, so any consistent assignment is a valid product).  Column n is right-hand side n / B of query n % B.
// Jobs are (tile ti, chunk kc of KC k-tiles), enumerated as in the split-k kernel, so the triangular matrix spreads evenly over the SMs
// (one CTA per tile left the last tile's 128 x N block to a single SM's load bandwidth); each job stores its 128 x 8 partial of W and
// trmv_reduce_kernel adds a tile's partials in k order, squares and sums the rows.
__global__ void __launch_bounds__(512) trmv_partial_kernel(const double* __restrict__ rhs, const double* __restrict__ Minv, long long ld, int B,
                                                           int nrhs, int Bpad, int KC, double* __restrict__ Wpart) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    int rem = blockIdx.x, ti = 0;
    for (;; ++ti) {
        const int n = (ti + KC) / KC;
        if (rem < n) break;
        rem -= n;
    }
    const int i0 = ti * TS + warp * 8;
    const bool valid = g < B * nrhs;
    const long long rrow = valid ? (long long)(g / B) * Bpad + (g % B) : 0;
    const double* rp = rhs + rrow * ld + 4 * t;
    const double* mp = Minv + (long long)(i0 + g) * ld + 4 * t;
    double c0 = 0.0, c1 = 0.0;
    const int kbeg = rem * KC * TS;
    const int kend = min(kbeg + KC * TS, i0 + 8);       // columns k < i0 + 8 can be non-zero in rows i0 .. i0 + 7
#pragma unroll 2
    for (int k0 = kbeg; k0 < kend; k0 += 16) {
        double2 a01 = *reinterpret_cast<const double2*>(mp + k0);
        double2 a23 = *reinterpret_cast<const double2*>(mp + k0 + 2);
        double2 b01 = *reinterpret_cast<const double2*>(rp + k0);
        double2 b23 = *reinterpret_cast<const double2*>(rp + k0 + 2);
        if (!valid) { b01 = make_double2(0.0, 0.0); b23 = b01; }
        if (k0 + 16 > i0) {                             // the diagonal block: L^-1 is lower triangular, the buffer holds its mirror above
            const int k = k0 + 4 * t, i = i0 + g;
            if (k > i) a01.x = 0.0;
            if (k + 1 > i) a01.y = 0.0;
            if (k + 2 > i) a23.x = 0.0;
            if (k + 3 > i) a23.y = 0.0;
        }
        dmma884(c0, c1, a01.x, b01.x);
        dmma884(c0, c1, a01.y, b01.y);
        dmma884(c0, c1, a23.x, b23.x);
        dmma884(c0, c1, a23.y, b23.y);
    }
    // c0, c1 = partial W[i0 + g][n = 2t, 2t + 1]
    *reinterpret_cast<double2*>(Wpart + (long long)blockIdx.x * (TS * 8) + (warp * 8 + g) * 8 + 2 * t) = make_double2(c0, c1);
}

__global__ void __launch_bounds__(128) trmv_reduce_kernel(const double* __restrict__ Wpart, int B, int nrhs, int Bpad, int KC, long long rows_total,
                                                          double* __restrict__ part) {
    __shared__ double red[4][8];
    const int ti = blockIdx.x, i = threadIdx.x, lane = i & 31, warp = i >> 5;
    int off = 0;
    for (int t = 0; t < ti; ++t) off += (t + KC) / KC;
    const int n = (ti + KC) / KC;
    double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int kc = 0; kc < n; ++kc) {
        const double* p = Wpart + (long long)(off + kc) * (TS * 8) + i * 8;
#pragma unroll
        for (int c = 0; c < 8; c += 2) {
            const double2 v = *reinterpret_cast<const double2*>(p + c);
            s[c] += v.x;
            s[c + 1] += v.y;
        }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        double q = s[c] * s[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        if (lane == 0) red[warp][c] = q;
    }
    __syncthreads();
    if (i < 8 && i < B * nrhs) part[(long long)ti * rows_total + (long long)(i / B) * Bpad + (i % B)] = (red[0][i] + red[1][i]) + (red[2][i] + red[3][i]);
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
                                                       QueryOut out, long long q_off, long long Mtot, const unsigned* __restrict__ qperm) {
    constexpr int NACC = P + P * D;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= B) return;
    const long long qsrc = qperm ? (long long)qperm[q] : (long long)q;    // spatial mode: row q of the batch is query qsrc
    const long long gq = q_off + qsrc;
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
                for (int a = 0; a < D; ++a) v[a] = vel[qsrc * D + a];
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

// ------------------------------------------------------------------------------------------------------------
// Orientation epilogue (policy_transportation.py:61-77): q_hat = quat(Jphi) (x) q, with quat() the Bar-Itzhack
// quaternion of a possibly non-orthogonal 3x3 matrix (numpy-quaternion's from_rotation_matrix(nonorthogonal=True)):
// dominant eigenvector of the symmetric 4x4 matrix K3, found here with cyclic Jacobi rotations (one thread per point,
// everything in registers).  Quaternions are (w, x, y, z); the eigenvector sign is normalised to w >= 0.
// ------------------------------------------------------------------------------------------------------------
// Bar-Itzhack quaternion (w, x, y, z; sign normalised to w >= 0) of a possibly non-orthogonal 3x3 matrix: dominant eigenvector of the
// symmetric 4x4 matrix K3 by cyclic Jacobi rotations, everything in registers.
__device__ __forceinline__ void quat_of_matrix(const double (&R)[3][3], double (&qa)[4]) {
    double A[4][4];
    A[0][0] = (R[0][0] - R[1][1] - R[2][2]) / 3.0;
    A[0][1] = (R[1][0] + R[0][1]) / 3.0;
    A[0][2] = (R[2][0] + R[0][2]) / 3.0;
    A[0][3] = (R[1][2] - R[2][1]) / 3.0;
    A[1][1] = (R[1][1] - R[0][0] - R[2][2]) / 3.0;
    A[1][2] = (R[2][1] + R[1][2]) / 3.0;
    A[1][3] = (R[2][0] - R[0][2]) / 3.0;
    A[2][2] = (R[2][2] - R[0][0] - R[1][1]) / 3.0;
    A[2][3] = (R[0][1] - R[1][0]) / 3.0;
    A[3][3] = (R[0][0] + R[1][1] + R[2][2]) / 3.0;
#pragma unroll
    for (int r = 1; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < r; ++c) A[r][c] = A[c][r];
    double V[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) V[r][c] = (r == c) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 24; ++sweep) {
        double off = 0.0, diag = 0.0;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            diag += A[r][r] * A[r][r];
#pragma unroll
            for (int c = r + 1; c < 4; ++c) off += A[r][c] * A[r][c];
        }
        if (off <= 1e-34 * (diag + off)) break;
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int q = p + 1; q < 4; ++q) {
                const double apq = A[p][q];
                if (apq != 0.0) {
                    const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
                    const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                    const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {           // A <- A J
                        const double akp = A[k][p], akq = A[k][q];
                        A[k][p] = c * akp - sn * akq;
                        A[k][q] = sn * akp + c * akq;
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {           // A <- J^T A
                        const double apk = A[p][k], aqk = A[q][k];
                        A[p][k] = c * apk - sn * aqk;
                        A[q][k] = sn * apk + c * aqk;
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {           // V <- V J
                        const double vkp = V[k][p], vkq = V[k][q];
                        V[k][p] = c * vkp - sn * vkq;
                        V[k][q] = sn * vkp + c * vkq;
                    }
                }
            }
    }
    int best = 0;
#pragma unroll
    for (int r = 1; r < 4; ++r)
        if (A[r][r] > A[best][best]) best = r;
    double e[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        double v = V[r][0];
        if (best == 1) v = V[r][1];
        if (best == 2) v = V[r][2];
        if (best == 3) v = V[r][3];
        e[r] = v;
    }
    double nrm = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2] + e[3] * e[3]);
    double sgn = (e[3] < 0.0) ? -1.0 / nrm : 1.0 / nrm;
    qa[0] = e[3] * sgn; qa[1] = -e[0] * sgn; qa[2] = -e[1] * sgn; qa[3] = -e[2] * sgn;
}
__device__ __forceinline__ void quat_mul(const double (&a)[4], const double (&b)[4], double (&o)[4]) {
    o[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    o[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    o[2] = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
    o[3] = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
}

// mode 0 (policy_transportation.py:61-77): mat = Jphi (M,3,3), out = quat(Jphi) (x) q.
// mode 1 (gaussian_process_transportation_diffeomorphic.py:94-101): mat = Jpsi at the ROTATED positions, out = quat(I + Jpsi) (x) (quat(Raff) (x) q).
__global__ void __launch_bounds__(128) quat_transport_kernel(const double* __restrict__ mat, const double* __restrict__ ori, long long M,
                                                             double* __restrict__ out, int mode, Affine af) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    double R[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) R[r][c] = mat[i * 9 + r * 3 + c] + ((mode == 1 && r == c) ? 1.0 : 0.0);
    double qa[4], qb[4] = {ori[i * 4 + 0], ori[i * 4 + 1], ori[i * 4 + 2], ori[i * 4 + 3]}, qo[4];
    quat_of_matrix(R, qa);
    if (mode == 1) {
        double Ra[3][3], qr[4], qt[4];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) Ra[r][c] = af.R[r][c];
        quat_of_matrix(Ra, qr);
        quat_mul(qr, qb, qt);
        quat_mul(qa, qt, qo);
    } else {
        quat_mul(qa, qb, qo);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) out[i * 4 + k] = qo[k];
}

// ------------------------------------------------------------------------------------------------------------
// Stiffness epilogue: K_hat = Jphi K Jphi^T per point (a stiffness / damping matrix is a (0,2)-type quantity of the demo
// frame; it follows the local linearisation of the map like the velocity does).  The reference code has no stiffness
// transport (its README announces it); for an orthogonal Jphi every candidate form reduces
// to this congruence.  One thread per point, everything in registers.
// ------------------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(128) stiffness_transport_kernel(const double* __restrict__ jphi, const double* __restrict__ Kin, long long M,
                                                                  double* __restrict__ Kout) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    double J[D][D], K[D][D], T[D][D];
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
        for (int c = 0; c < D; ++c) {
            J[r][c] = jphi[i * D * D + r * D + c];
            K[r][c] = Kin[i * D * D + r * D + c];
        }
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
        for (int c = 0; c < D; ++c) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < D; ++k) s = fma(J[r][k], K[k][c], s);
            T[r][c] = s;
        }
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
        for (int c = 0; c < D; ++c) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < D; ++k) s = fma(T[r][k], J[c][k], s);
            Kout[i * D * D + r * D + c] = s;
        }
}


// ------------------------------------------------------------------------------------------------------------
// Dense transport grids (BASELINE config 5; the query shape of plot_utils.py:10-15, 353-358): the lattice points are generated on
// the device from (origin, step, dims) -- point g of the row-major lattice (last dimension fastest) -- and the outputs are reduced to
// per-column statistics, so a 2^29-point grid needs neither host input nor host output buffers.
// ------------------------------------------------------------------------------------------------------------
struct GridSpec {
    double origin[MAXD], step[MAXD];
    long long dims[MAXD];
};

__global__ void __launch_bounds__(256) grid_points_kernel(GridSpec g, int d, long long first, int B, double* __restrict__ xq) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= B) return;
    long long idx = first + q;
    double x[MAXD];
    for (int a = d - 1; a >= 0; --a) {
        const long long i = idx % g.dims[a];
        idx /= g.dims[a];
        x[a] = __dadd_rn(g.origin[a], __dmul_rn(g.step[a], (double)i));      // no FMA contraction: bit-identical to numpy's origin + step * i
    }
    for (int a = 0; a < d; ++a) xq[(long long)q * d + a] = x[a];
}

// packed row of one lattice point: [mean (p) | std (1, the p columns are identical) | jac (p*d)], whichever were requested
__global__ void __launch_bounds__(256) grid_pack_kernel(const double* __restrict__ mean, const double* __restrict__ std, const double* __restrict__ jac,
                                                        int B, int p, int d, unsigned flags, double* __restrict__ pack) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= B) return;
    const int ncol = ((flags & 1u) ? p : 0) + ((flags & 2u) ? 1 : 0) + ((flags & 4u) ? p * d : 0);
    double* dst = pack + (long long)q * ncol;
    int c = 0;
    if (flags & 1u) for (int o = 0; o < p; ++o) dst[c++] = mean[(long long)q * p + o];
    if (flags & 2u) dst[c++] = std[(long long)q * p];
    if (flags & 4u) for (int v = 0; v < p * d; ++v) dst[c++] = jac[(long long)q * p * d + v];
}

// Column statistics of a row-major (B, ncol) block: partial[b][col] = {sum, sum of squares, min, max} over the rows block b owns
// (fixed assignment, fixed order: the result does not depend on scheduling).  GRID_STAT_BLOCKS blocks of 256 threads.
constexpr int GRID_STAT_BLOCKS = 128;
__global__ void __launch_bounds__(256) grid_stats_partial_kernel(const double* __restrict__ v, int B, int ncol, double* __restrict__ partial) {
    __shared__ double red[4][256];
    for (int col = 0; col < ncol; ++col) {
        double s = 0.0, s2 = 0.0, mn = 1e300, mx = -1e300;
        for (int r = blockIdx.x * 256 + threadIdx.x; r < B; r += GRID_STAT_BLOCKS * 256) {
            const double x = v[(long long)r * ncol + col];
            s += x;
            s2 = fma(x, x, s2);
            mn = fmin(mn, x);
            mx = fmax(mx, x);
        }
        red[0][threadIdx.x] = s; red[1][threadIdx.x] = s2; red[2][threadIdx.x] = mn; red[3][threadIdx.x] = mx;
        __syncthreads();
        for (int w = 128; w > 0; w >>= 1) {
            if (threadIdx.x < w) {
                red[0][threadIdx.x] += red[0][threadIdx.x + w];
                red[1][threadIdx.x] += red[1][threadIdx.x + w];
                red[2][threadIdx.x] = fmin(red[2][threadIdx.x], red[2][threadIdx.x + w]);
                red[3][threadIdx.x] = fmax(red[3][threadIdx.x], red[3][threadIdx.x + w]);
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            double* dst = partial + ((long long)blockIdx.x * ncol + col) * 4;
            dst[0] = red[0][0]; dst[1] = red[1][0]; dst[2] = red[2][0]; dst[3] = red[3][0];
        }
        __syncthreads();
    }
}
// acc[col] (+)= fixed-order reduction of the block partials; first != 0 initialises the accumulator
__global__ void grid_stats_final_kernel(const double* __restrict__ partial, int ncol, double* __restrict__ acc, int first) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol) return;
    double s = 0.0, s2 = 0.0, mn = 1e300, mx = -1e300;
    for (int b = 0; b < GRID_STAT_BLOCKS; ++b) {
        const double* src = partial + ((long long)b * ncol + col) * 4;
        s += src[0]; s2 += src[1]; mn = fmin(mn, src[2]); mx = fmax(mx, src[3]);
    }
    double* dst = acc + (long long)col * 4;
    if (first) { dst[0] = s; dst[1] = s2; dst[2] = mn; dst[3] = mx; }
    else { dst[0] += s; dst[1] += s2; dst[2] = fmin(dst[2], mn); dst[3] = fmax(dst[3], mx); }
}
// every `stride`-th lattice point of the block [first, first + B) -> packed sample rows (ncol each)
__global__ void grid_sample_kernel(const double* __restrict__ v, int ncol, long long first, int B, long long stride, long long first_sample,
                                   double* __restrict__ out) {
    // sample j is lattice point j * stride; this block holds samples j with first <= j*stride < first + B
    const long long j0 = (first + stride - 1) / stride;
    const long long j = j0 + blockIdx.x * blockDim.x + threadIdx.x;
    const long long g = j * stride;
    if (g >= first + B) return;
    for (int c = 0; c < ncol; ++c) out[(j - first_sample) * ncol + c] = v[(g - first) * ncol + c];
}


// ------------------------------------------------------------------------------------------------------------
// Minimum-variance stabilised rollout (plot_utils.py:298-310): pos <- pos + mean(pos) - gain * std(pos) * g / |g| with
// g = derivative_of_variance(pos) (gaussian_process.py:104-126); one thread per start point, the step is recorded in traj.
// ------------------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(128) rollout_step_kernel(double* __restrict__ pos, const double* __restrict__ mean, const double* __restrict__ std,
                                                           const double* __restrict__ dvar, long long K, double gain, double* __restrict__ traj_t) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    double g[D], nrm = 0.0;
#pragma unroll
    for (int a = 0; a < D; ++a) {
        g[a] = dvar[(long long)a * K + k];
        nrm = fma(g[a], g[a], nrm);
    }
    nrm = sqrt(nrm);
    const double sd = std[k * D];
#pragma unroll
    for (int a = 0; a < D; ++a) {
        const double x = pos[k * D + a] + mean[k * D + a] - gain * sd * (g[a] / nrm);
        pos[k * D + a] = x;
        traj_t[k * D + a] = x;
    }
}

}  // namespace gptb
