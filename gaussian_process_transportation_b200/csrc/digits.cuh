// Error-free digit cascade shared by the row slicer (ozaki.cuh) and the fused generator (query.cuh).
//
// A value y0 = x * 2^-e is written as  y0 = sum_{t=1..S} digit_t * B^-t + r_S * B^-S  with B = 2^BITS; every step
// (scaling by a power of two, rounding to an integer, subtracting it) is exact in FP64, so the identity holds exactly and
// the only error of the scheme is the dropped remainder r_S.
//   BITS = 7: balanced digits in [-64, 64]   (rint), needs |y0| <= 1/2;
//   BITS = 8: the full int8 range [-128, 127].  A non-redundant base-256 digit set has no slack, so the rounding point is
//             shifted by 1/510: remainders live in [-1/2 - 1/510, 1/2 - 1/510), digit = rint(256 r + 1/510) is in
//             [-128, 127] and the remainder interval maps into itself.  The clamp only catches the case where the FP64
//             sum 256 r + 1/510 rounds onto the tie 127.5 (the exact sum is below it), where 127 is the correct digit.
//             Needs y0 in [-128/255, 127/255]; digit_scale_exp() leaves that head-room.
#pragma once
#include <math.h>

namespace gptb {

template <int BITS>
__device__ __forceinline__ double digit_step(double& y) {
    static_assert(BITS == 7 || BITS == 8, "digit planes are 7-bit balanced or 8-bit full-range");
    double d;
    if constexpr (BITS == 7) {
        y *= 128.0;
        d = rint(y);
    } else {
        y *= 256.0;
        d = fmin(fmax(rint(y + (1.0 / 510.0)), -128.0), 127.0);
    }
    y -= d;
    return d;
}

// ------------------------------------------------------------------------------------------------------------
// 8-bit planes without the cascade: Q = rint(y0 * 256^S) is one DFMA (magic-number rounding: the integer lands in the
// mantissa of 1.5*2^52 + y0*256^S), and the balanced base-256 digits of Q are the bytes of (Q + C) xor C with
// C = 0x80..80 (S bytes): adding 128 per byte position turns signed digits into unsigned bytes with the carries resolved
// by the integer add, the xor maps them back to [-128, 127].  Same digit set and the same range of y0 as digit_step<8>
// (the two differ only in how the dropped remainder is rounded), but ~6 instructions per value instead of ~9 per digit,
// none of them on the FP64 conversion pipe.  Four values are transposed into one 32-bit word per plane (PRMT), plane 0
// = most significant digit.  mul = 2^-e * 256^S.
// ------------------------------------------------------------------------------------------------------------
template <int S>
__device__ __forceinline__ void digits8_pack4(const double (&val)[4], double mul, unsigned (&packed)[S]) {
    static_assert(S >= 1 && S <= 6, "8-bit digit planes: Q must fit the 51 mantissa bits below the magic constant");
    constexpr unsigned long long C = (S >= 8) ? ~0ull : (0x8080808080808080ull >> (8 * (8 - S)));
    constexpr long long MAGIC_BITS = 0x4338000000000000LL;          // bit pattern of 1.5 * 2^52
    unsigned lo[4], hi[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double t = fma(val[j], mul, 6755399441055744.0);
        const unsigned long long u = ((unsigned long long)(__double_as_longlong(t) - MAGIC_BITS) + C) ^ C;
        lo[j] = (unsigned)u;
        hi[j] = (unsigned)(u >> 32);
    }
    // byte transpose: word b of the result holds byte b of the four inputs
    auto transpose4 = [](const unsigned (&w)[4], unsigned (&o)[4]) {
        const unsigned p01l = __byte_perm(w[0], w[1], 0x5140), p01h = __byte_perm(w[0], w[1], 0x7362);
        const unsigned p23l = __byte_perm(w[2], w[3], 0x5140), p23h = __byte_perm(w[2], w[3], 0x7362);
        o[0] = __byte_perm(p01l, p23l, 0x5410);
        o[1] = __byte_perm(p01l, p23l, 0x7632);
        o[2] = __byte_perm(p01h, p23h, 0x5410);
        o[3] = __byte_perm(p01h, p23h, 0x7632);
    };
    unsigned bl[4], bh[4];
    transpose4(lo, bl);
    transpose4(hi, bh);
#pragma unroll
    for (int t = 0; t < S; ++t) {
        const int byte = S - 1 - t;                     // plane t = digit of weight 256^-(t+1) = byte S-1-t of Q
        packed[t] = (byte < 4) ? bl[byte] : bh[byte - 4];
    }
}

// 8-bit planes do not need a power-of-two scale: Q = rint(x * mul) is a single correctly-rounded FMA for ANY mul, the digits
// represent Q exactly, and x = Q / mul up to one more rounding (1e-16, far below the 256^-S truncation).  Using the tightest
// scale -- the row maximum lands exactly on the 0.498 head-room limit -- wins 0.5 bit on average per operand over the next
// power of two (std error 1.2e-8 -> 5.9e-9 in the N = 2048 Morton-ordered simulation).  Returns the scale (x = y0 * scale).
__host__ __device__ inline double digit_scale8(double bound) { return (bound > 0.0) ? bound / 0.498 : 1.0; }

// exponent e of the power-of-two row scale: |bound| * 2^-e fits the first digit's range (see above)
__host__ __device__ inline int digit_scale_exp(double bound, int bits) {
    if (!(bound > 0.0)) return 0;
    int ex = 0;
    const double m = frexp(bound, &ex);      // bound = m * 2^ex, m in [0.5, 1)  ->  bound * 2^-(ex+1) = m/2 < 1/2
    int e = ex + 1;
    if (bits == 8 && m * 0.5 > 0.498) e += 1;
    return e;
}

}  // namespace gptb
