"""CPU restatement (numpy) of the 8-bit digit-plane split used by the INT8-sliced variance path
(gaussian_process_transportation_b200/csrc/digits.cuh: digits8_pack4, digit_scale8).  TEST INFRASTRUCTURE: it documents and checks
the arithmetic identity the CUDA code relies on; nothing in the product imports it.

    x  ~  scale * Q / 256^S,   Q = rint(x / scale * 256^S),   Q = sum_{t=0}^{S-1} d_t * 256^(S-1-t),   d_t in [-128, 127]

The balanced digits are the bytes of (Q + C) xor C with C = 0x80..80 (S bytes): adding 128 per byte position turns signed digits
into unsigned bytes with the carries resolved by the integer add, the xor maps them back."""
import numpy as np


def digit_scale8(bound):
    return bound / 0.498 if bound > 0 else 1.0


def split8(x, S, scale):
    x = np.asarray(x, dtype=np.float64)
    q = np.rint(x * (256.0 ** S / scale)).astype(np.int64)
    c = sum(0x80 << (8 * k) for k in range(S))
    u = (q + c) ^ c
    planes = np.stack([((u >> (8 * (S - 1 - t))) & 0xFF).astype(np.uint8).view(np.int8) for t in range(S)])   # plane 0 = most significant
    return planes, q


def combine8(planes):
    S = planes.shape[0]
    return sum(planes[t].astype(np.int64) * (256 ** (S - 1 - t)) for t in range(S))


def sliced_dot(a_planes, b_planes, sa, sb):
    """sum_k a_k b_k from the plane products with a + b <= S - 1 (0-based), exact integer accumulation, FP64 recombination."""
    S = a_planes.shape[0]
    acc = 0.0
    for a in range(S):
        for b in range(S - a):
            acc += float(np.dot(a_planes[a].astype(np.int64), b_planes[b].astype(np.int64))) * 2.0 ** (-8 * (a + b + 2))
    return acc * sa * sb
