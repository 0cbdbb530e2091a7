"""Developer study (CPU, numpy): where the std error of the 8-bit digit-plane scheme comes from -- operand truncation vs the dropped
plane-pair products.  Result at N = 4096 (Morton order + shuffled chunks, benchmark hyper-parameters; profiles/r02_plane_error_study.log):
the operands truncated to 40 bits cost 5.5e-10, the 15-product scheme 9.0e-9 -- the error is the dropped diagonal a + b = 5, and adding
its four products (19 in all) recovers 5.5e-10."""
import sys, numpy as np, scipy.linalg as sla
sys.path.insert(0, '/root/repo')
from oracle.digit_planes import split8, digit_scale8
from bench import synthetic_pairs, kabsch
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
c, ell, s2 = 0.1, 0.1, 1e-4
S_, T_ = synthetic_pairs(N, 3, seed=0)
R, Sc, Tc = kabsch(S_, T_); X = (R @ (S_ - Sc).T).T + Tc
# Morton order + chunk shuffle (like gptb_set_train in spatial mode; the shuffle permutation differs, the structure is the same)
lo, hi = X.min(0), X.max(0)
cell = np.clip(((X - lo) / (hi - lo) * 1023).astype(np.int64), 0, 1023)
code = np.zeros(N, dtype=np.int64)
for b in range(10):
    for a in range(3):
        code |= ((cell[:, a] >> b) & 1) << (3 * b + a)
perm = np.argsort(code, kind='stable')
chunks = perm.reshape(-1, 64)
rng = np.random.default_rng(1); chunks = chunks[rng.permutation(len(chunks))]
X = X[chunks.ravel()]
def kern(A, B):
    d2 = ((A[:, None, :] - B[None, :, :]) ** 2).sum(-1) / ell ** 2
    return c * np.exp(-0.5 * d2)
K = kern(X, X); K[np.diag_indices(N)] += s2 + 1e-10
L = np.linalg.cholesky(K)
Li = sla.solve_triangular(L, np.eye(N), lower=True)
rq = np.random.default_rng(7)
xq = np.vstack([X.min(0) + (X.max(0) - X.min(0)) * rq.random((256, 3)), X[rq.choice(N, 256, replace=False)] + 1e-3 * rq.standard_normal((256, 3))])
ks = kern(xq, X)                       # (M, N)
W = ks @ Li.T                          # w[q, i] = sum_k ks[q,k] Li[i,k]
var = c + s2 - (W ** 2).sum(1)
std = np.sqrt(np.maximum(var, 0))
S = 5
sa = digit_scale8(c)
Ap, Aq = split8(ks, S, sa)             # planes (S, M, N)
Bp = np.zeros((S, N, N), dtype=np.int8); Bq = np.zeros((N, N)); sb = np.zeros(N)
for i in range(N):
    sb[i] = digit_scale8(np.abs(Li[i]).max())
    p, q = split8(Li[i], S, sb[i]); Bp[:, i, :] = p; Bq[i] = q
At = Aq * (sa / 256.0 ** S); Bt = Bq * (sb[:, None] / 256.0 ** S)
def stderr(Wx):
    v = c + s2 - (Wx ** 2).sum(1)
    return np.max(np.abs(np.sqrt(np.maximum(v, 0)) - std)) / np.sqrt(c + s2)
print("row scale of L^-1: median", np.median(sb), "max", sb.max(), " diag/|offdiag|max median:", np.median(np.abs(np.diag(Li)) / np.maximum(np.abs(Li - np.diag(np.diag(Li))).max(1), 1e-300)))
print("A truncated only      :", stderr(At @ Li.T))
print("B truncated only      :", stderr(ks @ Bt.T))
print("both truncated, all 25:", stderr(At @ Bt.T))
def pairs_sum(maxd):
    acc = np.zeros((len(xq), N))
    for a in range(S):
        for b in range(S):
            if a + b <= maxd:
                acc += (Ap[a].astype(np.float64) @ Bp[b].astype(np.float64).T) * 2.0 ** (-8 * (a + b + 2))
    return acc * sa * sb[None, :]
for maxd in (4, 5, 6):
    print(f"pairs a+b <= {maxd} ({sum(1 for a in range(S) for b in range(S) if a+b<=maxd)} products):", stderr(pairs_sum(maxd)))
# variant: diagonal of L^-1 handled in FP64, planes scaled by the off-diagonal row maximum
Lo = Li - np.diag(np.diag(Li))
sbo = np.array([digit_scale8(np.abs(Lo[i]).max()) for i in range(N)])
Bq2 = np.rint(Lo * (256.0 ** S / sbo[:, None])); Bt2 = Bq2 * (sbo[:, None] / 256.0 ** S)
print("B off-diagonal planes + exact diagonal, A truncated, all pairs:", stderr(At @ Bt2.T + ks * np.diag(Li)[None, :]))
