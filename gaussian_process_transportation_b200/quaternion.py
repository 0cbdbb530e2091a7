"""Minimal quaternion algebra for the orientation epilogue (policy_transportation.py:61-77), standing in for the
numpy-quaternion calls `from_rotation_matrix(rot, nonorthogonal=True)` and `*` (dependency absent from the reference
tree and this image: restated from the published Bar-Itzhack algorithm; SURVEY.md App. A.6b, parity unpinned).
Batched over leading axes; quaternions are (w, x, y, z)."""
import numpy as np


def from_rotation_matrix_nonorthogonal(rot):
    rot = np.asarray(rot, dtype=np.float64)
    shape = rot.shape[:-2]
    R = rot.reshape(-1, 3, 3)
    K3 = np.empty((R.shape[0], 4, 4))
    K3[:, 0, 0] = (R[:, 0, 0] - R[:, 1, 1] - R[:, 2, 2]) / 3.0
    K3[:, 0, 1] = (R[:, 1, 0] + R[:, 0, 1]) / 3.0
    K3[:, 0, 2] = (R[:, 2, 0] + R[:, 0, 2]) / 3.0
    K3[:, 0, 3] = (R[:, 1, 2] - R[:, 2, 1]) / 3.0
    K3[:, 1, 1] = (R[:, 1, 1] - R[:, 0, 0] - R[:, 2, 2]) / 3.0
    K3[:, 1, 2] = (R[:, 2, 1] + R[:, 1, 2]) / 3.0
    K3[:, 1, 3] = (R[:, 2, 0] - R[:, 0, 2]) / 3.0
    K3[:, 2, 2] = (R[:, 2, 2] - R[:, 0, 0] - R[:, 1, 1]) / 3.0
    K3[:, 2, 3] = (R[:, 0, 1] - R[:, 1, 0]) / 3.0
    K3[:, 3, 3] = (R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2]) / 3.0
    for i in range(4):
        for j in range(i):
            K3[:, i, j] = K3[:, j, i]
    _, vecs = np.linalg.eigh(K3)                      # ascending eigenvalues: last column = dominant eigenvector
    e = vecs[:, :, 3]
    q = np.stack([e[:, 3], -e[:, 0], -e[:, 1], -e[:, 2]], axis=-1)
    return q.reshape(shape + (4,))


def multiply(a, b):
    aw, ax, ay, az = np.moveaxis(np.asarray(a, dtype=np.float64), -1, 0)
    bw, bx, by, bz = np.moveaxis(np.asarray(b, dtype=np.float64), -1, 0)
    return np.stack([aw * bw - ax * bx - ay * by - az * bz,
                     aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw], axis=-1)
