"""Rigid(+scale) pre-alignment gamma(x) = s R (x - Sbar) + Tbar of paired point sets -- drop-in for
policy_transportation/models/affine_trasformation.py:8-57.  The O(N d^2) moments and the d x d SVD run in numpy on the host
(bit-parity with the reference); the *apply* is fused into the GPU query prologue when the
delta map is the B200 GaussianProcess, and is also available here for standalone use."""
import numpy as np


class AffineTransform():
    def __init__(self, do_scale=False, do_rotation=True):
        self.do_scale = do_scale
        self.do_rotation = do_rotation
        self.scale = 1

    def fit(self, source_points, target_points):
        assert len(source_points) == len(target_points)
        src = np.asarray(source_points)
        tgt = np.asarray(target_points)
        n, d = src.shape
        self.S_centroid = np.mean(src, axis=0)
        self.T_centroid = np.mean(tgt, axis=0)
        self.source_points_centered = src - self.S_centroid
        self.target_points_centered = tgt - self.T_centroid
        too_few = (d == 2 and n < 2) or (d == 3 and n < 3)
        if not self.do_rotation or too_few:
            self.rotation_matrix = np.eye(d)
        else:
            cross_cov = np.dot(self.source_points_centered.T, self.target_points_centered)
            U, _, Vt = np.linalg.svd(cross_cov)
            V = Vt.T
            self.rotation_matrix = V @ U.T
            if np.linalg.det(self.rotation_matrix) < 0:          # reflection fix
                V[:, -1] *= -1
                self.rotation_matrix = V @ U.T
        if self.do_scale:
            rotated = (self.rotation_matrix @ self.source_points_centered.T).T
            self.scale = np.sum(rotated * self.target_points_centered) / np.sum(rotated ** 2)
        print("Rotation Matrix of the Affine Matrix:")
        print(self.rotation_matrix)
        print("Scaling factor:", self.scale)
        self.translation = self.T_centroid - self.S_centroid

    def predict(self, x):
        return self.scale * (self.rotation_matrix @ (x - self.S_centroid).T).T + self.T_centroid

    def derivative(self, x):
        # the scale is not part of the Jacobian in the reference (quirk Q6)
        return np.repeat(self.rotation_matrix[np.newaxis, :, :], x.shape[0], axis=0)
