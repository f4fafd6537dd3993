"""TEST INFRASTRUCTURE ONLY.

Restatement of the third-party `thin-plate-spline==1.2.0` package (requirements.txt:19; imported
as `tps` at BDRC/image_dewarping.py:17, used at :61-65).  The package is absent from this image
and is not vendored by the reference, so its published algorithm is restated here:

    U(r) = r^2 log r  (U(0) = 0),   [[K + alpha*I, P], [P^T, 0]] [w; a] = [Y; 0],   P = [1, X]
    transform(X) = [U(|X - c_j|), 1, X] @ parameters

Parity: UNPINNED (the reference has no test or golden vector for it, and its only call site,
run_tps, raises TypeError before reaching it: image_dewarping.py:42).
"""
from __future__ import annotations

import numpy as np


class ThinPlateSpline:
    def __init__(self, alpha: float = 0.0):
        self.alpha = float(alpha)
        self.control_points = None
        self.parameters = None

    def _radial(self, X: np.ndarray) -> np.ndarray:
        d = np.sqrt(((X[:, None, :] - self.control_points[None, :, :]) ** 2).sum(-1))
        d[d == 0] = 1.0
        return d * d * np.log(d)

    def fit(self, X, Y):
        X = np.asarray(X, dtype=np.float64).reshape(len(X), -1)
        Y = np.asarray(Y, dtype=np.float64).reshape(len(Y), -1)
        n, ds = X.shape
        dt = Y.shape[1]
        self.control_points = X
        phi = self._radial(X)
        Xp = np.hstack([np.ones((n, 1)), X])
        A = np.vstack([
            np.hstack([phi + self.alpha * np.identity(n), Xp]),
            np.hstack([Xp.T, np.zeros((ds + 1, ds + 1))]),
        ])
        Yp = np.vstack([Y, np.zeros((ds + 1, dt))])
        self.parameters = np.linalg.solve(A, Yp)
        return self

    def transform(self, X):
        X = np.asarray(X, dtype=np.float64).reshape(len(X), -1)
        phi = self._radial(X)
        return np.hstack([phi, np.ones((len(X), 1)), X]) @ self.parameters
