"""Shared helpers for the test-suite."""
import numpy as np


def unit_square_mesh(n: int, jitter: float = 0.0, seed: int = 0):
    """Structured triangulation of [0,1]^2 with n x n squares split alternately (2 n^2 cells)."""
    xs = np.linspace(0.0, 1.0, n + 1)
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    xy = np.stack([X.ravel(), Y.ravel()], axis=1)
    if jitter:
        rng = np.random.default_rng(seed)
        interior = (xy[:, 0] > 0) & (xy[:, 0] < 1) & (xy[:, 1] > 0) & (xy[:, 1] < 1)
        xy[interior] += jitter / n * rng.uniform(-1, 1, size=(interior.sum(), 2))
    idx = lambda i, j: i * (n + 1) + j  # noqa: E731
    tri = []
    for i in range(n):
        for j in range(n):
            a, b, c, d = idx(i, j), idx(i + 1, j), idx(i + 1, j + 1), idx(i, j + 1)
            if (i + j) % 2 == 0:
                tri += [(a, b, c), (a, c, d)]
            else:
                tri += [(a, b, d), (b, c, d)]
    return xy, np.array(tri, dtype=np.int32)


def near(a, b, tol=3e-16):
    return np.abs(a - b) <= tol
