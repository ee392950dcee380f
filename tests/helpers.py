import numpy as np


def square_mesh(n, jitter=0.1, seed=3):
    """unit square, n x n x 2 triangles, every third cell marked solid (2), jittered."""
    from turtlefsi_b200.mesh import Mesh
    rng = np.random.default_rng(seed)
    xs = np.linspace(0, 1, n + 1)
    P = np.array([(x, y) for x in xs for y in xs])
    ii, jj = np.meshgrid(np.arange(n + 1), np.arange(n + 1), indexing="ij")
    interior = ((ii > 0) & (ii < n) & (jj > 0) & (jj < n)).ravel()
    P[interior] += jitter / n * rng.uniform(-1, 1, size=(interior.sum(), 2))
    cells = []
    for i in range(n):
        for j in range(n):
            v0 = i * (n + 1) + j
            v1, v2, v3 = v0 + 1, v0 + n + 1, v0 + n + 2
            cells += [[v0, v1, v3], [v0, v2, v3]]
    cells = np.sort(np.array(cells, dtype=np.int32), axis=1)
    markers = np.where(np.arange(len(cells)) % 3 == 0, 2, 1).astype(np.int32)
    return Mesh(P, cells, markers)
