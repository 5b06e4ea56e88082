"""Shared synthetic inputs for the tests (seeded)."""
import numpy as np


def clouds(seed, B, N, kind="cube", dup_frac=0.05):
    """'cube': U[0,1)^3.  'body': points on a wavy tube of height 1.7 (a 2-D surface, like real scans).
    A fraction of exact duplicate points exercises the FPS tie-breaks (the reference loader oversamples,
    utils/dataloader.py:35-44)."""
    rs = np.random.RandomState(seed)
    if kind == "cube":
        x = rs.rand(B, N, 3).astype(np.float32)
    else:
        u = rs.rand(B, N).astype(np.float32) * np.float32(2 * np.pi)
        h = rs.rand(B, N).astype(np.float32) * np.float32(1.7)
        r = np.float32(0.15) + np.float32(0.05) * np.sin(h * np.float32(7.0)).astype(np.float32)
        x = np.stack([r * np.cos(u), h, r * np.sin(u)], axis=-1).astype(np.float32)
    nd = int(N * dup_frac)
    for b in range(B):
        if nd:
            src = rs.randint(0, N, nd)
            dst = rs.randint(0, N, nd)
            x[b, dst] = x[b, src]
    return x
