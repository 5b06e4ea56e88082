"""Seeded synthetic inputs of the reference's shapes (no dataset or licensed SMPL file is available offline):
SMPL-shaped constants (V=6890, J=24, 10 betas, 207 pose features; distributions from SURVEY.md section 8(d)) and
point clouds.  numpy only; deterministic in the seed (legacy RandomState streams are stable across versions)."""
import numpy as np

# Standard SMPL kinematic tree (kintree_table[0] of the SMPL pkl, read at smplx/smplx/body_models.py:245-247;
# the table itself is not in the reference repository).
SMPL_PARENTS = np.array([-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21],
                        dtype=np.int64)
f32 = np.float32


def synthetic_smpl(V=6890, J=24, nbetas=10, seed=0, sparse_weights=True):
    rs = np.random.RandomState(seed)
    v_template = (rs.randn(V, 3) * 0.3).astype(f32)
    shapedirs = (rs.randn(V, 3, nbetas) * 0.01).astype(f32)
    posedirs = (rs.randn((J - 1) * 9, V * 3) * 0.001).astype(f32)
    Jr = rs.rand(J, V).astype(np.float64) ** 8       # peaky, row-stochastic like the real regressor
    J_regressor = (Jr / Jr.sum(1, keepdims=True)).astype(f32)
    if sparse_weights:
        W = np.zeros((V, J), np.float64)
        cols = np.stack([rs.permutation(J)[:4] for _ in range(V)])
        vals = rs.rand(V, 4) + 0.05
        vals /= vals.sum(1, keepdims=True)
        np.put_along_axis(W, cols, vals, axis=1)
    else:
        W = rs.rand(V, J) + 0.01
        W /= W.sum(1, keepdims=True)
    parents = SMPL_PARENTS[:J].copy()
    return dict(v_template=v_template, shapedirs=shapedirs, posedirs=posedirs, J_regressor=J_regressor,
                parents=parents, lbs_weights=W.astype(f32))


def synthetic_frames(F, J=24, nbetas=10, seed=1):
    rs = np.random.RandomState(seed)
    betas = rs.randn(F, nbetas).astype(f32)
    pose = (rs.randn(F, J * 3) * 0.3).astype(f32)
    return betas, pose


def body_clouds(seed, C, N, dup_frac=0.05):
    """C clouds of N points on a body-like 2-D surface (union of 6 capsules: torso, head, 2 arms, 2 legs; height ~1.7,
    coordinates in metres like CLOTH3D), shuffled, with a fraction of exact duplicates (the reference loader
    oversamples, utils/dataloader.py:35-44)."""
    rs = np.random.RandomState(seed)
    # capsule: (centre xyz, axis, half-length, radius), weights ~ area
    caps = [((0.0, 1.15, 0.0), 1, 0.30, 0.15), ((0.0, 1.62, 0.0), 1, 0.06, 0.10),
            ((-0.45, 1.38, 0.0), 0, 0.28, 0.045), ((0.45, 1.38, 0.0), 0, 0.28, 0.045),
            ((-0.10, 0.42, 0.0), 1, 0.40, 0.07), ((0.10, 0.42, 0.0), 1, 0.40, 0.07)]
    area = np.array([2 * np.pi * r * (2 * h + 2 * r) for _, _, h, r in caps])
    out = np.empty((C, N, 3), f32)
    for c in range(C):
        which = rs.choice(len(caps), size=N, p=area / area.sum())
        u = rs.rand(N) * 2 * np.pi
        t = rs.rand(N) * 2 - 1
        pts = np.zeros((N, 3))
        for i, (ctr, ax, h, r) in enumerate(caps):
            sel = which == i
            a, b = [d for d in range(3) if d != ax]
            p = np.zeros((sel.sum(), 3))
            p[:, ax] = t[sel] * (h + r)
            shrink = np.clip((np.abs(p[:, ax]) - h) / r, 0, 1)          # hemispherical caps
            rr = r * np.sqrt(np.clip(1 - shrink ** 2, 0, 1))
            p[:, a] = rr * np.cos(u[sel])
            p[:, b] = rr * np.sin(u[sel])
            pts[sel] = p + np.array(ctr)
        pts += rs.randn(1, 3) * 0.02                                     # per-frame jitter of the whole body
        nd = int(N * dup_frac)
        if nd:
            pts[rs.randint(0, N, nd)] = pts[rs.randint(0, N, nd)]
        out[c] = pts.astype(f32)
    return out


def cube_clouds(seed, C, N):
    return np.random.RandomState(seed).rand(C, N, 3).astype(f32)
