"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the model-side callers of the hot path (modules/mesh_encoder.py)."""
import numpy as np


def calc_segmentation_results(x, sem_logits, n, feature, garment_label):
    """mesh_encoder.py:109-125, frame by frame like the reference's Python loop.
    x (C,N,3), sem_logits (C,N,cls), feature (C,Cf,N) channel-major -> garment_v (C,n,3), feat (C,n,Cf)."""
    C, N, _ = x.shape
    feature = feature.transpose(0, 2, 1)                      # :111  feature.transpose(1, 2)
    labels = np.argmax(sem_logits, axis=2)                    # :113  (first maximal class, like torch.argmax)
    gv = np.zeros((C, n, 3), np.float32)
    gf = np.zeros((C, n, feature.shape[2]), np.float32)
    for i in range(C):                                        # :116-124
        sel = labels[i] == garment_label
        cur_x, cur_f = x[i][sel], feature[i][sel]
        k = min(n, cur_x.shape[0])
        gv[i, :k] = cur_x[:k]
        gf[i, :k] = cur_f[:k]
    return gv, gf


def knn_points(p1, p2, K=1):
    """chamferdist.knn_points as used at mesh_encoder.py:321-324 (pytorch3d-style: squared distances ascending, their indices).
    PARITY UNPINNED: chamferdist is an un-vendored, unpinned dependency of the reference and is not installed here; this
    restates its published contract.  Order among equal distances: ascending index (this repository's choice).
    p1 (B,N,3), p2 (B,P,3) -> dists (B,N,K) fp32, idx (B,N,K) int64."""
    B, N, _ = p1.shape
    d = np.empty((B, N, K), np.float32)
    ix = np.empty((B, N, K), np.int64)
    for b in range(B):
        dx = p1[b, :, None, 0] - p2[b, None, :, 0]
        dy = p1[b, :, None, 1] - p2[b, None, :, 1]
        dz = p1[b, :, None, 2] - p2[b, None, :, 2]
        # dist += diff * diff in x, y, z order with fp32 fused multiply-adds (emulated in float64: exact products, one rounding each)
        acc = (dx * dx).astype(np.float32)
        acc = (dy.astype(np.float64) * dy.astype(np.float64) + acc.astype(np.float64)).astype(np.float32)
        acc = (dz.astype(np.float64) * dz.astype(np.float64) + acc.astype(np.float64)).astype(np.float32)
        order = np.argsort(acc, axis=1, kind="stable")[:, :K]
        ix[b] = order
        d[b] = np.take_along_axis(acc, order, axis=1)
    return d, ix


def _interp_weights(dists):
    """mesh_encoder.py:341-345 / :371-375 on (B,N,K): 1/d, inf -> 0, normalise over K, inf -> 0."""
    with np.errstate(divide="ignore", invalid="ignore"):
        w = (np.float32(1.0) / dists).astype(np.float32)
        w[np.isinf(w)] = 0
        w = (w / w.sum(-1, keepdims=True, dtype=np.float32)).astype(np.float32)
        w[np.isinf(w)] = 0
    return w


def lbs_garment_interpolation(pred_template_garment_v, Tpose_vertices, Tpose_root_joints, zeropose_vertices, parents, gt_pose,
                              T_J_regressor, T_lbs_weights, smooth_adj, K=3, smooth_iters=100, coeff=0.1):
    """MeshEncoder.lbs_garment_interpolation (mesh_encoder.py:312-410) in numpy fp32.  PARITY UNPINNED (see knn_points).
    pred_template_garment_v (B,G,3), Tpose_vertices (B,P,3), Tpose_root_joints (B,3), zeropose_vertices (B,T,P,3), gt_pose (B,T,72),
    T_J_regressor (B,T,J,P), T_lbs_weights (B,T,P,J), smooth_adj: dense (G,G) = normalize(adj_old) - I (:386).
    -> lbs_pred (B,T,G,3), (nn_dists (B,G,1), nn_idx (B,G,1)), stage1 (B,T,G,3)"""
    from . import lbs as L
    B, G, _ = pred_template_garment_v.shape
    T = gt_pose.shape[1]
    J = T_J_regressor.shape[2]
    gt_pose_mat = L.batch_rodrigues(gt_pose.reshape(-1, 3)).reshape(B * T, 24, 3, 3)                        # :318
    q = (pred_template_garment_v + Tpose_root_joints.reshape(B, 1, 3)).astype(np.float32)                   # :320
    body = Tpose_vertices.reshape(B, -1, 3)
    dK, iK = knn_points(q, body, K)                                                                          # :321
    K64 = min(64, K)
    d64, i64 = dK[:, :, :K64], iK[:, :, :K64]                                                                # :323 (a prefix of the K result)
    nn = (dK[:, :, :1].copy(), iK[:, :, :1].copy())                                                          # :324
    inv_pose = np.zeros((B, 24, 3), np.float32)                                                              # :326-329
    inv_pose[:, 0, 0] = -np.pi / 2
    inv_pose[:, 1, 1] = 0.15
    inv_pose[:, 2, 1] = -0.15
    inv_mat = L.batch_rodrigues(inv_pose.reshape(-1, 3)).reshape(B, 24, 3, 3)
    inv_J = L.vertices2jointsB(T_J_regressor[:, 0], body)                                                    # :333
    _, inv_A = L.batch_rigid_transform(inv_mat, inv_J, parents)                                              # :335
    w64 = _interp_weights(d64)                                                                               # :341-345
    W0 = T_lbs_weights[:, 0]                                                                                 # (B,P,J)
    inv_nn_W = np.zeros((B, G, J), np.float32)
    for b in range(B):
        inv_nn_W[b] = (W0[b][i64[b]] * w64[b][:, :, None]).sum(1, dtype=np.float32)                          # :340,346
    stage1_b = L.skin(q, inv_A, inv_nn_W)                                                                    # :347-362
    stage1 = np.repeat(stage1_b[:, None], T, axis=1).reshape(B * T, G, 3)
    Jz = L.vertices2jointsB(T_J_regressor.reshape(B * T, J, -1), zeropose_vertices.reshape(B * T, -1, 3))     # :366-368
    _, A = L.batch_rigid_transform(gt_pose_mat, Jz, parents)                                                 # :369
    wK = _interp_weights(dK)                                                                                 # :371-375
    Wf = T_lbs_weights.reshape(B * T, -1, J)
    nn_W = np.zeros((B * T, G, J), np.float32)
    for f in range(B * T):
        b = f // T
        nn_W[f] = (Wf[f][iK[b]] * wK[b][:, :, None]).sum(1, dtype=np.float32)                                # :377-379
    if K > 1:                                                                                                # :382-389
        adj = smooth_adj.astype(np.float32)
        for _ in range(smooth_iters):
            nn_W = (nn_W + np.float32(coeff) * np.einsum("gh,fhj->fgj", adj, nn_W).astype(np.float32)).astype(np.float32)
    out = L.skin(stage1, A, nn_W)                                                                            # :391-408
    return out.reshape(B, T, G, 3), nn, stage1.reshape(B, T, G, 3)
