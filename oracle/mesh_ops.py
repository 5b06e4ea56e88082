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
