"""Host glue for the fused feature-propagation + head kernel (g4d_fp_interp_mlp): parameter folding/packing and the call."""
import ctypes

import torch
import torch.nn as nn

from .. import _lib
from . import pytorch_utils as pt_utils


def _fold_conv1d_block(block):
    """(W [out,in], b [out], relu?) of a pytorch_utils.Conv1d block: conv [+ eval BN] [+ ReLU]; None if not that shape."""
    kids = dict(block.named_children())
    conv = kids.get("conv")
    if conv is None or list(block.named_children())[0][0] != "conv" or "in" in kids:
        return None
    if tuple(conv.kernel_size) != (1,) or tuple(conv.stride) != (1,) or tuple(conv.padding) != (0,):
        return None
    w = conv.weight.detach().reshape(conv.out_channels, conv.in_channels).float()
    b = conv.bias.detach().float() if conv.bias is not None else torch.zeros(conv.out_channels, device=w.device)
    if "bn" in kids:
        bn = kids["bn"][0]
        scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
        w = w * scale[:, None]
        b = (b - bn.running_mean.detach().float()) * scale + bn.bias.detach().float()
    act = kids.get("activation")
    if act is not None and not isinstance(act, nn.ReLU):
        return None
    return w.contiguous(), b.contiguous(), act is not None


class PackedFpHead:
    def __init__(self, desc, params, c2, h2):
        self.desc, self.params, self.c2, self.h2 = desc, params, c2, h2


def pack_fp_head(fp_mlp, fc_layer, device):
    """Packs PointnetFPModule.mlp (2 conv+BN+ReLU blocks) and FC_layer = [Conv1d+BN+ReLU, Dropout, Conv1d] for the kernel.
    Returns None when the stacks do not have that shape."""
    folded = pt_utils.fold_shared_mlp(fp_mlp)
    if folded is None or len(folded) != 2:
        return None
    blocks = [m for m in fc_layer if not isinstance(m, nn.Dropout)]
    if len(blocks) != 2:
        return None
    ha, hb = _fold_conv1d_block(blocks[0]), _fold_conv1d_block(blocks[1])
    if ha is None or hb is None or not ha[2] or hb[2]:
        return None
    (w1, b1), (w2, b2) = folded
    c1, c_in = w1.shape
    c2 = w2.shape[0]
    h1, h2 = ha[0].shape[0], hb[0].shape[0]
    ok16 = lambda v: 16 <= v <= 256 and v % 16 == 0
    if not (ok16(c_in) and ok16(c1) and ok16(c2) and ok16(h1) and 1 <= h2 <= 16):
        return None
    L = _lib.lib()
    desc = _lib.FpDesc(c_in, c1, c2, h1, h2)
    nbytes = L.g4d_fp_param_bytes(ctypes.byref(desc))
    if nbytes == 0:
        raise _lib.G4DError("g4d_fp_param_bytes: " + L.g4d_last_error().decode())
    host = [t.cpu().contiguous() for t in (w1, b1, w2, b2, ha[0], ha[1], hb[0], hb[1])]
    blob = torch.empty(nbytes, dtype=torch.uint8)
    rc = L.g4d_fp_pack_params(ctypes.byref(desc), *(t.data_ptr() for t in host), blob.data_ptr())
    _lib.check(rc, "g4d_fp_pack_params")
    return PackedFpHead(desc, blob.to(device), c2, h2)


def fp_interp_mlp(packed, unknown, known, known_feats):
    """unknown (B,n,3), known (B,m,3), known_feats (B,C,m) fp32 -> (features (B,c2,n) fp32, logits (B,n,h2) fp32).
    The kernel's last epilogue also writes the arg-max labels (B,n) uint8; they ride on the logits tensor (segmentation_labels())."""
    from .pointnet2_utils import three_nn_raw, point_major_of
    unknown = unknown if unknown.is_contiguous() else unknown.contiguous()
    known = known if known.is_contiguous() else known.contiguous()
    B, n, _ = unknown.shape
    m = known.shape[1]
    dist2 = torch.empty(B, n, 3, dtype=torch.float32, device=unknown.device)
    idx = torch.empty(B, n, 3, dtype=torch.int32, device=unknown.device)
    three_nn_raw(unknown, known, dist2, idx)
    known_pm = point_major_of(known_feats)       # emitted by the producing FP level's epilogue; None when stale
    if known_pm is None:
        known_pm = known_feats.detach().transpose(1, 2).to(torch.float16).contiguous()
    feat = torch.empty(B, packed.c2, n, dtype=torch.float32, device=unknown.device)
    logits = torch.empty(B, n, packed.h2, dtype=torch.float32, device=unknown.device)
    labels = torch.empty(B, n, dtype=torch.uint8, device=unknown.device)
    rc = _lib.lib().g4d_fp_interp_mlp_labels(ctypes.byref(packed.desc), _lib.ptr(packed.params), B, n, m, _lib.ptr(dist2), _lib.ptr(idx),
                                             _lib.ptr(known_pm), _lib.ptr(feat), _lib.ptr(logits), _lib.ptr(labels), _lib.stream_ptr())
    _lib.check(rc, "g4d_fp_interp_mlp_labels")
    try:
        logits._g4d_labels = (labels, logits._version)
    except Exception:
        pass
    return feat, logits


def segmentation_labels(sem_logits):
    """argmax(sem_logits, dim=2) as uint8 (mesh_encoder.py:113): the labels the fused head kernel wrote next to these logits when
    they are still current, otherwise computed here."""
    hit = getattr(sem_logits, "_g4d_labels", None)
    if hit is not None and hit[1] == sem_logits._version and hit[0].shape == sem_logits.shape[:2]:
        return hit[0]
    return sem_logits.argmax(dim=2).to(torch.uint8)
