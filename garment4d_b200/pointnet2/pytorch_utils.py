"""Layer containers with the reference's constructor signatures and state-dict key names
(modules/pointnet2/pointnet2/pytorch_utils.py), so reference checkpoints load unchanged:

    <mlp>.layer{i}.conv.weight                      [out, in, 1, 1]   (no bias when bn=True)
    <mlp>.layer{i}.bn.bn.{weight,bias,running_mean,running_var,num_batches_tracked}

``SharedMLP`` = stack of 1x1 ``Conv2d`` -> ``BatchNorm2d`` -> ``ReLU`` blocks (pytorch_utils.py:5-32).  In eval
mode the set-abstraction modules do not run these layers through cuDNN at all: they fold conv+BN into one
affine map per layer (``fold_shared_mlp``) and hand the result to the tcgen05 grouped-MLP kernel.
"""
from typing import List, Tuple

import torch
import torch.nn as nn


class _NormWrap(nn.Sequential):
    """One norm layer registered under the name '<name>bn' (reference: _BNBase, pytorch_utils.py:104-111)."""

    def __init__(self, channels, norm_cls, name=""):
        super().__init__()
        self.add_module(name + "bn", norm_cls(channels))
        nn.init.constant_(self[0].weight, 1.0)
        nn.init.constant_(self[0].bias, 0)


class BatchNorm1d(_NormWrap):
    def __init__(self, in_size: int, *, name: str = ""):
        super().__init__(in_size, nn.BatchNorm1d, name)


class BatchNorm2d(_NormWrap):
    def __init__(self, in_size: int, name: str = ""):
        super().__init__(in_size, nn.BatchNorm2d, name)


class _ConvBase(nn.Sequential):
    """conv [+ bn] [+ activation] [+ instance norm], or the pre-activation order (pytorch_utils.py:35-101)."""

    def __init__(self, in_size, out_size, kernel_size, stride, padding, activation, bn, init, conv=None,
                 batch_norm=None, bias=True, preact=False, name="", instance_norm=False, instance_norm_func=None):
        super().__init__()
        use_bias = bias and not bn
        conv_unit = conv(in_size, out_size, kernel_size=kernel_size, stride=stride, padding=padding, bias=use_bias)
        init(conv_unit.weight)
        if use_bias:
            nn.init.constant_(conv_unit.bias, 0)
        norm_channels = in_size if preact else out_size
        tail = []
        if bn:
            tail.append((name + "bn", batch_norm(norm_channels)))
        if activation is not None:
            tail.append((name + "activation", activation))
        if not bn and instance_norm:
            tail.append((name + "in", instance_norm_func(norm_channels, affine=False, track_running_stats=False)))
        if preact:
            for k, mod in tail:
                self.add_module(k, mod)
        self.add_module(name + "conv", conv_unit)
        if not preact:
            for k, mod in tail:
                self.add_module(k, mod)


class Conv1d(_ConvBase):
    def __init__(self, in_size: int, out_size: int, *, kernel_size: int = 1, stride: int = 1, padding: int = 0,
                 activation=nn.ReLU(inplace=True), bn: bool = False, init=nn.init.kaiming_normal_, bias: bool = True,
                 preact: bool = False, name: str = "", instance_norm=False):
        super().__init__(in_size, out_size, kernel_size, stride, padding, activation, bn, init, conv=nn.Conv1d,
                         batch_norm=BatchNorm1d, bias=bias, preact=preact, name=name, instance_norm=instance_norm,
                         instance_norm_func=nn.InstanceNorm1d)


class Conv2d(_ConvBase):
    def __init__(self, in_size: int, out_size: int, *, kernel_size: Tuple[int, int] = (1, 1),
                 stride: Tuple[int, int] = (1, 1), padding: Tuple[int, int] = (0, 0), activation=nn.ReLU(inplace=True),
                 bn: bool = False, init=nn.init.kaiming_normal_, bias: bool = True, preact: bool = False,
                 name: str = "", instance_norm=False):
        super().__init__(in_size, out_size, kernel_size, stride, padding, activation, bn, init, conv=nn.Conv2d,
                         batch_norm=BatchNorm2d, bias=bias, preact=preact, name=name, instance_norm=instance_norm,
                         instance_norm_func=nn.InstanceNorm2d)


class FC(nn.Sequential):
    """Linear [+ bn] [+ activation] (pytorch_utils.py:200-236)."""

    def __init__(self, in_size: int, out_size: int, *, activation=nn.ReLU(inplace=True), bn: bool = False, init=None,
                 preact: bool = False, name: str = ""):
        super().__init__()
        fc = nn.Linear(in_size, out_size, bias=not bn)
        if init is not None:
            init(fc.weight)
        if not bn:
            nn.init.constant_(fc.bias, 0)
        tail = []
        if bn:
            tail.append((name + "bn", BatchNorm1d(in_size if preact else out_size)))
        if activation is not None:
            tail.append((name + "activation", activation))
        if preact:
            for k, mod in tail:
                self.add_module(k, mod)
        self.add_module(name + "fc", fc)
        if not preact:
            for k, mod in tail:
                self.add_module(k, mod)


class SharedMLP(nn.Sequential):
    """args = [c_in, c_1, ..., c_L]: L blocks named 'layer0'..'layer{L-1}' (pytorch_utils.py:5-32)."""

    def __init__(self, args: List[int], *, bn: bool = False, activation=nn.ReLU(inplace=True), preact: bool = False,
                 first: bool = False, name: str = "", instance_norm: bool = False):
        super().__init__()
        for i in range(len(args) - 1):
            plain = first and preact and i == 0     # the reference drops bn+activation only in this case
            self.add_module(
                name + "layer{}".format(i),
                Conv2d(args[i], args[i + 1], bn=(not plain) and bn, activation=None if plain else activation,
                       preact=preact, instance_norm=instance_norm))

    # Training-mode route: the layers are torch's (cuDNN 1x1 convolution, BatchNorm2d with batch statistics, ReLU).  On the grouped
    # (B, C, P, K) tensors of this model the NCHW BatchNorm kernels of cuDNN run one block per channel (16-128 blocks on 148 SMs):
    # 61 % of a fwd+bwd micro-batch.  In channels_last memory the same layers take the NHWC kernels: 48.8 -> 26.7 ms per 60-cloud
    # micro-batch (tools/train_profile.py).  Values and gradients are the same tensors, only the strides differ.
    channels_last_training = True

    def forward(self, x):
        if self.channels_last_training and self.training and x.dim() == 4 and x.is_cuda and x.dtype == torch.float32:
            x = x.contiguous(memory_format=torch.channels_last)
        return super().forward(x)


def fold_shared_mlp(mlp: SharedMLP):
    """Folds every conv(+eval-mode BN) block of a SharedMLP into (W [out,in] fp32, b [out] fp32).

    Returns None when the stack is not a plain conv->bn->ReLU / conv->ReLU chain (pre-activation order,
    instance norm, missing ReLU, non-1x1 kernels): the caller then keeps the layer-by-layer path.
    """
    folded = []
    for block in mlp:
        kids = dict(block.named_children())
        conv = kids.get("conv")
        if conv is None or list(block.named_children())[0][0] != "conv":
            return None
        if tuple(conv.kernel_size) != (1, 1) or tuple(conv.stride) != (1, 1) or tuple(conv.padding) != (0, 0):
            return None
        if "in" in kids or not isinstance(kids.get("activation"), nn.ReLU):
            return None
        w = conv.weight.detach().reshape(conv.out_channels, conv.in_channels).float()
        b = conv.bias.detach().float() if conv.bias is not None else torch.zeros(conv.out_channels, device=w.device)
        if "bn" in kids:
            bn = kids["bn"][0]
            if bn.running_mean is None or bn.running_var is None:
                return None
            scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
            w = w * scale[:, None]
            b = (b - bn.running_mean.detach().float()) * scale + bn.bias.detach().float()
        folded.append((w.contiguous(), b.contiguous()))
    return folded


def shared_mlp_version(mlp: SharedMLP):
    """Cheap fingerprint of everything fold_shared_mlp reads (tensor identities and in-place version counters)."""
    sig = []
    for t in list(mlp.parameters()) + list(mlp.buffers()):
        sig.append((t.data_ptr(), t._version, t.device))
    return tuple(sig)
