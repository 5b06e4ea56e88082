#!/usr/bin/env python
"""Stage-by-stage CUDA-event timing of the FP1 / FP2 modules' eval route at c3 sizes (L2 flushed before each stage)."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from garment4d_b200 import _lib, synthetic
from garment4d_b200.encoder import Pointnet2MSGSEG
from garment4d_b200.pointnet2 import pointnet2_utils as pu

C, N = 240, 8192
dev = torch.device("cuda:0")
L = _lib.lib()
torch.manual_seed(1234)
model = Pointnet2MSGSEG(input_channels=0, bn=True, global_feat=False).to(dev).eval()
pc = torch.from_numpy(np.tile(synthetic.body_clouds(4234, 16, N), ((C + 15) // 16, 1, 1))[:C].copy()).to(dev)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)


def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.fill_(0.0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / reps * 1e3


with torch.no_grad():
    lx, lf = model.sa_stack(pc)
    f2 = model.FP_modules[2](lx[2], lx[3], lf[2], lf[3])
    for name, fp, unknown, known, skip, kf in (("FP2", model.FP_modules[2], lx[2], lx[3], lf[2], lf[3]), ("FP1", model.FP_modules[1], lx[1], lx[2], lf[1], f2)):
        B, n, _ = unknown.shape
        m, c2, c1 = known.shape[1], kf.shape[1], skip.shape[1]
        folded = fp._folded_mlp(kf, skip)["half"]
        d2 = torch.empty(B, n, 3, device=dev); i3 = torch.empty(B, n, 3, dtype=torch.int32, device=dev)
        x = torch.empty(c2 + c1, B * n, dtype=torch.float16, device=dev)
        kpm = kf._g4d_pm
        sp = _lib.stream_ptr
        print(f"{name}: n={n} m={m} c2={c2} c1={c1} mlp={[w.shape[0] for w, _ in folded]}  whole module {t(lambda: fp(unknown, known, skip, kf)):.1f} us")
        print(f"   three_nn            {t(lambda: pu.three_nn_raw(unknown, known, d2, i3)):8.1f} us")
        print(f"   interp_concat (cm)  {t(lambda: L.g4d_fp_interp_concat_cbn_h(B, c2, c1, m, n, _lib.ptr(d2), _lib.ptr(i3), _lib.ptr(kf), _lib.ptr(skip), _lib.ptr(x), sp())):8.1f} us")
        print(f"   interp_concat (pm)  {t(lambda: L.g4d_fp_interp_concat_pm_cbn_h(B, c2, c1, m, n, _lib.ptr(d2), _lib.ptr(i3), _lib.ptr(kpm), _lib.ptr(skip), _lib.ptr(x), sp())):8.1f} us")
        (w1, b1), (w2, b2) = folded
        print(f"   mm1 {tuple(w1.shape)} x {tuple(x.shape)} {t(lambda: torch.mm(w1, x)):8.1f} us")
        y1 = torch.mm(w1, x)
        print(f"   bias_relu_h         {t(lambda: L.g4d_bias_relu_h(w1.shape[0], B * n, _lib.ptr(y1), _lib.ptr(b1), 1, sp())):8.1f} us")
        print(f"   mm2 -> fp32         {t(lambda: torch.mm(w2, y1, out_dtype=torch.float32)):8.1f} us")
        print(f"   mm2 -> fp16         {t(lambda: torch.mm(w2, y1)):8.1f} us")
        y2 = torch.mm(w2, y1, out_dtype=torch.float32)
        out = torch.empty(B, w2.shape[0], n, device=dev); pm = torch.empty(B, n, w2.shape[0], dtype=torch.float16, device=dev)
        print(f"   unpack (+pm)        {t(lambda: L.g4d_bias_relu_unpack(B, w2.shape[0], n, _lib.ptr(y2), 0, _lib.ptr(b2), 1, _lib.ptr(out), _lib.ptr(pm), sp())):8.1f} us")
        spm = skip._g4d_pm
        xr = torch.empty(B * n, c2 + c1, dtype=torch.float16, device=dev)
        print(f"   interp_concat rows  {t(lambda: L.g4d_fp_interp_concat_rows_h(B, c2, c1, m, n, _lib.ptr(d2), _lib.ptr(i3), _lib.ptr(kpm), _lib.ptr(spm), _lib.ptr(xr), sp())):8.1f} us")
        print(f"   linear1 rows        {t(lambda: torch.nn.functional.linear(xr, w1)):8.1f} us")
        yr1 = torch.nn.functional.linear(xr, w1)
        print(f"   bias_relu_rows_h    {t(lambda: L.g4d_bias_relu_rows_h(B * n, w1.shape[0], _lib.ptr(yr1), _lib.ptr(b1), 1, sp())):8.1f} us")
        print(f"   mm2 rows -> fp32    {t(lambda: torch.mm(yr1, w2.t(), out_dtype=torch.float32)):8.1f} us")
        yr2 = torch.mm(yr1, w2.t(), out_dtype=torch.float32)
        print(f"   rows_unpack (+pm)   {t(lambda: L.g4d_bias_relu_rows_unpack(B, w2.shape[0], n, _lib.ptr(yr2), _lib.ptr(b2), 1, _lib.ptr(out), _lib.ptr(pm), sp())):8.1f} us")
        print(f"   allocs (4 empty)    {t(lambda: [torch.empty(c2 + c1, B * n, dtype=torch.float16, device=dev), torch.empty(B, n, 3, device=dev)]):8.1f} us")
