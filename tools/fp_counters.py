#!/usr/bin/env python
"""Role-level cycle counters of g4d_fp_interp_mlp at c3 sizes (g4d_debug_fp_counters): who waits for whom."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from garment4d_b200 import _lib, synthetic
from garment4d_b200.encoder import Pointnet2MSGSEG
from garment4d_b200.pointnet2 import pointnet2_cuda_bridge as bridge

C, N = int(sys.argv[1]) if len(sys.argv) > 1 else 240, 8192
dev = torch.device("cuda:0")
L = _lib.lib()
torch.manual_seed(1234)
model = Pointnet2MSGSEG(input_channels=0, bn=True, global_feat=False).to(dev).eval()
pc = torch.from_numpy(np.tile(synthetic.body_clouds(4234, 16, N), ((C + 15) // 16, 1, 1))[:C].copy()).to(dev)
with torch.no_grad():
    lx, lf = model.sa_stack(pc)
    f2 = model.FP_modules[2](lx[2], lx[3], lf[2], lf[3])
    f1 = model.FP_modules[1](lx[1], lx[2], lf[1], f2)
    assert model._fused_fp0_head(lx, [None, f1, f2, lf[3]]) is not None
    packed = model._fp0_cache[str(dev)][1]
    buf = torch.zeros(22, dtype=torch.int64, device=dev)
    for _ in range(2):
        bridge.fp_interp_mlp(packed, lx[0], lx[1], f1)
    torch.cuda.synchronize()
    L.g4d_debug_fp_counters(ctypes.c_void_p(buf.data_ptr()))
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); bridge.fp_interp_mlp(packed, lx[0], lx[1], f1); e.record()
    torch.cuda.synchronize()
    L.g4d_debug_fp_counters(None)
c = buf.cpu().tolist()
print(f"NA={os.environ.get('G4D_FP_NA', 'default')}: three_nn + fp_interp_mlp {s.elapsed_time(e):.3f} ms; tiles of CTA 0: {c[5]}")
print(f"  consumer group 0: loop {c[4]} cycles, issuer waiting for a full A buffer {c[2]} ({100.0 * c[2] / max(c[4], 1):.1f} %), "
      f"waiting for layer-1 MMAs {c[3]} ({100.0 * c[3] / max(c[4], 1):.1f} %)")
tiles = max(c[5] // 2, 1)
names = ["L1 issue (incl. wait for A)", "L1 MMAs", "epilogue 1 + barrier", "L2 issue", "L2 MMAs", "epilogue 2 (+ 64 channel stores) + barrier",
         "L3 issue", "L3 MMAs", "epilogue 3 + barrier", "L4 issue", "L4 MMAs", "epilogue 4 (logits, labels) + barrier"]
print("  phases of consumer group 0, cycles per tile (its issuing lane):")
for n, v in zip(names, c[8:20]):
    print(f"      {n:44s} {v / tiles:8.0f}")
print(f"      {'sum':44s} {sum(c[8:20]) / tiles:8.0f}")
print(f"      inside epilogue 4: after tcgen05.ld {c[20] / tiles:.0f}, after the stores {c[21] / tiles:.0f}, after the group barrier {c[19] / tiles:.0f}")
