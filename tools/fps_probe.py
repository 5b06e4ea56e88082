import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from garment4d_b200 import synthetic
from garment4d_b200.pointnet2 import pointnet2_utils as pu
dev = torch.device("cuda:0")
def timeit(x, m, label):
    pu.build_grid(x, 0.1)
    for _ in range(2): pu.furthest_point_sample_and_gather(x, m)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5): pu.furthest_point_sample_and_gather(x, m)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    print(f"{label:40s} {ms:8.3f} ms  -> {ms*1e6/(m-1):8.1f} ns/iter")
C = int(sys.argv[1]) if len(sys.argv) > 1 else 240
body = torch.from_numpy(np.tile(synthetic.body_clouds(1, 16, 8192), (C // 16 + 1, 1, 1))[:C].copy()).to(dev)
timeit(body, 1024, f"body C={C}")
timeit(body[:148].contiguous(), 1024, "body C=148")
timeit(body[:1].contiguous(), 1024, "body C=1")
co = torch.full((C, 8192, 3), 0.25, device=dev)
timeit(co, 1024, f"coincident (no updates) C={C}")
timeit(co[:1].contiguous(), 1024, "coincident C=1")
cube = torch.rand(C, 8192, 3, device=dev)
timeit(cube, 1024, f"cube C={C}")
