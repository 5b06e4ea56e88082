#!/usr/bin/env python
"""Where the c4 fwd+bwd micro-batch (60 clouds x 8192 points, encoder in training mode, operator route) spends its GPU time:
torch.profiler kernel table.   python tools/train_profile.py [channels_last]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
import bench
from garment4d_b200.encoder import Pointnet2MSGSEG
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = Pointnet2MSGSEG(input_channels=0, bn=True, global_feat=False).to(dev).train()
if len(sys.argv) > 1 and sys.argv[1] == "channels_last":
    model = model.to(memory_format=torch.channels_last)
x = torch.from_numpy(bench.make_inputs("body", 5, 60, 8192)).to(dev)
y = torch.from_numpy(np.random.RandomState(1).randint(0, 7, (60 * 8192,))).to(dev)


def step():
    sem = model(x)[1]
    loss = F.cross_entropy(sem.reshape(-1, sem.shape[-1]), y)
    loss.backward()
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record(); step(); e.record(); torch.cuda.synchronize()
print(f"micro-batch fwd+bwd: {s.elapsed_time(e):.2f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
