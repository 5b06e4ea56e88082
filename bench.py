#!/usr/bin/env python
"""Benchmark of the Garment4D hot path on B200: PointNet++ encoder forward (Pointnet2MSGSEG, eval mode) + SMPL lbs()
for the same frames.  Metric (BASELINE.json): frames/s, one frame = one cloud of N points.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c3] [--impl b200|reference]

One "step" = one pass of the hot path over one batch of C = B*T synthetic frames (default workload: BASELINE config
c3, B=8 x T=30 x N=8192, encoder + LBS with V=6890 -- the configuration the metric "encoder+LBS fwd" is quoted on;
per-GPU work is fixed as N grows: weak scaling, frames are independent so there is no data-path collective).
Prints ONE JSON line (see README / DESIGN.md for the keys).
"""
import argparse
import copy
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {  # name: (B, T, N)   BASELINE.json configs[1..4]
    "c2": (8, 1, 8192), "c3": (8, 30, 8192), "c4": (32, 30, 8192), "c5": (64, 30, 16384),
}
V_SMPL = 6890
# algorithmic work per cloud of the SA stack (independent of N except the FPS/ball terms)
SA_BRANCHES = [  # (level, n_in or None=N, npoint, K, c_in, (c1,c2,c3))
    (0, None, 1024, 16, 0, (16, 16, 32)), (0, None, 1024, 32, 0, (32, 32, 64)),
    (1, 1024, 256, 16, 96, (32, 32, 64)), (1, 1024, 256, 32, 96, (64, 64, 128)),
    (2, 256, 64, 32, 192, (64, 64, 128)), (2, 256, 64, 64, 192, (128, 128, 256)),
]


def branch_macs(K, c_in, mlp, npoint):
    c1, c2, c3 = mlp
    return npoint * K * ((c_in + 3) * c1 + c1 * c2 + c2 * c3)


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tc_tflops": d["bf16_tflops"], "tc_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tc_tflops": 1590.0, "tc_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            busy = [s for s in sm if s >= 0.5 * max(sm)] or sm
            out.update(sm_mhz=float(np.median(busy)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU port of the path (oracle/), bounded sample

def cpu_path_frames_per_s(n_points, sample_clouds, steps, warmup, seed=1234):
    """Times oracle.encoder_cpu (C oracle ops + torch-CPU conv stacks) + oracle.lbs on `sample_clouds` frames per step."""
    import torch
    from garment4d_b200.encoder import Pointnet2MSGSEG
    from garment4d_b200 import synthetic
    from oracle import lbs as olbs
    from oracle import pointnet2 as orc
    from oracle.encoder_cpu import encoder_forward_cpu
    torch.manual_seed(seed)
    model = Pointnet2MSGSEG(input_channels=0, bn=True, global_feat=False).eval()
    pc = synthetic.body_clouds(seed, sample_clouds, n_points)
    smpl = synthetic.synthetic_smpl(seed=seed)
    betas, pose = synthetic.synthetic_frames(sample_clouds, seed=seed + 1)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        encoder_forward_cpu(model, pc)
        olbs.lbs(betas, pose, **smpl)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    t = float(np.mean(times))
    return sample_clouds / t, t, max(orc.num_threads(), torch.get_num_threads())


def run_reference(args, rank, world):
    if rank != 0:
        return
    B, T, N = CONFIGS[args.config]
    sample = args.ref_clouds
    fps, t, cores = cpu_path_frames_per_s(N, sample, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "frames/sec (B*T*N pts) encoder+LBS fwd", "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config}: B={B} T={T} N={N}, Pointnet2MSGSEG fwd (eval) + SMPL lbs V={V_SMPL}",
                   "sample": f"{sample} frames per step"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} frames of the same workload per step (the reference has no CPU pointnet2 ops; "
                                   "oracle/ is the CPU restatement of its CUDA kernels, lbs = numpy restatement of lbs.py)"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--ref-clouds", type=int, default=8, help="frames per step of the CPU reference arm / cpu_baseline")
    ap.add_argument("--chunks", type=int, default=4, help="frame groups per step, each on its own CUDA stream (EncoderLBSRunner)")
    ap.add_argument("--no-graph", action="store_true", help="launch kernel by kernel instead of replaying the captured CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-breakdown", action="store_true")
    ap.add_argument("--seed-rank", type=int, default=None, help="debug: generate the synthetic inputs of this rank (single-GPU repro of a multi-GPU run)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from garment4d_b200 import _lib, synthetic
    from garment4d_b200 import lbs as glbs
    from garment4d_b200.encoder import Pointnet2MSGSEG
    from garment4d_b200.pointnet2 import pointnet2_utils as pu

    assert torch.cuda.is_available(), "bench.py --impl b200 needs a GPU"
    L = _lib.lib()      # raises if the CUDA extension is missing: no fallback
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B, T, N = CONFIGS[args.config]
    C = B * T
    seed = 1234 + 1000 * int(args.config[1]) + (rank if args.seed_rank is None else args.seed_rank)
    torch.manual_seed(1234)                     # same weights on every rank
    model = Pointnet2MSGSEG(input_channels=0, bn=True, global_feat=False).to(dev).eval()
    smpl_np = synthetic.synthetic_smpl(seed=1234)
    smpl = [torch.from_numpy(np.ascontiguousarray(smpl_np[k])).to(dev) for k in
            ("v_template", "shapedirs", "posedirs", "J_regressor", "parents", "lbs_weights")]
    # a few distinct clouds tiled to C frames (generation cost), each frame jittered so no two are identical
    base = synthetic.body_clouds(seed, min(C, 16), N)
    reps = (C + base.shape[0] - 1) // base.shape[0]
    pc_host = np.tile(base, (reps, 1, 1))[:C].copy()
    pc_host += (np.random.RandomState(seed).randn(C, 1, 3) * 0.01).astype(np.float32)
    betas_np, pose_np = synthetic.synthetic_frames(C, seed=seed + 1)
    pc_pin = torch.from_numpy(pc_host).pin_memory()
    betas_pin, pose_pin = torch.from_numpy(betas_np).pin_memory(), torch.from_numpy(pose_np).pin_memory()
    pc_dev, betas_dev, pose_dev = pc_pin.to(dev), betas_pin.to(dev), pose_pin.to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)     # 256 MB > 126 MB L2

    from garment4d_b200.runner import EncoderLBSRunner, GraphedEncoderLBSRunner
    if args.no_graph:
        runner = EncoderLBSRunner(model, smpl, chunks=args.chunks, device=dev)
    else:
        runner = GraphedEncoderLBSRunner(model, smpl, chunks=args.chunks, device=dev)

    def step(pc, betas, pose):
        if args.no_graph:
            return runner.forward_device(pc, betas, pose)
        return runner.replay_device()          # static input buffers hold pc_dev / betas_dev / pose_dev (resident in HBM)

    lab_pin = torch.empty(C, N, dtype=torch.uint8).pin_memory()
    verts_pin = torch.empty(C, V_SMPL, 3, dtype=torch.float32).pin_memory()
    joints_pin = torch.empty(C, 24, 3, dtype=torch.float32).pin_memory()

    def step_e2e():
        if args.no_graph:
            runner.forward_host(pc_pin, betas_pin, pose_pin, lab_pin, verts_pin, joints_pin)
        else:
            runner.replay_host()

    if not args.no_graph:
        runner.capture(pc_dev, betas_dev, pose_dev)
        runner.capture_host(pc_pin, betas_pin, pose_pin, lab_pin, verts_pin, joints_pin)

    h2d = pc_pin.numel() * 4 + betas_pin.numel() * 4 + pose_pin.numel() * 4
    d2h = lab_pin.numel() + verts_pin.numel() * 4 + joints_pin.numel() * 4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        n0 = L.g4d_launch_count()
        sampler = ClockSampler(local_rank) if rank == 0 else None
        for s, e in ev:
            flush.fill_(0.0)                      # evict L2 between timed iterations (not timed)
            s.record()
            fn()
            e.record()
        barrier()
        clocks = sampler.stop() if sampler else None
        launches = (L.g4d_launch_count() - n0) // steps
        total_ms = sum(s.elapsed_time(e) for s, e in ev)
        if world > 1:
            t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        return total_ms / steps, launches, clocks

    ms, launches, clocks = timed(lambda: step(pc_dev, betas_dev, pose_dev), args.steps, args.warmup)
    if not args.no_graph:
        launches = runner.kernels_per_replay      # kernels of libgarment4d_b200.so replayed by the graph each step
    ms_e2e, _, _ = timed(step_e2e, args.steps, args.warmup)

    peaks = read_peaks()
    kernels, roof = [], None
    if not args.no_kernel_breakdown:
        kernels = kernel_breakdown(torch, L, model, pc_dev, betas_dev, pose_dev, smpl, flush, peaks, C, N)
        if rank == 0 and kernels:
            top = max((k for k in kernels if k.get("roofline")), key=lambda k: k["ms"])
            roof = dict(top["roofline"], kernel=top["name"], ms_per_launch=top["ms"], peak_source=peaks["source"], note=top.get("note", ""))

    cpu = None
    if rank == 0 and args.gpus == 1 and not args.no_cpu_baseline:
        # bounded sample: ~10 s of CPU work (6 timed passes + 1 warm-up over 6 x ref_clouds frames of the same workload)
        n_cpu, passes = 6 * args.ref_clouds, 6
        fps_cpu, t_cpu, cores = cpu_path_frames_per_s(N, n_cpu, passes, 1)
        cpu = {"value": fps_cpu, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": f"{n_cpu} frames of the same workload per pass, {passes} timed passes (oracle/: C restatement of the reference CUDA "
                         f"kernels with OpenMP + torch-CPU conv stacks + numpy lbs), {t_cpu:.2f} s per pass"}

    if rank == 0:
        frames = C * world
        line = {
            "metric": "frames/sec (B*T*N pts) encoder+LBS fwd", "value": frames / (ms * 1e-3), "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16 operands / f32 accumulate (grouped MLP); f32 (FPS, ball query, LBS)",
            "data": "synthetic",
            "config": {"workload": f"{args.config}: B={B} T={T} N={N} per GPU, Pointnet2MSGSEG fwd (eval, 3 SA + 3 FP + seg head) + SMPL lbs V={V_SMPL}",
                       "frames_per_step_per_gpu": C, "l2": "flushed between timed iterations (256 MB fill)",
                       "streams": f"{args.chunks} frame groups per step on separate CUDA streams (copy/compute overlap)",
                       "launch": "kernel by kernel" if args.no_graph else "one captured CUDA graph per step",
                       "parallelism": f"dp{world} (frames sharded, no data-path collective)"},
            "e2e": {"value": frames / (ms_e2e * 1e-3), "unit": "frames/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "result": "uint8 segmentation labels + posed vertices + joints"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "kernels": kernels,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def ncu_traffic(ncu_rows, key):
    """DRAM bytes of one launch from the rows of profiles/*_ncu_summary.json.  key = (kernel-name substring, occurrence) or a
    list of such pairs whose launches together make up the operator (summed); None / not captured -> None."""
    if key is None or not ncu_rows:
        return None
    keys = key if isinstance(key, list) else [key]
    total = 0.0
    for sub, occ in keys:
        hits = [r for r in ncu_rows if sub in r["kernel"]]
        if occ >= len(hits) or hits[occ].get("traffic_bytes") is None:
            return None
        total += hits[occ]["traffic_bytes"]
    return total


def kernel_breakdown(torch, L, model, pc, betas, pose, smpl, flush, peaks, C, N, reps=5):
    """Per-kernel device time (CUDA events on the launching stream, L2 flushed before each launch) and the
    algorithmic-work roofline of each (bytes / flops per cloud from SURVEY.md section 8(d))."""
    from garment4d_b200 import lbs as glbs
    from garment4d_b200.pointnet2 import pointnet2_utils as pu
    import ctypes
    from garment4d_b200 import _lib
    out = []

    def t(fn):
        fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(reps):
            flush.fill_(0.0)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record()
            torch.cuda.synchronize()
            tot += s.elapsed_time(e)
        return tot / reps

    # DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the same c3-sized launch, from the committed
    # `ncu --set full` capture (profiles/*_ncu_summary.json, written by tools/summarize_ncu.py); None when not captured.
    ncu_rows = []
    if C == 240 and N == 8192:
        import glob
        cands = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_summary.json")))
        if cands:
            ncu_rows = json.load(open(cands[-1]))

    def traffic(key):
        return ncu_traffic(ncu_rows, key)

    def hbm(name, ms, bytes_per_cloud, note="", ncu=None):
        ach = bytes_per_cloud * C / (ms * 1e-3) / 1e9
        out.append({"name": name, "ms": ms, "roofline": {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                                         "frac": ach / peaks["hbm_gbs"], "traffic": traffic(ncu)}, "note": note})

    def tensor(name, ms, macs_per_cloud, ncu=None):
        ach = 2.0 * macs_per_cloud * C / (ms * 1e-3) / 1e12
        out.append({"name": name, "ms": ms, "roofline": {"bound": "tensor", "achieved": ach, "peak": peaks["tc_tflops"], "unit": "TFLOP/s",
                                                         "frac": ach / peaks["tc_tflops"], "traffic": traffic(ncu)}})

    with torch.no_grad():
        xyz, feats = pc.contiguous(), None
        lx, lf = [xyz], [None]
        for lvl, sa in enumerate(model.SA_modules):
            n_in, P = xyz.shape[1], sa.npoint
            ms = t(lambda: pu.furthest_point_sample_and_gather(xyz, P))
            hbm(f"fps_gather L{lvl} ({n_in}->{P})", ms, 12 * n_in + 16 * P,
                note=f"serial ALU chain: {10 * n_in * (P - 1) * C / (ms * 1e-3) / 1e12:.2f} T lane-ops/s of 37.2 peak",
                ncu=(("fps_pruned_kernel<512", 0), ("fps_kernel<256, 4", 0), ("fps_kernel<256, 1", 0))[lvl] if lvl < 3 else None)
            _, new_xyz = pu.furthest_point_sample_and_gather(xyz, P)
            g0, g1 = sa.groupers
            ms = t(lambda: pu.ball_query_pair(g0.radius, g0.nsample, g1.radius, g1.nsample, xyz, new_xyz))
            hbm(f"ball_query2 L{lvl}", ms, 12 * n_in + 12 * P + 4 * P * (g0.nsample + g1.nsample),
                ncu=("ball_query_grid_kernel<2>", 0) if lvl == 0 else ("ball_query_kernel<2>", lvl - 1))
            idxs = pu.ball_query_pair(g0.radius, g0.nsample, g1.radius, g1.nsample, xyz, new_xyz)
            c_in = 0 if feats is None else feats.shape[1]
            # the north star's fused ball-query+group operator (materialises the grouped tensor; not on the fused route)
            for gi, g in enumerate((g0, g1)):
                ms = t(lambda: pu.QueryAndGroup(g.radius, g.nsample)(xyz, new_xyz, feats))
                # the operator = one ball query (grid kernel at level 0, brute force below) + the fused grouping pass
                bq = ("ball_query_grid_kernel<1>", gi) if lvl == 0 else ("ball_query_kernel<1>", 2 * (lvl - 1) + gi)
                hbm(f"query_and_group L{lvl} K={g.nsample}", ms,
                    12 * n_in + 12 * P + 4 * c_in * n_in + 4 * P * g.nsample + 4 * (c_in + 3) * P * g.nsample,
                    ncu=[bq, ("group_fused_kernel", 2 * lvl + gi)])
            new_xyz2, new_feats = sa(xyz, feats)
            feat_pm = None if feats is None else getattr(feats, "_g4d_pm")
            ctot = new_feats.shape[1]
            out_cm = torch.empty_like(new_feats)
            out_pm = torch.empty(C, P, ctot, dtype=torch.float16, device=pc.device)
            off = 0
            for i, (g, idx) in enumerate(zip((g0, g1), idxs)):
                br = sa._branch(i, c_in, xyz.device)
                def run(br=br, idx=idx, off=off):
                    rc = L.g4d_sa_mlp_max(ctypes.byref(br.desc), _lib.ptr(br.params), C, n_in, P, _lib.ptr(xyz), _lib.ptr(new_xyz),
                                          _lib.ptr(idx), _lib.ptr(feat_pm), _lib.ptr(out_cm), _lib.ptr(out_pm), ctot, off, _lib.stream_ptr())
                    _lib.check(rc, "g4d_sa_mlp_max")
                ms = t(run)
                d = br.desc
                tensor(f"sa_mlp_max L{lvl} K={g.nsample} ({c_in}+3->{d.c1},{d.c2},{d.c3})", ms,
                       branch_macs(g.nsample, c_in, (d.c1, d.c2, d.c3), P),
                       ncu=(f"sa_mlp_max_kernel<{g.nsample}, {1 if c_in else 0}, ", 1 if (lvl == 2 and g.nsample == 32) else 0))
                off += br.c_out
            xyz, feats = new_xyz2, new_feats
            lx.append(xyz); lf.append(feats)
        # feature propagation: FP2 / FP1 modules (our prologue + library GEMMs + our epilogues), then the fused finest level + head
        from garment4d_b200.pointnet2 import pointnet2_cuda_bridge as bridge
        fp2, fp1 = model.FP_modules[2], model.FP_modules[1]
        ms = t(lambda: fp2(lx[2], lx[3], lf[2], lf[3]))
        f2 = fp2(lx[2], lx[3], lf[2], lf[3])
        out.append({"name": "FP2 module (three_nn + interp/concat kernel + 2 library GEMMs + bias/ReLU kernels)", "ms": ms})
        ms = t(lambda: fp1(lx[1], lx[2], lf[1], f2))
        f1 = fp1(lx[1], lx[2], lf[1], f2)
        out.append({"name": "FP1 module (three_nn + interp/concat kernel + 2 library GEMMs + bias/ReLU(+fp16 point-major) kernels)", "ms": ms})
        d2 = torch.empty(C, N, 3, dtype=torch.float32, device=pc.device)
        i3 = torch.empty(C, N, 3, dtype=torch.int32, device=pc.device)
        ms_nn = t(lambda: pu.three_nn_raw(lx[0], lx[1], d2, i3))
        hbm(f"three_nn L0 ({N} -> {lx[1].shape[1]}, grid)", ms_nn, 12 * N + 12 * lx[1].shape[1] + 24 * N, ncu=("three_nn_grid_kernel", 0))
        if model._fused_fp0_head(lx, [None, f1, f2, lf[3]]) is not None:
            packed = model._fp0_cache[str(pc.device)][1]
            ms = t(lambda: bridge.fp_interp_mlp(packed, lx[0], lx[1], f1)) - ms_nn
            d = packed.desc
            hbm(f"fp_interp_mlp ({d.c_in}->{d.c1},{d.c2} + head {d.h1},{d.h2}; tcgen05)", ms,
                4 * d.c2 * N + 4 * d.h2 * N + 24 * N + 2 * d.c_in * lx[1].shape[1],
                note=f"{2.0 * N * (d.c_in * d.c1 + d.c1 * d.c2 + d.c2 * d.h1 + d.h1 * 16) * C / (ms * 1e-3) / 1e12:.1f} TFLOP/s on the tensor pipe",
                ncu=("fp_interp_mlp_kernel", 0))
        ms = t(lambda: glbs.lbs(betas, pose, *smpl))
        hbm("lbs (6 kernels)", ms, 83872, note=f"{2 * 7737470 * C / (ms * 1e-3) / 1e12:.2f} TFLOP/s fp32 of ~74.5 FFMA peak")
        ms_sa = t(lambda: model.sa_stack(pc))
        ms_all = t(lambda: model(pc))
        out.append({"name": "SA stack (3 levels, all kernels)", "ms": ms_sa})
        out.append({"name": "FP stack + seg head (torch/cuDNN + three_nn/three_interpolate)", "ms": ms_all - ms_sa})
    return out


if __name__ == "__main__":
    main()
