#!/usr/bin/env python
"""Benchmark of the Garment4D hot path on B200: PointNet++ encoder forward (Pointnet2MSGSEG, eval mode) + SMPL lbs()
for the same frames.  Metric (BASELINE.json): frames/s, one frame = one cloud of N points.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c4] [--scaling strong|weak] [--impl b200|reference]

One "step" = one pass of the hot path over one batch of C = B*T synthetic frames.  Default workload: BASELINE config c4
(B=32 x T=30 x N=8192 = 960 frames per step, encoder + LBS with V=6890), the configuration the multi-GPU metric is quoted
on; it fits one GPU, so N=1 runs the same global batch.  STRONG scaling: the global batch is fixed
and its sequences are dealt out to the ranks like the reference's DistributedSampler (garment4d_b200/sharding.py); frames
are independent, so the forward pass has no data-path collective.  The `train` object of the line is the fwd+bwd step of
config c4 (encoder in training mode, segmentation loss, ONE flat NCCL all-reduce of the gradients, Adam step).
Prints ONE JSON line (see README / DESIGN.md for the keys).
"""
import argparse
import datetime
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
    "c4_8": (4, 30, 8192),      # (not a BASELINE config: one rank's share of c4 at 8 GPUs, for single-GPU experiments)
}
V_SMPL = 6890
METRIC = "frames/sec (B*T*N pts) encoder+LBS fwd"
# algorithmic work per cloud of the SA stack (independent of N except the FPS/ball terms)
SA_BRANCHES = [  # (level, n_in or None=N, npoint, K, c_in, (c1,c2,c3))
    (0, None, 1024, 16, 0, (16, 16, 32)), (0, None, 1024, 32, 0, (32, 32, 64)),
    (1, 1024, 256, 16, 96, (32, 32, 64)), (1, 1024, 256, 32, 96, (64, 64, 128)),
    (2, 256, 64, 32, 192, (64, 64, 128)), (2, 256, 64, 64, 192, (128, 128, 256)),
]
# FPS: measured floor of one serial step (block-wide arg-max + barrier with no distance update), ns
FPS_STEP_FLOOR_NS = 300.0


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


def progress(rank, msg):
    """Per-rank progress line on stderr: tells a hung kernel from a hung collective in a multi-GPU log."""
    print(f"[bench rank {rank} t={time.time() % 10000:8.2f}] {msg}", file=sys.stderr, flush=True)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU port of the path (oracle/), bounded sample

def cpu_path_frames_per_s(n_points, sample_clouds, steps, warmup, threads, seed=1234):
    """Times oracle.encoder_cpu (C oracle ops, OpenMP over clouds, + torch-CPU conv stacks) + oracle.lbs on `sample_clouds`
    frames per step with `threads` host threads (set explicitly: torchrun exports OMP_NUM_THREADS=1)."""
    os.environ["OMP_NUM_THREADS"] = str(threads)          # read by the OpenMP runtime of liboracle.so when it is loaded (below)
    import torch
    torch.set_num_threads(threads)
    from garment4d_b200.encoder import Pointnet2MSGSEG
    from garment4d_b200 import synthetic
    from oracle import lbs as olbs
    from oracle import pointnet2 as orc
    from oracle.encoder_cpu import encoder_forward_cpu
    orc.set_num_threads(threads)
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
    return sample_clouds / t, t, min(threads, max(orc.num_threads(), torch.get_num_threads()))


def run_reference(args, rank, world):
    if rank != 0:
        return
    B, T, N = CONFIGS[args.config]
    cores = host_cores()
    sample = args.ref_clouds if args.ref_clouds > 0 else max(8, cores)       # >= one cloud per host thread (OpenMP over clouds)
    fps, t, used = cpu_path_frames_per_s(N, sample, args.steps, args.warmup, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config}: B={B} T={T} N={N}, Pointnet2MSGSEG fwd (eval) + SMPL lbs V={V_SMPL}",
                   "sample": f"{sample} frames per step"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": used, "kind": "port",
                         "sample": f"{sample} frames of the same workload per step on {used} of {cores} host threads (the reference has no CPU "
                                   "pointnet2 ops; oracle/ is the CPU restatement of its CUDA kernels, lbs = numpy restatement of lbs.py)"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------

def make_inputs(kind, seed, C, N):
    """C synthetic clouds: a few distinct ones tiled to C frames (generation cost), each frame jittered so no two are equal."""
    from garment4d_b200 import synthetic
    nbase = min(C, 16)
    if kind == "body":
        base = synthetic.body_clouds(seed, nbase, N)
    else:
        base = np.random.RandomState(seed).rand(nbase, N, 3).astype(np.float32)       # SURVEY 8(d) cloud A: U[0,1)^3
    reps = (C + nbase - 1) // nbase
    pc = np.tile(base, (reps, 1, 1))[:C].copy()
    pc += (np.random.RandomState(seed + 7).randn(C, 1, 3) * 0.01).astype(np.float32)
    return pc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c4", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: the config's global batch is sharded over the ranks; weak: every rank runs the whole config")
    ap.add_argument("--ref-clouds", type=int, default=0, help="frames per step of the CPU reference arm (0 = one per host thread, >= 8)")
    ap.add_argument("--chunks", type=int, default=4, help="frame groups per step, each on its own CUDA stream (EncoderLBSRunner)")
    ap.add_argument("--no-graph", action="store_true", help="launch kernel by kernel instead of replaying the captured CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-breakdown", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the fwd+bwd (config c4) arm")
    ap.add_argument("--no-extras", action="store_true", help="skip the cube-cloud run and the label-agreement check")
    ap.add_argument("--train-steps", type=int, default=3)
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
    from garment4d_b200.encoder import Pointnet2MSGSEG
    from garment4d_b200.sharding import shard_sequences

    assert torch.cuda.is_available(), "bench.py --impl b200 needs a GPU"
    L = _lib.lib()      # raises if the CUDA extension is missing: no fallback
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ctl = None          # control-plane group (barriers, max over ranks): gloo, so that timing never waits inside a GPU collective
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        progress(rank, "init_process_group(nccl)")
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=240))
        ctl = dist.new_group(backend="gloo", timeout=datetime.timedelta(seconds=600))
        progress(rank, "process groups up")

    B, T, N = CONFIGS[args.config]
    seqs = list(range(B)) if (world == 1 or args.scaling == "weak") else shard_sequences(B, rank, world)
    C = len(seqs) * T                       # frames this rank processes per step
    C_global = B * T * (world if args.scaling == "weak" else 1)
    srank = rank if args.seed_rank is None else args.seed_rank
    seed = 1234 + 1000 * int(args.config[1]) + srank
    torch.manual_seed(1234)                     # same weights on every rank
    model = Pointnet2MSGSEG(input_channels=0, bn=True, global_feat=False).to(dev).eval()
    smpl_np = synthetic.synthetic_smpl(seed=1234)
    smpl = [torch.from_numpy(np.ascontiguousarray(smpl_np[k])).to(dev) for k in
            ("v_template", "shapedirs", "posedirs", "J_regressor", "parents", "lbs_weights")]
    pc_host = make_inputs("body", seed, C, N)
    betas_np, pose_np = synthetic.synthetic_frames(C, seed=seed + 1)
    pc_pin = torch.from_numpy(pc_host).pin_memory()
    betas_pin, pose_pin = torch.from_numpy(betas_np).pin_memory(), torch.from_numpy(pose_np).pin_memory()
    pc_dev, betas_dev, pose_dev = pc_pin.to(dev), betas_pin.to(dev), pose_pin.to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)     # 256 MB > 126 MB L2

    from garment4d_b200.runner import EncoderLBSRunner, GraphedEncoderLBSRunner

    def make_runner():
        cls = EncoderLBSRunner if args.no_graph else GraphedEncoderLBSRunner
        return cls(model, smpl, chunks=args.chunks, device=dev)

    runner = make_runner()

    lab_pin = torch.empty(C, N, dtype=torch.uint8).pin_memory()
    verts_pin = torch.empty(C, V_SMPL, 3, dtype=torch.float32).pin_memory()
    joints_pin = torch.empty(C, 24, 3, dtype=torch.float32).pin_memory()

    def step_dev():
        if args.no_graph:
            return runner.forward_device(pc_dev, betas_dev, pose_dev)
        return runner.replay_device()          # static input buffers hold pc_dev / betas_dev / pose_dev (resident in HBM)

    def step_e2e():
        if args.no_graph:
            runner.forward_host(pc_pin, betas_pin, pose_pin, lab_pin, verts_pin, joints_pin)
        else:
            runner.replay_host()

    if not args.no_graph:
        progress(rank, "capturing CUDA graphs")
        runner.capture(pc_dev, betas_dev, pose_dev)
        runner.capture_host(pc_pin, betas_pin, pose_pin, lab_pin, verts_pin, joints_pin)

    h2d = pc_pin.numel() * 4 + betas_pin.numel() * 4 + pose_pin.numel() * 4
    d2h = lab_pin.numel() + verts_pin.numel() * 4 + joints_pin.numel() * 4

    nbar = [0]

    def barrier(tag=""):
        """device idle -> all ranks here -> device idle.  The rendezvous is a gloo (host) barrier: no GPU kernel spins on a
        peer, so a rank that is late (or dead) shows up as a host-side timeout with the rank named, not as 8 busy GPUs."""
        torch.cuda.synchronize()
        if world > 1:
            nbar[0] += 1
            progress(rank, f"reached barrier {nbar[0]} {tag}")
            dist.barrier(group=ctl)
        torch.cuda.synchronize()

    def max_ranks(v):
        if world == 1:
            return float(v)
        t = torch.tensor([v], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=ctl)
        return float(t.item())

    def timed(fn, steps, warmup, tag, sample_clocks=False):
        for _ in range(warmup):
            fn()
        barrier(f"{tag}: warm-up done")
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        n0 = L.g4d_launch_count()
        sampler = ClockSampler(local_rank) if (rank == 0 and sample_clocks) else None
        for s, e in ev:
            flush.fill_(0.0)                      # evict L2 between timed iterations (not timed)
            s.record()
            fn()
            e.record()
        barrier(f"{tag}: timed steps done")
        clocks = sampler.stop() if sampler else None
        launches = (L.g4d_launch_count() - n0) // steps
        total_ms = max_ranks(sum(s.elapsed_time(e) for s, e in ev))        # device time, max over ranks
        return total_ms / steps, launches, clocks

    ms, launches, clocks = timed(step_dev, args.steps, args.warmup, "device-resident", sample_clocks=True)
    if not args.no_graph:
        launches = runner.kernels_per_replay      # kernels of libgarment4d_b200.so replayed by the graph each step
    ms_e2e, _, _ = timed(step_e2e, args.steps, args.warmup, "end-to-end")

    peaks = read_peaks()
    extras = {}
    if not args.no_extras:
        extras = run_extras(torch, args, model, smpl, dev, seed, C, C_global, N, make_runner, timed, rank)

    kernels, roof = [], None
    if not args.no_kernel_breakdown and rank == 0:
        progress(rank, "kernel breakdown")
        Ck = min(C, 240)                       # per-kernel times at the c3 launch size (the committed ncu captures' size)
        kernels = kernel_breakdown(torch, L, model, pc_dev[:Ck].contiguous(), betas_dev[:Ck].contiguous(), pose_dev[:Ck].contiguous(),
                                   smpl, flush, peaks, Ck, N)
        tops = [k for k in kernels if k.get("roofline")]
        if tops:
            top = max(tops, key=lambda k: k["ms"])
            roof = dict(top["roofline"], kernel=top["name"], ms_per_launch=top["ms"], clouds_per_launch=Ck, peak_source=peaks["source"],
                        note=top.get("note", ""))

    train = None
    if not args.no_train:
        train = train_arm(torch, dist, args, dev, rank, world, seqs, T, N, seed, barrier, max_ranks)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # bounded sample: ~10-20 s of CPU work on all host threads, at least one cloud per thread (OpenMP runs over clouds)
        cores = host_cores()
        n_cpu, passes = max(48, 2 * cores), 5
        fps_cpu, t_cpu, used = cpu_path_frames_per_s(N, n_cpu, passes, 1, cores)
        cpu = {"value": fps_cpu, "unit": "frames/s", "cores": used, "kind": "port",
               "sample": f"{n_cpu} frames of the same workload per pass, {passes} timed passes on {used} of {cores} host threads (oracle/: C "
                         f"restatement of the reference CUDA kernels with OpenMP + torch-CPU conv stacks + numpy lbs), {t_cpu:.2f} s per pass"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": C_global / (ms * 1e-3), "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f16 operands / f32 accumulate (grouped MLP, FP MLPs); f32 (FPS, ball query, three_nn, LBS)",
            "data": "synthetic",
            "config": {"workload": f"{args.config}: B={B} T={T} N={N} global batch ({C_global} frames per step), Pointnet2MSGSEG fwd (eval, 3 SA + 3 FP "
                                   f"+ seg head) + SMPL lbs V={V_SMPL}",
                       "frames_per_step_global": C_global, "frames_per_step_per_gpu": C,
                       "clouds": "body (2-D surface of capsules, 5 % duplicate points); cube clouds under `cube`",
                       "l2": "flushed between timed iterations (256 MB fill)",
                       "streams": f"{args.chunks} frame groups per step on separate CUDA streams (copy/compute overlap)",
                       "launch": "kernel by kernel" if args.no_graph else "one captured CUDA graph per step",
                       "parallelism": f"dp{world}: sequences dealt out like the reference's DistributedSampler, no data-path collective in forward; "
                                      "barriers / max over ranks on a gloo control group"},
            "e2e": {"value": C_global / (ms_e2e * 1e-3), "unit": "frames/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "result": "uint8 segmentation labels + posed vertices + joints"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "kernels": kernels,
            "cpu_baseline": cpu,
            "train": train,
        }
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        barrier("exit")
        dist.destroy_process_group()


def run_extras(torch, args, model, smpl, dev, seed, C, C_global, N, make_runner, timed, rank):
    """(1) the same step on 'cube' clouds (volumetric U[0,1)^3: the worst case of the pruned FPS and of the grid searches);
    (2) label agreement: fraction of arg-max segmentation labels of the fused route (fp16 operands) equal to the fp32
    module-by-module route (torch conv/BN/ReLU stacks, TF32 off) on the same clouds."""
    from garment4d_b200 import synthetic
    out = {}
    progress(rank, "extras: cube clouds")
    pc_cube = torch.from_numpy(make_inputs("cube", seed, C, N)).to(dev)
    b_np, p_np = synthetic.synthetic_frames(C, seed=seed + 1)
    betas, pose = torch.from_numpy(b_np).to(dev), torch.from_numpy(p_np).to(dev)
    r2 = make_runner()
    if not args.no_graph:
        r2.capture(pc_cube, betas, pose)
        fn = r2.replay_device
    else:
        fn = lambda: r2.forward_device(pc_cube, betas, pose)
    steps = max(3, min(args.steps, 10))
    ms_cube, _, _ = timed(fn, steps, 3, "cube clouds")
    out["cube"] = {"value": C_global / (ms_cube * 1e-3), "unit": "frames/s", "ms_per_step": ms_cube, "steps": steps,
                   "clouds": "U[0,1)^3 (SURVEY 8(d) cloud A)"}
    del r2
    if rank == 0:
        progress(rank, "extras: label agreement")
        import copy
        n_chk = min(C, 16)
        pcs = torch.from_numpy(make_inputs("body", seed, n_chk, N)).to(dev)
        ref = copy.deepcopy(model).eval()
        ref.fused = False
        for sa in ref.SA_modules:
            sa.fused = False
        for fp in ref.FP_modules:
            fp.fused = False
        tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            with torch.no_grad():
                sem_fused = model(pcs)[1]
                sem_ref = ref(pcs)[1]      # module by module in fp32: the reference's operator sequence, BN not folded, no fp16 anywhere
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
        agree = float((sem_fused.argmax(2) == sem_ref.argmax(2)).float().mean().item())
        out["label_agreement"] = {"value": agree, "frames": n_chk, "points": n_chk * N,
                                  "max_logit_err": float((sem_fused - sem_ref).abs().max().item()),
                                  "vs": "fp32 module-by-module route (torch conv+BN+ReLU, TF32 off) on the same clouds"}
    return out


def train_arm(torch, dist, args, dev, rank, world, seqs, T, N, seed, barrier, max_ranks):
    """Config c4's fwd+bwd step on the hot path (canonical stage: the encoder trains, train_temporal.py:224-298): encoder in
    training mode through the operator route (our FPS / ball query / grouping / three_nn / interpolate kernels and their
    backward kernels; torch conv+BN+ReLU stacks), the reference's segmentation cross-entropy (smplx/loss/temporal_loss.py),
    gradients accumulated over micro-batches of sequences, ONE flat NCCL all-reduce of all gradients on a side stream
    (garment4d_b200/sharding.FlatGradientReducer; replaces DDP's bucketed reducer, train_temporal.py:186-187), Adam step.
    BatchNorm uses per-micro-batch statistics (no SyncBN: the north star allows a collective on gradients only)."""
    import torch.nn.functional as F
    from garment4d_b200.encoder import Pointnet2MSGSEG
    from garment4d_b200.sharding import FlatGradientReducer
    progress(rank, "train arm: build")
    torch.manual_seed(4321)
    model = Pointnet2MSGSEG(input_channels=0, bn=True, global_feat=False).to(dev).train()
    reducer = FlatGradientReducer(model.parameters())
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)          # train_temporal.py:121
    mb_seqs = 2                                                    # sequences per micro-batch (60 clouds)
    n_seq = len(seqs)
    pcs = torch.from_numpy(make_inputs("body", seed + 31, n_seq * T, N)).to(dev).view(n_seq, T, N, 3)
    labels = torch.from_numpy(np.random.RandomState(seed + 32).randint(0, 7, (n_seq, T, N))).to(dev)
    side = torch.cuda.Stream(device=dev)
    ev_ar = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))

    def one_step(timed_ar=False):
        reducer.zero()
        nmb = 0
        for lo in range(0, n_seq, mb_seqs):
            x = pcs[lo:lo + mb_seqs].reshape(-1, N, 3)
            y = labels[lo:lo + mb_seqs].reshape(-1)
            sem = model(x)[1]                                     # (clouds, N, 7)
            loss = F.cross_entropy(sem.reshape(-1, sem.shape[-1]), y)
            loss.backward()
            nmb += 1
        reducer.rebind()
        if nmb > 1:
            reducer.flat.div_(nmb)
        if world > 1:
            cur = torch.cuda.current_stream(dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                if timed_ar:
                    ev_ar[0].record()
                dist.all_reduce(reducer.flat)                    # the ONE data-path collective of a training step (NCCL)
                reducer.flat.div_(world)
                if timed_ar:
                    ev_ar[1].record()
            cur.wait_stream(side)
        opt.step()
        return loss

    progress(rank, "train arm: warm-up")
    for _ in range(2):                                             # the library picks its convolution / BatchNorm kernels in the first passes
        one_step()
    barrier("train: warm-up done")
    steps = max(1, args.train_steps)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        loss = one_step(timed_ar=True)
    e.record()
    barrier("train: timed steps done")
    ms = max_ranks(s.elapsed_time(e) / steps)
    ar_us = max_ranks(ev_ar[0].elapsed_time(ev_ar[1]) * 1e3) if world > 1 else 0.0
    B = CONFIGS[args.config][0]
    Cg = B * T * (world if args.scaling == "weak" else 1)
    return {"workload": f"{args.config} fwd+bwd: encoder in training mode (operator route) + seg cross-entropy + flat gradient all-reduce + Adam",
            "value": Cg / (ms * 1e-3), "unit": "frames/s", "ms_per_step": ms, "steps": steps, "micro_batch_frames": mb_seqs * T,
            "allreduce_us": ar_us, "allreduce_bytes": reducer.nbytes,
            "collective": "NCCL all_reduce (sum) of one flat fp32 buffer, side stream" if world > 1 else "none (1 rank)",
            "loss": float(loss.item())}


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
    algorithmic-work roofline of each (bytes / flops per cloud from SURVEY.md section 8(d)).  `ref_ms` = the reference's own
    CUDA kernel (oracle/_ref: its .cu files compiled unmodified for sm_100a) on the same inputs, where that library was built."""
    from garment4d_b200 import lbs as glbs
    from garment4d_b200.pointnet2 import pointnet2_utils as pu
    import ctypes
    from garment4d_b200 import _lib
    out = []
    refgpu = None
    try:
        from oracle import refgpu as _rg          # baseline timing only (the "reference on B200" column); never on the product path
        if _rg.available():
            refgpu = _rg
    except Exception:
        refgpu = None

    def t(fn, reps=reps):
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

    def tref(fn):
        if refgpu is None:
            return None
        try:
            return t(fn, reps=2)
        except Exception:
            return None

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

    def hbm(name, ms, bytes_per_cloud, note="", ncu=None, ref_ms=None):
        ach = bytes_per_cloud * C / (ms * 1e-3) / 1e9
        out.append({"name": name, "ms": ms, "ref_ms": ref_ms,
                    "roofline": {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                 "frac": ach / peaks["hbm_gbs"], "traffic": traffic(ncu)}, "note": note})

    def tensor(name, ms, macs_per_cloud, ncu=None, ref_ms=None):
        ach = 2.0 * macs_per_cloud * C / (ms * 1e-3) / 1e12
        out.append({"name": name, "ms": ms, "ref_ms": ref_ms,
                    "roofline": {"bound": "tensor", "achieved": ach, "peak": peaks["tc_tflops"], "unit": "TFLOP/s",
                                 "frac": ach / peaks["tc_tflops"], "traffic": traffic(ncu)}})

    def latency(name, ms, steps, bytes_per_cloud, ncu=None, ref_ms=None):
        ns = ms * 1e6 / steps
        out.append({"name": name, "ms": ms, "ref_ms": ref_ms,
                    "roofline": {"bound": "latency", "achieved": ns, "peak": FPS_STEP_FLOOR_NS, "unit": "ns per serial step",
                                 "frac": FPS_STEP_FLOOR_NS / ns, "traffic": traffic(ncu),
                                 "hbm_gbs": bytes_per_cloud * C / (ms * 1e-3) / 1e9},
                    "note": f"{steps} dependent arg-max steps per cloud; floor = one step with no distance update ({FPS_STEP_FLOOR_NS:.0f} ns, measured)"})

    sa_total = 0.0
    with torch.no_grad():
        xyz, feats = pc.contiguous(), None
        lx, lf = [xyz], [None]
        for lvl, sa in enumerate(model.SA_modules):
            n_in, P = xyz.shape[1], sa.npoint
            ms = t(lambda: pu.furthest_point_sample_and_gather(xyz, P))
            latency(f"fps_gather L{lvl} ({n_in}->{P})", ms, P - 1, 12 * n_in + 16 * P,
                    ncu=(("fps_pruned_kernel<512", 0), ("fps_kernel<256, 4", 0), ("fps_kernel<256, 1", 0))[lvl] if lvl < 3 else None,
                    ref_ms=tref(lambda: refgpu.furthest_point_sample(xyz, P)))
            _, new_xyz = pu.furthest_point_sample_and_gather(xyz, P)
            g0, g1 = sa.groupers
            ms = t(lambda: pu.ball_query_pair(g0.radius, g0.nsample, g1.radius, g1.nsample, xyz, new_xyz))
            hbm(f"ball_query2 L{lvl}", ms, 12 * n_in + 12 * P + 4 * P * (g0.nsample + g1.nsample),
                ncu=("ball_query_grid_kernel<2>", 0) if lvl == 0 else ("ball_query_kernel<2>", lvl - 1),
                ref_ms=tref(lambda: (refgpu.ball_query(g0.radius, g0.nsample, xyz, new_xyz), refgpu.ball_query(g1.radius, g1.nsample, xyz, new_xyz))))
            idxs = pu.ball_query_pair(g0.radius, g0.nsample, g1.radius, g1.nsample, xyz, new_xyz)
            c_in = 0 if feats is None else feats.shape[1]
            # the north star's fused ball-query+group operator (materialises the grouped tensor; not on the fused route)
            for gi, g in enumerate((g0, g1)):
                ms = t(lambda: pu.QueryAndGroup(g.radius, g.nsample)(xyz, new_xyz, feats))
                bq = ("ball_query_grid_kernel<1>", gi) if lvl == 0 else ("ball_query_kernel<1>", 2 * (lvl - 1) + gi)

                def ref_qg(g=g):
                    # the reference's QueryAndGroup.forward operator sequence on its own kernels (pointnet2_utils.py:250-258)
                    idx = refgpu.ball_query(g.radius, g.nsample, xyz, new_xyz)
                    gx = refgpu.grouping_operation(xyz.transpose(1, 2).contiguous(), idx)
                    gx -= new_xyz.transpose(1, 2).unsqueeze(-1)
                    return gx if feats is None else torch.cat([gx, refgpu.grouping_operation(feats, idx)], dim=1)
                hbm(f"query_and_group L{lvl} K={g.nsample}", ms,
                    12 * n_in + 12 * P + 4 * c_in * n_in + 4 * P * g.nsample + 4 * (c_in + 3) * P * g.nsample,
                    ncu=[bq, ("group_fused_kernel", gi) if lvl == 0 else ("group_rows_kernel", 2 * (lvl - 1) + gi)], ref_ms=tref(ref_qg))
            new_xyz2, new_feats = sa(xyz, feats)
            feat_pm = None if feats is None else pu.point_major_of(feats)
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
                sa_total += ms
                d = br.desc

                def ref_branch(i=i, g=g):
                    # the reference's module sequence for this scale: QueryAndGroup (its kernels) -> SharedMLP (cuDNN) -> max_pool2d
                    idx = refgpu.ball_query(g.radius, g.nsample, xyz, new_xyz)
                    gx = refgpu.grouping_operation(xyz.transpose(1, 2).contiguous(), idx)
                    gx -= new_xyz.transpose(1, 2).unsqueeze(-1)
                    x = gx if feats is None else torch.cat([gx, refgpu.grouping_operation(feats, idx)], dim=1)
                    y = sa.mlps[i](x)
                    return torch.nn.functional.max_pool2d(y, kernel_size=[1, y.size(3)])
                tensor(f"sa_mlp_max L{lvl} K={g.nsample} ({c_in}+3->{d.c1},{d.c2},{d.c3})", ms,
                       branch_macs(g.nsample, c_in, (d.c1, d.c2, d.c3), P),
                       ncu=(f"sa_mlp_max_kernel<{g.nsample}, {1 if c_in else 0}", 1 if (lvl == 2 and g.nsample == 32) else 0),
                       ref_ms=tref(ref_branch))
                out[-1]["ref_note"] = "reference module sequence for this scale: its ball query + grouping kernels, SharedMLP on cuDNN, max_pool2d"
                off += br.c_out
            xyz, feats = new_xyz2, new_feats
            lx.append(xyz); lf.append(feats)
        total_macs = sum(branch_macs(K, ci, mlp, P) for _, _, P, K, ci, mlp in SA_BRANCHES)
        ach = 2.0 * total_macs * C / (sa_total * 1e-3) / 1e12
        out.append({"name": "sa_mlp_max, all six branches", "ms": sa_total,
                    "aggregate_roofline": {"bound": "tensor", "achieved": ach, "peak": peaks["tc_tflops"], "unit": "TFLOP/s",
                                           "frac": ach / peaks["tc_tflops"]}})
        # feature propagation: FP2 / FP1 modules, then the fused finest level + head
        from garment4d_b200.pointnet2 import pointnet2_cuda_bridge as bridge
        fp2, fp1 = model.FP_modules[2], model.FP_modules[1]
        ms = t(lambda: fp2(lx[2], lx[3], lf[2], lf[3]))
        f2 = fp2(lx[2], lx[3], lf[2], lf[3])
        out.append({"name": "FP2 module (three_nn + interpolation + 2-layer MLP)", "ms": ms})
        ms = t(lambda: fp1(lx[1], lx[2], lf[1], f2))
        f1 = fp1(lx[1], lx[2], lf[1], f2)
        out.append({"name": "FP1 module (three_nn + interpolation + 2-layer MLP)", "ms": ms})
        d2 = torch.empty(C, N, 3, dtype=torch.float32, device=pc.device)
        i3 = torch.empty(C, N, 3, dtype=torch.int32, device=pc.device)
        ms_nn = t(lambda: pu.three_nn_raw(lx[0], lx[1], d2, i3))
        hbm(f"three_nn L0 ({N} -> {lx[1].shape[1]}, grid)", ms_nn, 12 * N + 12 * lx[1].shape[1] + 24 * N, ncu=("three_nn_grid_kernel", 0),
            ref_ms=tref(lambda: refgpu.three_nn_raw(lx[0], lx[1])))
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
        out.append({"name": "FP stack + seg head", "ms": ms_all - ms_sa})
    return out


if __name__ == "__main__":
    main()
