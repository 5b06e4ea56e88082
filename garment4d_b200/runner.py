"""Stream-pipelined driver of the hot path: PointNet++ encoder forward + SMPL lbs() for a batch of frames.

Frames are independent, so a step is cut into `chunks` groups of frames, each issued on its own
CUDA stream: the host->device copy of chunk i+1, the kernels of chunk i and the device->host copy of chunk i-1 overlap,
and the latency-bound kernels of different chunks (FPS is a serial chain per cloud, the grouped MLP hands off between
warps) fill each other's idle SMs.  Results are identical to one big call: every kernel works per cloud.
"""
import torch

from . import _lib
from . import lbs as glbs
from .pointnet2 import pointnet2_cuda_bridge as _bridge


class EncoderLBSRunner:
    """model: garment4d_b200.encoder.Pointnet2MSGSEG (eval);  smpl: [v_template, shapedirs, posedirs, J_regressor, parents,
    lbs_weights] device tensors."""

    def __init__(self, model, smpl, chunks=4, device=None):
        self.model = model
        self.smpl = smpl
        self.chunks = max(1, int(chunks))
        self.device = device if device is not None else next(model.parameters()).device
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(self.chunks)]
        self.lbs_stream = torch.cuda.Stream(device=self.device)    # lbs() does not depend on the encoder: its own stream

    def _bounds(self, C):
        n = min(self.chunks, C)
        base, extra = divmod(C, n)
        out, lo = [], 0
        for i in range(n):
            hi = lo + base + (1 if i < extra else 0)
            out.append((lo, hi))
            lo = hi
        return out

    @torch.no_grad()
    def forward_device(self, pc, betas, pose):
        """Inputs resident on the device.  Returns (sem_logits (C,N,classes), verts (C,V,3), joints (C,J,3))."""
        C = pc.shape[0]
        _lib.lib().g4d_fps_concurrency_hint(int(C))              # the frame groups' FPS launches share the GPU: C clouds at once
        cur = torch.cuda.current_stream(self.device)
        sems = []
        capturing = torch.cuda.is_current_stream_capturing()
        self.lbs_stream.wait_stream(cur)
        with torch.cuda.stream(self.lbs_stream):
            verts, joints = glbs.lbs(betas, pose, *self.smpl)
            if not capturing:
                verts.record_stream(cur); joints.record_stream(cur)
        for (lo, hi), st in zip(self._bounds(C), self.streams):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                _, sem, _, _ = self.model(pc[lo:hi])
                if not capturing:
                    sem.record_stream(cur)
            sems.append(sem)
        for st in self.streams:
            cur.wait_stream(st)
        cur.wait_stream(self.lbs_stream)
        _lib.lib().g4d_fps_concurrency_hint(0)
        return torch.cat(sems), verts, joints

    @torch.no_grad()
    def forward_host(self, pc_pin, betas_pin, pose_pin, labels_pin, verts_pin, joints_pin):
        """End to end with pinned HOST buffers: per chunk H2D of the clouds -> encoder -> argmax labels (uint8) -> D2H; the SMPL
        parameters go up, lbs() runs for all frames and the posed vertices / joints (90 % of the result bytes) come back on the
        lbs stream, overlapped with the encoder.  Asynchronous: synchronise the current stream (or the device) before reading
        the outputs."""
        C = pc_pin.shape[0]
        _lib.lib().g4d_fps_concurrency_hint(int(C))
        cur = torch.cuda.current_stream(self.device)
        self.lbs_stream.wait_stream(cur)
        with torch.cuda.stream(self.lbs_stream):
            bt = betas_pin.to(self.device, non_blocking=True)
            ps = pose_pin.to(self.device, non_blocking=True)
            v, j = glbs.lbs(bt, ps, *self.smpl)
            verts_pin.copy_(v, non_blocking=True)
            joints_pin.copy_(j, non_blocking=True)
        for (lo, hi), st in zip(self._bounds(C), self.streams):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                pc = pc_pin[lo:hi].to(self.device, non_blocking=True)
                _, sem, _, _ = self.model(pc)
                # the segmentation the model consumes (mesh_encoder.py:113): written by the head kernel's last epilogue
                labels_pin[lo:hi].copy_(_bridge.segmentation_labels(sem), non_blocking=True)
        for st in self.streams:
            cur.wait_stream(st)
        cur.wait_stream(self.lbs_stream)
        _lib.lib().g4d_fps_concurrency_hint(0)


class GraphedEncoderLBSRunner(EncoderLBSRunner):
    """The same step captured once into CUDA graphs (shapes are static): one graph launch per step instead of ~100 kernel
    launches per chunk, which otherwise makes the host the bottleneck as soon as several chunks are in flight.

        r = GraphedEncoderLBSRunner(model, smpl, chunks=4)
        r.capture(pc, betas, pose)                         # device tensors of the step's shapes (contents irrelevant)
        sem, verts, joints = r.replay_device(pc, betas, pose)
        r.capture_host(pc_pin, betas_pin, pose_pin, labels_pin, verts_pin, joints_pin);  r.replay_host()
    """

    def __init__(self, model, smpl, chunks=4, device=None):
        super().__init__(model, smpl, chunks, device)
        self.g_dev = self.g_host = None

    def _warm(self, fn):
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(3):                 # builds every parameter cache / lazy kernel attribute outside the capture
                fn()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)

    def capture(self, pc, betas, pose):
        self.s_in = (pc.clone(), betas.clone(), pose.clone())
        self._warm(lambda: self.forward_device(*self.s_in))
        self.g_dev = torch.cuda.CUDAGraph()
        n0 = _lib.lib().g4d_launch_count()
        with torch.cuda.graph(self.g_dev):
            self.s_out = self.forward_device(*self.s_in)
        self.kernels_per_replay = int(_lib.lib().g4d_launch_count() - n0)     # libgarment4d_b200 kernel nodes in the graph
        return self

    def replay_device(self, pc=None, betas=None, pose=None):
        """pc/betas/pose = None: run on the contents already in the static input buffers (self.s_in)."""
        if pc is not None:
            self.s_in[0].copy_(pc); self.s_in[1].copy_(betas); self.s_in[2].copy_(pose)
        self.g_dev.replay()
        return self.s_out

    def capture_host(self, pc_pin, betas_pin, pose_pin, labels_pin, verts_pin, joints_pin):
        args = (pc_pin, betas_pin, pose_pin, labels_pin, verts_pin, joints_pin)
        self._warm(lambda: self.forward_host(*args))
        self.g_host = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_host):
            self.forward_host(*args)
        return self

    def replay_host(self):
        """H2D from / D2H into the pinned buffers given to capture_host (fill them before, read them after a synchronize)."""
        self.g_host.replay()
