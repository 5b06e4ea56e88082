"""CPU emulation of the pruned FPS's per-step work for different point orderings (no GPU): for every serial step, how many
clumps (threads) and warps take the update path.  Guides the ordering / clump-size choice of fps_pruned.cu."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from garment4d_b200 import synthetic


def morton3(cx, cy, cz, bits):
    key = np.zeros_like(cx, dtype=np.int64)
    for b in range(bits):
        key |= ((cx >> b) & 1) << (3 * b) | ((cy >> b) & 1) << (3 * b + 1) | ((cz >> b) & 1) << (3 * b + 2)
    return key


def order_coarse_rowmajor(p, cell):
    lo = p.min(0); c = np.floor((p - lo) / (cell * 1.0001)).astype(np.int64)
    d = c.max(0) + 1
    key = (c[:, 2] * d[1] + c[:, 1]) * d[0] + c[:, 0]
    return np.argsort(key, kind="stable")


def order_morton(p, bits):
    lo = p.min(0); ext = (p.max(0) - lo).max()
    h = ext / (1 << bits) * 1.0001
    c = np.minimum(np.floor((p - lo) / h).astype(np.int64), (1 << bits) - 1)
    return np.argsort(morton3(c[:, 0], c[:, 1], c[:, 2], bits), kind="stable")


def simulate(p, order, ppt, m, warp=32):
    q = p[order].astype(np.float64)
    n = len(q); nt = n // ppt
    cl = q.reshape(nt, ppt, 3)
    ctr = 0.5 * (cl.min(1) + cl.max(1))
    rad = np.sqrt(((cl - ctr[:, None]) ** 2).sum(2).max(1)) * 1.0001
    temp = np.full((nt, ppt), 1e10)
    thr = np.full(nt, np.inf)
    start = np.where(order == 0)[0][0]
    last = q[start]
    thr_cnt, warp_cnt, upd_pts = [], [], []
    for j in range(1, m):
        d2c = ((ctr - last) ** 2).sum(1)
        need = d2c < thr
        if j == 1: need[:] = True
        thr_cnt.append(need.sum()); warp_cnt.append(need.reshape(-1, warp).any(1).sum())
        idxs = np.where(need)[0]
        d = ((cl[idxs] - last) ** 2).sum(2)
        upd_pts.append((d < temp[idxs]).sum())
        temp[idxs] = np.minimum(temp[idxs], d)
        tmax = temp[idxs].max(1)
        thr[idxs] = (rad[idxs] + np.sqrt(tmax)) ** 2 * 1.0002
        flat = temp.argmax()
        last = cl.reshape(-1, 3)[flat]
    return np.array(thr_cnt), np.array(warp_cnt), np.array(upd_pts), nt


if __name__ == "__main__":
    kind = sys.argv[1] if len(sys.argv) > 1 else "body"
    N, m = 8192, 1024
    if kind == "body":
        p = synthetic.body_clouds(4234, 1, N)[0]
    else:
        p = np.random.RandomState(1).rand(N, 3).astype(np.float32)
    for name, order in [("coarse r=0.1 row-major", order_coarse_rowmajor(p, 0.1)), ("morton 4b", order_morton(p, 4)), ("morton 5b", order_morton(p, 5)),
                        ("morton 6b", order_morton(p, 6)), ("morton 7b", order_morton(p, 7)), ("morton 10b", order_morton(p, 10))]:
        for ppt in (16, 8, 4):
            t, w, u, nt = simulate(p, order, ppt, m)
            nw = nt // 32
            print(f"{kind:5s} {name:24s} ppt={ppt:2d} threads={nt:5d} warps={nw:3d}: need threads/step {t[1:].mean():7.1f} ({100*t[1:].mean()/nt:4.1f}%)  "
                  f"warps/step {w[1:].mean():5.2f} ({100*w[1:].mean()/nw:4.1f}%)  points that change/step {u[1:].mean():6.1f}  pts tested/step {t[1:].mean()*ppt:7.0f}")


def cost_model(p, order, m, design):
    """Issue-cycle model per step: max over the 4 SM sub-partitions of the instructions their warps execute
    (warp w lives on sub-partition w % 4) + nothing else.  design: ('thread', T, PPT, base, upd) thread-owned clumps;
    ('row', T, PPT, base, skip, per_clump) warp-row clumps of 32 points."""
    q = p[order].astype(np.float64); n = len(q)
    kind, T, PPT = design[:3]
    if kind == 'thread':
        cl = q.reshape(T, PPT, 3)                       # thread t owns PPT consecutive points
    else:
        cl = q.reshape(T // 32 * PPT, 32, 3)            # clump = 32 consecutive points; warp w owns clumps w*PPT .. w*PPT+PPT-1
    nc = cl.shape[0]; cpw = nc // (T // 32)
    ctr = 0.5 * (cl.min(1) + cl.max(1)); rad = np.sqrt(((cl - ctr[:, None]) ** 2).sum(2).max(1)) * 1.0001
    temp = np.full(cl.shape[:2], 1e10); thr = np.full(nc, np.inf)
    last = q[np.where(order == 0)[0][0]]
    tot = 0.0; instr = 0.0
    for j in range(1, m):
        need = ((ctr - last) ** 2).sum(1) < thr
        if j == 1: need[:] = True
        idxs = np.where(need)[0]
        temp[idxs] = np.minimum(temp[idxs], ((cl[idxs] - last) ** 2).sum(2))
        thr[idxs] = (rad[idxs] + np.sqrt(temp[idxs].max(1))) ** 2 * 1.0002
        last = cl.reshape(-1, 3)[temp.argmax()]
        per_warp = need.reshape(-1, cpw)
        if kind == 'thread':
            w = design[3] + design[4] * per_warp.any(1)
        else:
            cnt = per_warp.sum(1)
            w = design[3] + (cnt > 0) * design[4] + cnt * design[5]
        sp = w.reshape(-1, 4).sum(0)
        if j > 1: tot += sp.max(); instr += w.sum()
    return tot / (m - 2), instr / (m - 2)


if __name__ == "__main__" and len(sys.argv) > 2:
    print("cost model (issue cycles on the busiest sub-partition per step; total warp-instr per step)")
    designs = {"A now: thread clumps T512 P16": ('thread', 512, 16, 40, 280), "B thread clumps T1024 P8": ('thread', 1024, 8, 40, 150),
               "C warp-row T512 16x32": ('row', 512, 16, 40, 32, 33), "D warp-row T1024 8x32": ('row', 1024, 8, 40, 16, 33),
               "E warp-row T256 32x32": ('row', 256, 32, 40, 64, 33)}
    for oname, order in [("coarse row-major", order_coarse_rowmajor(p, 0.1)), ("morton 5b", order_morton(p, 5)), ("morton 10b", order_morton(p, 10))]:
        for dname, d in designs.items():
            c, i = cost_model(p, order, m, d)
            print(f"{kind:5s} {oname:18s} {dname:32s} busiest sub-partition {c:7.1f}   total {i:7.1f}")


if __name__ == "__main__" and len(sys.argv) > 3:
    print("warp-row design: flagged 32-point clumps per step / warps with any flagged clump")
    for oname, order in [("coarse row-major", order_coarse_rowmajor(p, 0.1)), ("morton 4b", order_morton(p, 4)), ("morton 5b", order_morton(p, 5))]:
        q = p[order].astype(np.float64)
        cl = q.reshape(-1, 32, 3); nc = cl.shape[0]
        ctr = 0.5 * (cl.min(1) + cl.max(1)); rad = np.sqrt(((cl - ctr[:, None]) ** 2).sum(2).max(1)) * 1.0001
        temp = np.full(cl.shape[:2], 1e10); thr = np.full(nc, np.inf)
        last = q[np.where(order == 0)[0][0]]
        nf, nw, mx = [], [], []
        for j in range(1, m):
            need = ((ctr - last) ** 2).sum(1) < thr
            if j == 1: need[:] = True
            idxs = np.where(need)[0]
            temp[idxs] = np.minimum(temp[idxs], ((cl[idxs] - last) ** 2).sum(2))
            thr[idxs] = (rad[idxs] + np.sqrt(temp[idxs].max(1))) ** 2 * 1.0002
            last = cl.reshape(-1, 3)[temp.argmax()]
            pw = need.reshape(-1, 16).sum(1)
            if j > 1: nf.append(need.sum()); nw.append((pw > 0).sum()); mx.append(pw.max())
        print(f"{kind:5s} {oname:18s} flagged clumps/step {np.mean(nf):5.2f}  warps touched {np.mean(nw):4.2f}  max flagged in one warp: mean {np.mean(mx):4.2f} p95 {np.percentile(mx,95):.0f}")
