"""Per-env parity table + coarse stage timing on the GPU box (writes gpurun_out/diag.txt)."""
import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import adapose_oracle as O
from rgbmanip_b200 import synth, weights
from rgbmanip_b200.estimator import AdaPoseEstimator_v5

g = np.load("tests/golden/e2e.npz")
batch = synth.make_batch(8, seed=0)
cfg = {"load": False, "direct_regression": True, "img_size": 224}
sd = weights.init_state_dict(0)
c1 = np.zeros((8, 1024), np.int32); c2 = np.zeros((8, 1024), np.int32)
for e in range(8):
    if g["valid"][e]: c1[e], c2[e] = g[f"env{e}_choose1"], g[f"env{e}_choose2"]
lines = []
for prec in ("bf16x3", "bf16"):
    est = AdaPoseEstimator_v5(None, cfg, None, state_dict=sd, max_envs=8, precision=prec)
    boxes = est.estimate(*batch.args(), choose=(c1, c2))
    eng = est.estimator
    for e in range(8):
        if not g["valid"][e]: continue
        px, deg, mm, cmm = O.parity_errors(boxes[e], g["boxes"][e], batch.K[e], batch.E1[e])
        dn = np.abs(eng.nocs[e].cpu().numpy() - g[f"env{e}_view1_nocs"]).max()
        dd = np.abs(eng.depth[e].cpu().numpy() - g[f"env{e}_view1_depth"])
        rr = O.rotation_angle_deg(eng.R[e].cpu().numpy().reshape(3, 3), g[f"env{e}_view1_r"])
        f = eng.feat[e].cpu().numpy().transpose(2, 0, 1)[:, ::8, ::8]
        df = np.abs(f - g[f"env{e}_feat1_sub"]).max() / np.abs(g[f"env{e}_feat1_sub"]).max()
        zc = (batch.E1[e][:3, :3] @ g["boxes"][e].T + batch.E1[e][:3, 3:4])[2]
        lines.append(f"{prec:7s} env{e} px {px:.4f} deg {deg:.5f} ctr-mm {mm:.4f} corner-mm {cmm:.4f} | nocs {dn:.2e} depth mean {dd.mean()*1e3:.3f}mm max {dd.max()*1e3:.3f}mm R {rr:.5f} feat-rel {df:.2e} s {float(eng.scale[e]):.5f} zmin {zc.min():.2f}")
    # timing
    ev = lambda: torch.cuda.Event(enable_timing=True)
    N = eng.E
    dev = eng.device
    t = lambda a, dt=None: (torch.from_numpy(np.ascontiguousarray(a)).to(dt) if dt else torch.from_numpy(np.ascontiguousarray(a))).to(dev)
    K, E1, E2 = t(batch.K, torch.float64), t(batch.E1, torch.float64), t(batch.E2, torch.float64)
    r1, r2, m1, m2 = t(batch.rgb1), t(batch.rgb2), t(batch.mask1), t(batch.mask2)
    for it in range(3):
        marks = [ev() for _ in range(6)]
        marks[0].record()
        eng.preprocess(0, r1, m1, K, N); eng.preprocess(1, r2, m2, K, N); marks[1].record()
        eng.run_backbone(2 * N); marks[2].record()
        eng.stereo(N, E1, E2); marks[3].record()
        torch.cuda.synchronize()
    evs = [("start", ev())]
    evs[0][1].record()
    def mark(name):
        e = ev(); e.record(); evs.append((name, e))
    eng.stereo(N, E1, E2, mark=mark)
    torch.cuda.synchronize()
    lines.append(f"{prec} stereo stages ms: " + ", ".join(f"{evs[i][0]}={evs[i-1][1].elapsed_time(evs[i][1]):.3f}" for i in range(1, len(evs))))
    lines.append(f"{prec} timing N={N}: preprocess {marks[0].elapsed_time(marks[1]):.2f} ms, backbone({2*N} frames) {marks[1].elapsed_time(marks[2]):.2f} ms, stereo {marks[2].elapsed_time(marks[3]):.2f} ms")
    # per-op timing
    for group, ops in (("backbone", eng.backbone_ops), ("costreg", eng.cr_ops)):
        n = 2 * N if group == "backbone" else N
        tt = []
        for name, op in ops:
            a, b = ev(), ev()
            a.record(); op(n); b.record(); torch.cuda.synchronize()
            tt.append((name, a.elapsed_time(b), getattr(op, "kind", "-")))
        lines.append(f"{prec} {group} per-op ms: " + ", ".join(f"{n.replace('img_extractor.feats.','')}[{k}]={v:.3f}" for n, v, k in tt))
    eng.close()
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/diag.txt", "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
