"""CUDA-event time of every stereo stage (volume, cost-reg ops, decode pieces, fit) for one chunk.
usage: python tools/time_stereo_tail.py [chunk_envs]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rgbmanip_b200 import synth, weights
from rgbmanip_b200.engine import Engine
E = int(sys.argv[1]) if len(sys.argv) > 1 else 64
tc = True
eng = Engine(weights.init_state_dict(0), max_envs=E)
b = synth.make_batch(E, seed=1, special=False, n_unique=8)
dev = eng.device
t = lambda a, dt=None: (torch.from_numpy(np.ascontiguousarray(a)).to(dt) if dt else torch.from_numpy(np.ascontiguousarray(a))).to(dev)
K, E1, E2 = t(b.K, torch.float64), t(b.E1, torch.float64), t(b.E2, torch.float64)
for it in range(2):
    eng.run_chunk(K, t(b.rgb1), t(b.mask1), E1, t(b.rgb2), t(b.mask2), E2)
torch.cuda.synchronize()
acc = {}
for it in range(5):
    evs = [("start", torch.cuda.Event(enable_timing=True))]
    evs[0][1].record()
    def mark(name):
        e = torch.cuda.Event(enable_timing=True); e.record(); evs.append((name, e))
    if tc:
        orig = eng.dec_ops
        eng.dec_ops = [(n, (lambda k, op=op, n=n: (op(k), mark("dec." + n)))) for n, op in orig]
    eng.stereo(E, E1, E2, mark=mark)
    if tc:
        eng.dec_ops = orig
    torch.cuda.synchronize()
    for (n0, e0), (n1, e1) in zip(evs[:-1], evs[1:]):
        acc.setdefault(n1, []).append(e0.elapsed_time(e1))
eng.check_error_flag()
tot = 0.0
for k, v in acc.items():
    m = float(np.median(v)); tot += m
    print(f"{k:18s} {m:8.3f} ms  ({m / E * 1e3:7.2f} us/env)")
print(f"total {tot:.3f} ms for {E} envs")
