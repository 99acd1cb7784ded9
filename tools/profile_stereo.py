"""One warm-up chunk, then one chunk of the whole pipeline between cudaProfilerStart/Stop (for ncu)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rgbmanip_b200 import synth, weights
from rgbmanip_b200.engine import Engine
E = int(sys.argv[1]) if len(sys.argv) > 1 else 16
eng = Engine(weights.init_state_dict(0), max_envs=E)
b = synth.make_batch(E, seed=1, special=False, n_unique=4)
dev = eng.device
t = lambda a, dt=None: (torch.from_numpy(np.ascontiguousarray(a)).to(dt) if dt else torch.from_numpy(np.ascontiguousarray(a))).to(dev)
K, E1, E2 = t(b.K, torch.float64), t(b.E1, torch.float64), t(b.E2, torch.float64)
args = (K, t(b.rgb1), t(b.mask1), E1, t(b.rgb2), t(b.mask2), E2)
eng.run_chunk(*args)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()      # ncu --profile-from-start off: exactly one steady-state chunk is captured
eng.run_chunk(*args)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
