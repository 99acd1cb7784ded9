import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rgbmanip_b200 import weights
from rgbmanip_b200.engine import Engine
eng = Engine(weights.init_state_dict(0), max_envs=16)
eng.vol.hi.normal_()
op = dict(eng.cr_ops)["cr.conv0"]
for _ in range(2): op(16)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): op(16)
b.record(); torch.cuda.synchronize()
print("ADP_C0_DBG", os.environ.get("ADP_C0_DBG"), "conv0 ms per 16 envs:", a.elapsed_time(b) / 5)
