"""One backbone pass at bench chunk size (for ncu)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rgbmanip_b200 import weights
from rgbmanip_b200.engine import Engine
E = int(sys.argv[1]) if len(sys.argv) > 1 else 64
eng = Engine(weights.init_state_dict(0), max_envs=E)
eng.crops.normal_()
for it in range(2):
    eng.run_backbone(2 * E)
torch.cuda.synchronize()
