"""Per-layer efficiency of the backbone at bench chunk size (writes gpurun_out/layer_eff.txt)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rgbmanip_b200 import weights
from rgbmanip_b200.engine import Engine

E = int(sys.argv[1]) if len(sys.argv) > 1 else 32
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16x2"
sd = weights.init_state_dict(0)
eng = Engine(sd, max_envs=E, precision=prec)
F = 2 * E
eng.crops.normal_()
def flops(name):
    w = None
    key = name
    if name == "conv1": w = sd["img_extractor.feats.conv1.weight"]; hw = 112 * 112
    elif name.startswith("img_extractor.feats.layer"):
        if name.endswith(".down"): w = sd[name.replace(".down", ".downsample.0.weight")]
        else: w = sd[name + ".weight"]
        li = int(name.split("layer")[1][0]); hw = {1: 56 * 56, 2: 28 * 28, 3: 28 * 28, 4: 28 * 28}[li]
    elif name in ("up_1", "up_2", "up_3"):
        w = sd[f"img_extractor.{name}.conv.0.weight"]; hw = {"up_1": 56 * 56, "up_2": 112 * 112, "up_3": 224 * 224}[name]
    elif name == "final": w = sd["img_extractor.final.weight"]; hw = 224 * 224
    if w is None: return 0.0
    return 2.0 * hw * np.prod(w.shape)
ev = lambda: torch.cuda.Event(enable_timing=True)
for it in range(2):
    rec = []
    for name, op in eng.backbone_ops:
        a, b = ev(), ev(); a.record(); op(F); b.record(); rec.append((name, a, b, getattr(op, "kind", "-")))
    torch.cuda.synchronize()
lines = []
tot = 0.0
for name, a, b, kind in rec:
    ms = a.elapsed_time(b); tot += ms
    fl = flops(name) * F
    lines.append(f"{name.replace('img_extractor.feats.',''):22s} {kind:4s} {ms:8.3f} ms  {fl/1e9:9.1f} GF  {fl/ms/1e9 if ms>0 else 0:8.1f} TF/s alg")
lines.append(f"total {tot:.3f} ms for {F} frames, precision {prec}")
os.makedirs("gpurun_out", exist_ok=True)
open(f"gpurun_out/layer_eff_{prec}.txt", "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
