"""Scratch experiment: emulate the device path's operand/storage rounding inside the oracle and report
the end-to-end error against the fp32 oracle (px, deg, mm).  Guides the precision policy in DESIGN.md."""
import sys, os, time
import numpy as np, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import adapose_oracle as O
from rgbmanip_b200 import synth, weights

sd = weights.init_state_dict(0)
cfg = {"img_size": 224, "direct_regression": True}
batch = synth.make_batch(8, seed=0)
envs = [0, 1, 2, 3]

def rnd(x, dt):
    return x.to(dt).float() if dt is not None else x

def run(op2d=None, op3d=None, store2d=None, store3d=None, store_feat=None, resid_fp32=True, label=""):
    c2, c3, ct3 = F.conv2d, F.conv3d, F.conv_transpose3d
    def conv2d(x, w, *a, **k):
        if op2d == "bf16x3":
            b = k.pop("bias", None) if "bias" in k else (a[0] if a else None)
            a2 = a[1:] if a else a
            xh = x.bfloat16().float(); xl = (x - xh).bfloat16().float()
            wh = w.bfloat16().float(); wl = (w - wh).bfloat16().float()
            return c2(xh, wh, b, *a2, **k) + c2(xl, wh, None, *a2, **k) + c2(xh, wl, None, *a2, **k)
        if op2d == "fp16a":      # activations single fp16, weights split hi+lo (2 MMA passes)
            return c2(x.half().float(), w, *a, **k)
        return c2(rnd(x, op2d), rnd(w, op2d), *a, **k)
    def conv3d(x, w, *a, **k): return c3(rnd(x, op3d), rnd(w, op3d), *a, **k)
    def convt3d(x, w, *a, **k): return ct3(rnd(x, op3d), rnd(w, op3d), *a, **k)
    O.F.conv2d, O.F.conv3d, O.F.conv_transpose3d = conv2d, conv3d, convt3d
    def q(name, x):
        if name == "feat": return rnd(x, store_feat)
        if name.startswith("cr.") or name == "fused1":
            return rnd(x, store3d) if name != "cr.prob" else x
        if name in ("logits", "fused_pts"): return x
        if resid_fp32 and name.startswith("img_extractor.feats.layer") and name.count(".") == 3 and not name.endswith(("conv1", "down")):
            return x
        return rnd(x, store2d)
    tap = O.Taps(q=q)
    res = []
    try:
        np.random.seed(0)
        for e in range(max(envs) + 1):
            if e not in envs:
                O.prepare_model_input(batch.rgb1[e], batch.mask1[e], batch.K[e]); O.prepare_model_input(batch.rgb2[e], batch.mask2[e], batch.K[e]); continue
            d = {}
            box = O.predict(sd, cfg, batch.K[e], batch.rgb1[e], batch.mask1[e], batch.E1[e], batch.rgb2[e], batch.mask2[e], batch.E2[e], both_views=False, tap=tap, details=d)
            res.append((box, d["depth"], d["nocs"], d["R"]))
    finally:
        O.F.conv2d, O.F.conv3d, O.F.conv_transpose3d = c2, c3, ct3
    return res

t0 = time.time()
ref = run(label="fp32")
print("fp32 ref done", time.time() - t0)
bf, hf = torch.bfloat16, torch.float16
variants = {
  "bf16x3 2d operands, fp32 storage": dict(op2d="bf16x3"),
  "bf16x3 2d + bf16 3d": dict(op2d="bf16x3", op3d=bf, store3d=bf),
  "bf16x3 2d + fp16 3d": dict(op2d="bf16x3", op3d=hf, store3d=hf),
  "bf16 2d operands only": dict(op2d=bf),
  "bf16 2d operands + bf16 storage, fp32 residual stream, fp32 feat": dict(op2d=bf, store2d=bf),
  "bf16 2d all (bf16 residual, bf16 feat)": dict(op2d=bf, store2d=bf, store_feat=bf, resid_fp32=False),
  "bf16 2d, fp32 resid, bf16 feat": dict(op2d=bf, store2d=bf, store_feat=bf),
  "fp16a 2d (fp16 act operands+storage, split weights) + fp16 3d": dict(op2d="fp16a", store2d=hf, store_feat=hf, resid_fp32=False, op3d=hf, store3d=hf),
  "fp16a 2d fp32resid + fp16 3d": dict(op2d="fp16a", store2d=hf, store_feat=hf, resid_fp32=True, op3d=hf, store3d=hf),
  "fp16 2d all": dict(op2d=hf, store2d=hf, store_feat=hf, resid_fp32=False),
  "bf16 3d only (operands+storage)": dict(op3d=bf, store3d=bf),
  "fp16 3d only": dict(op3d=hf, store3d=hf),
  "bf16 everything, fp32 resid": dict(op2d=bf, store2d=bf, store_feat=bf, op3d=bf, store3d=bf),
  "bf16 2d (fp32 resid, fp16 feat) + fp16 3d": dict(op2d=bf, store2d=bf, store_feat=hf, op3d=hf, store3d=hf),
}
sel = sys.argv[1:]
for name, kw in variants.items():
    if sel and not any(s in name for s in sel): continue
    out = run(**kw)
    errs = []
    for (b0, d0, n0, R0), (b1, d1, n1, R1), e in zip(ref, out, envs):
        px, deg, mm, cmm = O.parity_errors(b1, b0, batch.K[e], batch.E1[e])
        errs.append((px, deg, mm, cmm, float(np.abs(d1 - d0).max() * 1e3), float(np.abs(n1 - n0).max())))
    errs = np.array(errs)
    print(f"{name:70s} max px {errs[:,0].max():.4f} deg {errs[:,1].max():.4f} ctr-mm {errs[:,2].max():.3f} corner-mm {errs[:,3].max():.3f} depth-mm {errs[:,4].max():.2f} nocs {errs[:,5].max():.5f}", flush=True)
