"""AdaPose (StereoPoseNet_with_depth) parameter table, random init and checkpoint ingestion.

The table mirrors the reference module's ``state_dict()`` one-to-one (names, shapes, order):
  * backbone   models/pose_estimator/AdaPose/lib/pspnet.py:33-158
  * stereo net models/pose_estimator/AdaPose/lib/network_v5.py:260-376
Checkpoints written by the reference carry the ``module.`` prefix of ``nn.DataParallel``
(interface_v5.py:48,55-56); :func:`load_checkpoint` accepts both spellings.

Random init follows the reference constructor's distributions (pspnet.py:45-48 for the ResNet convs,
PyTorch defaults elsewhere) but draws from a numpy PCG64 stream so that the very same weights can be
regenerated on any machine (golden fixtures are produced with these weights loaded into the reference
module, see oracle/make_golden.py).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np

RESNET34_LAYERS = (3, 4, 6, 3)
RESNET_PLANES = (64, 128, 256, 512)
PSP_BINS = (1, 2, 3, 6)
BN_EPS = 1e-5


def _resnet_entries():
    ent = [("img_extractor.feats.conv1.weight", (64, 3, 7, 7), "resnet")]
    inpl = 64
    for li, (planes, blocks) in enumerate(zip(RESNET_PLANES, RESNET34_LAYERS), start=1):
        for b in range(blocks):
            pre = f"img_extractor.feats.layer{li}.{b}"
            cin = inpl if b == 0 else planes
            ent.append((f"{pre}.conv1.weight", (planes, cin, 3, 3), "resnet"))
            ent.append((f"{pre}.conv2.weight", (planes, planes, 3, 3), "resnet"))
            if b == 0 and (li == 2 or inpl != planes):
                ent.append((f"{pre}.downsample.0.weight", (planes, cin, 1, 1), "resnet"))
        inpl = planes
    return ent


def _conv_default(name, shape, bias=True):
    ent = [(f"{name}.weight", shape, "default")]
    if bias:
        ent.append((f"{name}.bias", (shape[0],), ("bias", int(np.prod(shape[1:])))))
    return ent


def _bn(name, c):
    return [(f"{name}.weight", (c,), "bn_w"), (f"{name}.bias", (c,), "bn_b"),
            (f"{name}.running_mean", (c,), "bn_m"), (f"{name}.running_var", (c,), "bn_v"),
            (f"{name}.num_batches_tracked", (), "bn_n")]


def param_table(regress_pose: bool = True, arch: str = "v5"):
    """[(name, shape, init-kind)] in the reference's state_dict order.  ``arch``: "v5" = StereoPoseNet_with_depth
    (network_v5.py:260-376); "baseline" = StereoPoseNet_with_depth_baseline (network_baseline.py:523-620: the cost volume and
    CostRegNet are replaced by 4 cross-view attention blocks, fusion.py:53-82, and a per-point depth MLP)."""
    if arch not in ("v5", "baseline"):
        raise ValueError(f"unknown architecture {arch!r}")
    ent = _resnet_entries()
    for s in range(4):
        ent.append((f"img_extractor.psp.stages.{s}.1.weight", (128, 512, 1, 1), "default"))
    for nm, cout, cin in (("up_1", 256, 1024), ("up_2", 64, 256), ("up_3", 64, 64)):
        ent += _conv_default(f"img_extractor.{nm}.conv.0", (cout, cin, 3, 3))
        ent.append((f"img_extractor.{nm}.conv.1.weight", (1,), "prelu"))
    ent += _conv_default("img_extractor.final", (32, 64, 1, 1))
    ent += _conv_default("instance_color.0", (64, 32, 1))
    if arch == "baseline":
        return ent + _baseline_tail(regress_pose)
    cr = "cost_regularization"
    for nm, cout, cin in (("conv0", 8, 32), ("conv1", 16, 8), ("conv2", 16, 16), ("conv3", 32, 16),
                          ("conv4", 32, 32), ("conv5", 64, 32), ("conv6", 64, 64)):
        ent.append((f"{cr}.{nm}.conv.weight", (cout, cin, 3, 3, 3), "default"))
        ent += _bn(f"{cr}.{nm}.bn", cout)
    for nm, cin, cout in (("conv7", 64, 32), ("conv9", 32, 16), ("conv11", 16, 8)):
        # ConvTranspose3d weights are (C_in, C_out, 3,3,3); torch computes fan_in from dim 1
        ent.append((f"{cr}.{nm}.conv.weight", (cin, cout, 3, 3, 3), "default"))
        ent += _bn(f"{cr}.{nm}.bn", cout)
    ent.append((f"{cr}.prob.weight", (1, 8, 3, 3, 3), "default"))
    ent += _conv_default("nocs_head.0", (128, 64, 1))
    ent += _conv_default("nocs_head.2", (64, 128, 1))
    ent += _conv_default("nocs_head.4", (3, 64, 1))
    if regress_pose:
        ent += _pose_entries()
    return ent


def _pose_entries():
    ent = _conv_default("nocs_pts_mlp.0", (32, 3, 1))
    ent += _conv_default("nocs_pts_mlp.2", (64, 32, 1))
    ent += _conv_default("pose_mlp1.0", (128, 96, 1))
    ent += _conv_default("pose_mlp1.2", (128, 128, 1))
    ent += _conv_default("pose_mlp2.0", (256, 256, 1))
    ent += _conv_default("pose_mlp2.2", (256, 256, 1))
    for head, nout in (("rotation_estimator", 6), ("translation_estimator", 3), ("size_estimator", 3)):
        ent += _conv_default(f"{head}.0", (256, 256))
        ent += _conv_default(f"{head}.2", (128, 256))
        ent += _conv_default(f"{head}.4", (nout, 128))
    return ent


FUSION_DEPTH, FUSION_HEADS, FUSION_DIM = 4, 4, 32      # ViewFusion(embed_dim=32, num_heads=4, depth=4), network_baseline.py:553


def _baseline_tail(regress_pose):
    """Everything behind ``instance_color`` of StereoPoseNet_with_depth_baseline, in constructor order."""
    ent = _conv_default("nocs_head.0", (128, 64, 1))
    ent += _conv_default("nocs_head.2", (64, 128, 1))
    ent += _conv_default("nocs_head.4", (3, 64, 1))
    for b in range(FUSION_DEPTH):
        for f in ("fusion1", "fusion2"):
            for l in range(4):        # q, k, v, output projections (fusion.py:33)
                ent += _conv_default(f"view_fusion.blocks.{b}.{f}.linears.{l}", (FUSION_DIM, FUSION_DIM))
    ent += _conv_default("depth_head.0", (64, 32, 1))
    ent += _conv_default("depth_head.2", (32, 64, 1))
    ent += _conv_default("depth_head.4", (1, 32, 1))
    if regress_pose:
        ent += _pose_entries()
    return ent


def init_state_dict(seed: int = 0, regress_pose: bool = True, randomize_bn: bool = True,
                    nocs_gain: float = 16.0, prob_gain: float = 4.0, arch: str = "v5", attn_gain: float = 1.0,
                    depth_bias: float = 0.8):
    """Random-init weights of the AdaPose architecture as ``OrderedDict[str, np.ndarray]``.

    ``randomize_bn`` perturbs the BatchNorm3d affine parameters and running statistics so that BN
    folding is exercised (the constructor default 1/0/0/1 is the identity).

    ``nocs_gain`` / ``prob_gain`` multiply the last NOCS layer's and the depth-logit conv's weights.  With
    the plain constructor init (gains 1) the NOCS map is almost constant (std ~0.01, the very threshold of
    the pair filter at utils.py:83) and the depth softmax is flat, so the scale fit divides by ~0 and
    amplifies any rounding by 1/|dNOCS|: box-level parity would measure that conditioning, not the
    kernels.  The default gains give NOCS a spread of a few tenths and a peaked depth distribution, like a
    trained network, while staying a seeded random init of the same architecture.  For ``arch="baseline"`` the same
    reasoning gives ``attn_gain`` (multiplies the query / key projections of every attention block, so the softmax is peaked
    rather than uniform) and ``depth_bias`` (added to the last depth-MLP bias: the head ends in a ReLU and predicts metres).
    """
    rng = np.random.default_rng(seed)
    sd = OrderedDict()
    for name, shape, kind in param_table(regress_pose, arch):
        if kind == "resnet":
            n = shape[2] * shape[3] * shape[0]
            w = rng.standard_normal(shape, dtype=np.float32) * np.float32(math.sqrt(2.0 / n))
        elif kind == "default":
            fan_in = int(np.prod(shape[1:]))
            b = 1.0 / math.sqrt(fan_in)
            w = rng.uniform(-b, b, size=shape).astype(np.float32)
        elif isinstance(kind, tuple) and kind[0] == "bias":
            b = 1.0 / math.sqrt(kind[1])
            w = rng.uniform(-b, b, size=shape).astype(np.float32)
        elif kind == "prelu":
            w = np.full(shape, 0.25, np.float32)
        elif kind == "bn_w":
            w = rng.uniform(0.6, 1.4, size=shape).astype(np.float32) if randomize_bn else np.ones(shape, np.float32)
        elif kind == "bn_b":
            w = (rng.standard_normal(shape) * 0.1).astype(np.float32) if randomize_bn else np.zeros(shape, np.float32)
        elif kind == "bn_m":
            w = (rng.standard_normal(shape) * 0.2).astype(np.float32) if randomize_bn else np.zeros(shape, np.float32)
        elif kind == "bn_v":
            w = rng.uniform(0.5, 1.5, size=shape).astype(np.float32) if randomize_bn else np.ones(shape, np.float32)
        elif kind == "bn_n":
            w = np.zeros((), np.int64)
        else:  # pragma: no cover
            raise AssertionError(kind)
        if name == "nocs_head.4.weight":
            w = w * np.float32(nocs_gain)
        elif name == "cost_regularization.prob.weight":
            w = w * np.float32(prob_gain)
        elif name.startswith("view_fusion.") and (".linears.0." in name or ".linears.1." in name):
            w = w * np.float32(attn_gain)
        elif name == "depth_head.4.bias":
            w = w + np.float32(depth_bias)
        sd[name] = w
    return sd


def strip_module_prefix(sd):
    """DataParallel checkpoints prefix every key with ``module.`` (interface_v5.py:48)."""
    out = OrderedDict()
    for k, v in sd.items():
        out[k[7:] if k.startswith("module.") else k] = v
    return out


def to_numpy_state_dict(sd):
    out = OrderedDict()
    for k, v in strip_module_prefix(sd).items():
        if hasattr(v, "detach"):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    return out


def check_state_dict(sd, regress_pose: bool = True, arch: str = "v5"):
    """Strict check (the reference uses ``load_state_dict(strict=True)``, interface_v5.py:56)."""
    table = param_table(regress_pose, arch)
    want = {n: s for n, s, _ in table}
    missing = [n for n in want if n not in sd]
    extra = [n for n in sd if n not in want]
    if missing or extra:
        raise KeyError(f"state dict mismatch: missing={missing[:4]} unexpected={extra[:4]}")
    for n, s in want.items():
        if tuple(np.shape(sd[n])) != tuple(s):
            raise ValueError(f"{n}: shape {np.shape(sd[n])} != {s}")


def load_checkpoint(path: str, regress_pose: bool = True, arch: str = "v5"):
    """Read a reference ``.pth`` (torch.save of a DataParallel state_dict)."""
    import torch
    sd = to_numpy_state_dict(torch.load(path, map_location="cpu"))
    check_state_dict(sd, regress_pose, arch)
    return sd


def fold_bn(sd, name: str):
    """Per-output-channel (scale, shift) of eval-mode BatchNorm3d ``name`` (network_v5.py:19,240):
    y = (x - mean) / sqrt(var + eps) * gamma + beta."""
    g = sd[f"{name}.weight"].astype(np.float64)
    b = sd[f"{name}.bias"].astype(np.float64)
    m = sd[f"{name}.running_mean"].astype(np.float64)
    v = sd[f"{name}.running_var"].astype(np.float64)
    scale = g / np.sqrt(v + BN_EPS)
    shift = b - m * scale
    return scale.astype(np.float32), shift.astype(np.float32)
