"""Generate tests/golden/*.npz by executing the UNMODIFIED reference (/root/reference) on seeded inputs.

Runs only in the build container (the reference does not exist on the GPU box).  The reference module is
imported as-is with MagicMock stand-ins for the simulator packages it imports transitively and no-op
``.cuda()`` shims; it is put in ``eval()`` mode (SURVEY.md finding 0.3-1).
Weights are ``rgbmanip_b200.weights.init_state_dict(seed)`` loaded with ``strict=True`` so every consumer
can regenerate them; inputs come from ``rgbmanip_b200.synth``.

    python oracle/make_golden.py            # writes tests/golden/{e2e,preprocess,units,branch_b,view_ring}.npz
"""
from __future__ import annotations

import logging
import math
import os
import sys
from unittest.mock import MagicMock

import numpy as np
import torch
import torch.nn as nn
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("RGBMANIP_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from rgbmanip_b200 import synth, weights  # noqa: E402

# init gains of the transformer-variant fixtures (weights.init_state_dict docstring); tests regenerate the weights with these
BASELINE_INIT = dict(attn_gain=5.0, depth_bias=0.8)
from oracle.view_ring_oracle import view_ring_script  # noqa: E402


def import_reference():
    for m in ["sapien", "sapien.core", "sapien.core.renderer", "sapien.utils", "mplib", "gym", "gym.spaces",
              "gym.vector", "gym.vector.utils", "gym.vector.utils.shared_memory", "ujson", "matplotlib",
              "matplotlib.pyplot", "trimesh", "transforms3d"]:
        sys.modules.setdefault(m, MagicMock())
    torch.Tensor.cuda = lambda s, *a, **k: s
    nn.Module.cuda = lambda s, *a, **k: s
    from models.pose_estimator.AdaPose import interface_v5
    from models.pose_estimator.AdaPose.lib import align, network_v5, rotation_utils, utils
    return interface_v5, network_v5, rotation_utils, utils, align


def build_reference_estimator(interface_v5, task="drawer", seed=0, direct_regression=True):
    cfg = yaml.safe_load(open(f"{REF}/cfg/pose_estimator/adapose_{task}.yaml"))
    cfg["load"] = False
    cfg["direct_regression"] = direct_regression
    est = interface_v5.AdaPoseEstimator_v5(None, cfg, logging.getLogger("golden"))
    sd = weights.init_state_dict(seed, regress_pose=direct_regression)
    est.estimator.load_state_dict({"module." + k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()},
                                  strict=True)
    est.estimator.eval()
    return est, cfg, sd


def golden_e2e(interface_v5, out_dir):
    torch.manual_seed(0)
    est, cfg, _ = build_reference_estimator(interface_v5)
    batch = synth.make_batch(8, seed=0)
    records = []
    orig_prepare = est.prepare_model_input
    orig_forward = est.estimator.module.forward
    cur = {}

    def prepare(rgb, mask, K, resize_size):
        r = orig_prepare(rgb, mask, K, resize_size)
        cur.setdefault("prep", []).append(r)
        return r

    def forward(*a, **k):
        out = orig_forward(*a, **k)
        cur["pred"] = {kk: v.detach().cpu().numpy() for kk, v in out.items()}
        return out

    est.prepare_model_input = prepare
    est.estimator.module.forward = forward
    feats = {}
    est.estimator.module.img_extractor.register_forward_hook(
        lambda m, i, o: feats.setdefault("f", []).append(o.detach().numpy().copy()))
    est.estimator.module.cost_regularization.register_forward_hook(
        lambda m, i, o: feats.setdefault("logits", []).append(o.detach().numpy().copy()))
    np.random.seed(0)
    boxes = []
    for e in range(len(batch)):
        cur.clear()
        feats.clear()
        b = batch.slice(e, e + 1)
        boxes.append(est.predict(b.K[0], b.rgb1[0], b.mask1[0], b.E1[0], b.rgb2[0], b.mask2[0], b.E2[0]))
        rec = {"valid": "pred" in cur}
        if rec["valid"]:
            (v1, ch1, _, K1), (v2, ch2, _, K2) = cur["prep"]
            rec.update(choose1=ch1, choose2=ch2, K1=K1, K2=K2,
                       view1_rgb_sub=v1.numpy()[:, ::8, ::8], view2_rgb_sub=v2.numpy()[:, ::8, ::8],
                       feat1_sub=feats["f"][0][0, :, ::8, ::8], feat2_sub=feats["f"][1][0, :, ::8, ::8],
                       logits1_sub=feats["logits"][0][0, 0, :, ::8, ::8])
            for k in ("view1_nocs", "view1_depth", "view1_r", "view1_t", "view1_s",
                      "view2_nocs", "view2_depth", "view2_r"):
                rec[k] = cur["pred"][k][0]
        records.append(rec)
    out = {"boxes": np.asarray(boxes), "valid": np.array([r["valid"] for r in records])}
    for e, r in enumerate(records):
        for k, v in r.items():
            if k != "valid":
                out[f"env{e}_{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(out_dir, "e2e.npz"), **out)
    print("e2e: valid", out["valid"], "boxes", out["boxes"].shape)


def golden_e2e_seed1(interface_v5, out_dir):
    """A second, independent end-to-end fixture: other weights (seed 1), other scenes (seed 5), boxes + pixel subsets only."""
    torch.manual_seed(1)
    est, cfg, _ = build_reference_estimator(interface_v5, seed=1)
    batch = synth.make_batch(8, seed=5, special=False)
    chooses = []
    orig_prepare = est.prepare_model_input

    def prepare(rgb, mask, K, resize_size):
        r = orig_prepare(rgb, mask, K, resize_size)
        chooses.append(np.asarray(r[1]))
        return r

    est.prepare_model_input = prepare
    np.random.seed(1)
    boxes = est.estimate(*batch.args())
    ch = np.stack(chooses).reshape(8, 2, -1).astype(np.int32)
    np.savez_compressed(os.path.join(out_dir, "e2e_seed1.npz"), boxes=boxes, choose1=ch[:, 0], choose2=ch[:, 1])
    print("e2e seed 1 boxes", boxes.shape, "finite", np.isfinite(boxes).all())


def golden_branch_b(interface_v5, out_dir):
    """direct_regression=False, use_depth=True -> RANSAC + Umeyama fit (interface_v5.py:322-338)."""
    est, cfg, _ = build_reference_estimator(interface_v5, direct_regression=False)
    batch = synth.make_batch(2, seed=3, special=False)
    np.random.seed(5)
    boxes = est.estimate(*batch.args())
    np.savez_compressed(os.path.join(out_dir, "branch_b.npz"), boxes=boxes)
    print("branch B boxes", boxes.shape, np.isfinite(boxes).all())


def branch_c_unit_case(seed, n_shared=700, noise=0.0015, focal=None, world_scale=1.0):
    """Two views of one rigid object with partly shared surface points: NOCS maps of both views, sampled crop pixels
    (choose + window -> pts2d exactly as interface_v5.py:136-145), intrinsics, extrinsics.  ``focal`` overrides fx = fy (a
    short focal length or a scene in other units (``world_scale``) makes the algebraic epipolar test of utils.py:160-166
    actually reject pairs: with metres and the 440 px focal length it passes everything)."""
    rng = np.random.default_rng(seed)
    P, S = 1024, 224
    K = synth.intrinsics()
    if focal is not None:
        K[0, 0] = K[1, 1] = focal
    E1, E2 = synth._camera_pair(rng)
    target = -E1[:3, :3].T @ E1[:3, 3] + E1[2, :3] * 0.7            # a point 0.7 m in front of camera 1
    Rw = np.linalg.qr(rng.standard_normal((3, 3)))[0]
    size = rng.uniform(0.15, 0.3)
    obj = rng.uniform(-0.5, 0.5, (2 * P - n_shared, 3))
    ids1 = np.arange(P)
    ids2 = np.concatenate([rng.permutation(P)[:n_shared], np.arange(P, 2 * P - n_shared)])
    rng.shuffle(ids2)
    out = {}
    for v, (ids, E) in enumerate(((ids1, E1), (ids2, E2)), 1):
        world = (size * obj[ids]) @ Rw.T + target
        cam = world @ E[:3, :3].T + E[:3, 3]
        uv = (cam @ K.T)
        uv = uv[:, :2] / uv[:, 2:3]
        crop = 40 * int(rng.integers(4, 9))                           # window sizes of get_bbox
        rmin = int(np.clip(np.median(uv[:, 1]) - crop / 2, 0, 480 - crop))
        cmin = int(np.clip(np.median(uv[:, 0]) - crop / 2, 0, 640 - crop))
        ratio = S / crop
        cx = np.clip(np.round((uv[:, 0] - cmin) * ratio), 0, S - 1).astype(np.int64)
        cy = np.clip(np.round((uv[:, 1] - rmin) * ratio), 0, S - 1).astype(np.int64)
        out[f"choose{v}"] = (cy * S + cx).astype(np.int32)
        out[f"win{v}"] = np.array([rmin, rmin + crop, cmin, cmin + crop], np.int32)
        out[f"nocs{v}"] = (obj[ids] + rng.normal(0, noise, (P, 3))).astype(np.float32)
    if world_scale != 1.0:          # the same images seen in other length units: only the translations change
        E1, E2 = E1.copy(), E2.copy()
        E1[:3, 3] *= world_scale
        E2[:3, 3] *= world_scale
    out.update(K=K, E1=E1, E2=E2)
    return out


def golden_branch_c_units(utils, align, out_dir):
    """utils.depth_estimation_from_nocs_matches + align.estimatePnPRansac of the reference on synthetic two-view cases."""
    import contextlib
    import io
    from oracle import adapose_oracle as orc
    out = {}
    cases = [dict(seed=11), dict(seed=12, n_shared=300, noise=0.004), dict(seed=13, focal=6.0), dict(seed=14, n_shared=1024, noise=0.0),
             dict(seed=15, world_scale=400.0)]
    for ci, kw in enumerate(cases):
        c = branch_c_unit_case(**kw)
        pts1 = orc.prepare_pts2d(c["choose1"].astype(np.int64), *c["win1"][:3])
        pts2 = orc.prepare_pts2d(c["choose2"].astype(np.int64), *c["win2"][:3])
        P1, P2 = np.eye(4), np.eye(4)
        P1[:3], P2[:3] = c["K"] @ c["E1"][:3], c["K"] @ c["E2"][:3]
        with contextlib.redirect_stdout(io.StringIO()):
            ls, rs, lp, rp = utils.depth_estimation_from_nocs_matches(pts1, c["nocs1"], P1, c["E1"], pts2, c["nocs2"], P2, c["E2"], c["K"])
            if np.isfinite(ls):
                ok, size, R, t, _ = align.estimatePnPRansac(c["nocs1"].astype(np.float32), pts1.astype(np.float32), ls, c["K"])
            else:
                ok, R, t = False, np.full((3, 3), np.nan), np.full((3, 1), np.nan)
        for k, v in c.items():
            out[f"case{ci}_{k}"] = v
        out.update({f"case{ci}_pts2d1": pts1, f"case{ci}_pts2d2": pts2, f"case{ci}_left_scale": ls, f"case{ci}_right_scale": rs,
                    f"case{ci}_left_pts": lp, f"case{ci}_right_pts": rp, f"case{ci}_pnp_ok": ok, f"case{ci}_pnp_R": R,
                    f"case{ci}_pnp_t": t})
        print(f"branch C unit case {ci}: matches {len(lp)}, scales {ls:.6f} {rs:.6f}, pnp ok {ok}")
    out["n_cases"] = len(cases)
    np.savez_compressed(os.path.join(out_dir, "branch_c_units.npz"), **out)


def golden_branch_c(interface_v5, out_dir):
    """direct_regression=False, use_depth=False -> NOCS matching + triangulation + cv2 PnP (interface_v5.py:339-349)."""
    import contextlib
    import io
    est, cfg, _ = build_reference_estimator(interface_v5, direct_regression=False)
    est.cfg["use_depth"] = False
    batch = synth.make_batch(4, seed=3, special=False)
    rec = {"match": [], "pnp": [], "prep": [], "pred": []}
    orig_match, orig_pnp = interface_v5.depth_estimation_from_nocs_matches, interface_v5.estimatePnPRansac
    orig_prepare, orig_forward = est.prepare_model_input, est.estimator.module.forward

    def match(*a):
        r = orig_match(*a)
        rec["match"].append(r)
        return r

    def pnp(*a):
        r = orig_pnp(*a)
        rec["pnp"].append(r)
        return r

    def prepare(rgb, mask, K, resize_size):
        r = orig_prepare(rgb, mask, K, resize_size)
        rec["prep"].append(r)
        return r

    def forward(*a, **k):
        o = orig_forward(*a, **k)
        rec["pred"].append({kk: v.detach().cpu().numpy() for kk, v in o.items()})
        return o

    interface_v5.depth_estimation_from_nocs_matches, interface_v5.estimatePnPRansac = match, pnp
    est.prepare_model_input, est.estimator.module.forward = prepare, forward
    np.random.seed(7)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            boxes = est.estimate(*batch.args())
    finally:
        interface_v5.depth_estimation_from_nocs_matches, interface_v5.estimatePnPRansac = orig_match, orig_pnp
    n = len(batch)
    out = dict(boxes=boxes,
               choose1=np.stack([rec["prep"][2 * e][1] for e in range(n)]).astype(np.int32),
               choose2=np.stack([rec["prep"][2 * e + 1][1] for e in range(n)]).astype(np.int32),
               pts2d1=np.stack([rec["prep"][2 * e][2] for e in range(n)]),
               pts2d2=np.stack([rec["prep"][2 * e + 1][2] for e in range(n)]),
               nocs1=np.stack([rec["pred"][e]["view1_nocs"][0] for e in range(n)]),
               nocs2=np.stack([rec["pred"][e]["view2_nocs"][0] for e in range(n)]),
               left_scale=np.array([rec["match"][e][0] for e in range(n)]),
               right_scale=np.array([rec["match"][e][1] for e in range(n)]),
               n_match=np.array([len(rec["match"][e][2]) for e in range(n)]),
               pnp_ok=np.array([bool(rec["pnp"][e][0]) for e in range(n)]),
               pnp_R=np.stack([rec["pnp"][e][2] for e in range(n)]), pnp_t=np.stack([np.asarray(rec["pnp"][e][3]).flatten() for e in range(n)]))
    for e in range(n):
        out[f"env{e}_left_pts"] = rec["match"][e][2]
        out[f"env{e}_right_pts"] = rec["match"][e][3]
    np.savez_compressed(os.path.join(out_dir, "branch_c.npz"), **out)
    print("branch C boxes", boxes.shape, "finite", np.isfinite(boxes).all(), "matches", out["n_match"], "scales", out["left_scale"])


def golden_baseline(out_dir):
    """The transformer variant (train.py:242-244: name = adapose_baseline -> interface_baseline.AdaPoseEstimator_baseline, the v5
    interface around StereoPoseNet_with_depth_baseline): 4 envs end to end, plus the attention statistics the init gains aim at."""
    from models.pose_estimator.AdaPose import interface_baseline
    cfg = yaml.safe_load(open(f"{REF}/cfg/pose_estimator/adapose_drawer.yaml"))
    cfg.update(load=False, name="adapose_baseline")
    torch.manual_seed(0)
    est = interface_baseline.AdaPoseEstimator_baseline(None, cfg, logging.getLogger("golden"))
    sd = weights.init_state_dict(0, arch="baseline", **BASELINE_INIT)
    est.estimator.load_state_dict({"module." + k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    est.estimator.eval()
    batch = synth.make_batch(4, seed=9, special=False)
    rec = {"prep": [], "pred": [], "fused": [], "attn": []}
    orig_prepare, orig_forward = est.prepare_model_input, est.estimator.module.forward

    def prepare(rgb, mask, K, resize_size):
        r = orig_prepare(rgb, mask, K, resize_size)
        rec["prep"].append(r)
        return r

    def forward(*a, **k):
        o = orig_forward(*a, **k)
        rec["pred"].append({kk: v.detach().cpu().numpy() for kk, v in o.items()})
        return o

    est.prepare_model_input, est.estimator.module.forward = prepare, forward
    est.estimator.module.view_fusion.register_forward_hook(lambda m, i, o: rec["fused"].append([t.detach().numpy().copy() for t in o]))
    blk0 = est.estimator.module.view_fusion.blocks[0].fusion1
    est.estimator.module.view_fusion.blocks[0].register_forward_hook(lambda m, i, o: rec["attn"].append(blk0.attn.detach().numpy().copy()))
    np.random.seed(11)
    boxes = est.estimate(*batch.args())
    n = len(batch)
    p = rec["attn"][0][0]                                  # [heads, N, N] of env 0, block 0
    ent = -(p * np.log(np.maximum(p, 1e-30))).sum(-1)
    print("baseline: attention entropy (block 0, env 0) mean %.3f of max %.3f; row max prob mean %.3f" % (ent.mean(), math.log(p.shape[-1]), p.max(-1).mean()))
    out = dict(boxes=boxes,
               choose1=np.stack([rec["prep"][2 * e][1] for e in range(n)]).astype(np.int32),
               choose2=np.stack([rec["prep"][2 * e + 1][1] for e in range(n)]).astype(np.int32),
               fused1=np.stack([rec["fused"][e][0][0] for e in range(n)]), fused2=np.stack([rec["fused"][e][1][0] for e in range(n)]),
               attn_entropy=ent.astype(np.float32))
    for k in ("view1_nocs", "view2_nocs", "view1_depth", "view2_depth", "view1_r", "view2_r"):
        out[k] = np.stack([rec["pred"][e][k][0] for e in range(n)])
    np.savez_compressed(os.path.join(out_dir, "baseline.npz"), **out)
    print("baseline boxes", boxes.shape, "finite", np.isfinite(boxes).all(), "depth range", out["view1_depth"].min(), out["view1_depth"].max(),
          "fused range", np.abs(out["fused1"]).max())


def golden_view_ring(out_dir):
    """Caller-side queues of the RL controller (models/controller/rl_pose.py:85-97,118-150,189-223), executed unmodified."""
    for m in ["tensorboard", "torch.utils.tensorboard", "ipdb", "open3d", "sapien.utils.viewer"]:
        sys.modules.setdefault(m, MagicMock())
    from models.controller import rl_pose
    calls = []

    class Stub:
        cfg = {"task_name": "one_drawer_cabinet"}

        def estimate(self, K, rgb1, m1, E1, rgb2, m2, E2):
            calls.append(dict(K=K[:, 0, 0].copy(), rgb1=rgb1[:, 0, 0, 0].copy(), rgb2=rgb2[:, 0, 0, 0].copy(), m1=m1.sum((1, 2)),
                              m2=m2.sum((1, 2)), E1=E1[:, 0, 0].copy(), E2=E2[:, 0, 0].copy()))
            return np.arange(K.shape[0] * 24, dtype=np.float64).reshape(-1, 8, 3)

    ci = object.__new__(rl_pose.ControlInterface)
    ci.num_envs, ci.max_steps, ci.estimator = 3, 5, Stub()
    ci.accumulate_steps = 0
    ci.reset_queue()
    rec = {}
    for t, (color, mask, K, E, pose) in enumerate(view_ring_script()):
        image = {"camera0": {"Color": color, "Mask": mask, "Intrinsic": K, "Extrinsic": E}}
        ci.add_view(image, pose)
        ci.accumulate_steps += 1                      # as reset_robot / step do (rl_pose.py:116,...)
        box = ci.get_estimation()
        c = calls[-1]
        rec[f"s{t}_available"] = ci.available.copy(); rec[f"s{t}_available_num"] = ci.available_num.copy()
        rec[f"s{t}_bbox_queue"] = ci.bbox_queue.copy(); rec[f"s{t}_pose_queue"] = ci.pose_queue.copy()
        for k, v in c.items():
            rec[f"s{t}_{k}"] = v
        rec[f"s{t}_box"] = box
    ci.estimator.cfg = {"task_name": "mugs"}
    rec["mug_box"] = ci.get_estimation()
    np.savez_compressed(os.path.join(out_dir, "view_ring.npz"), **rec)
    print("view ring golden:", len(rec), "arrays")


def golden_actor(out_dir):
    """Observation of the controller (rl_pose.py:173-187) and the actor forward (module.py:24-34,89-91), reference code
    executed unmodified: ControlInterface.get_observation on scripted queues, ActorCritic.act_inference on it."""
    for m in ["tensorboard", "torch.utils.tensorboard", "ipdb", "open3d", "sapien.utils.viewer"]:
        sys.modules.setdefault(m, MagicMock())
    from algo.ppo.ppo.module import ActorCritic
    from models.controller import rl_pose
    pol = yaml.safe_load(open(f"{REF}/cfg/controller/rl.yaml"))["policy"]
    torch.manual_seed(3)
    ac = ActorCritic((60,), (75,), (12,), 0.6, pol, asymmetric=False)
    with torch.no_grad():                      # move the last layer off its 0.01-gain init so that the output is informative
        ac.actor[-1].weight.mul_(30.0)
        for m in ac.actor:
            if isinstance(m, nn.Linear):
                m.bias.normal_(0, 0.1)
    ci = object.__new__(rl_pose.ControlInterface)
    ci.num_envs, ci.max_steps = 3, 5
    ci.accumulate_steps = 0
    ci.reset_queue()
    rec = {k: v.detach().numpy() for k, v in ac.state_dict().items() if k.startswith("actor.")}
    for t, (color, mask, K, E, pose) in enumerate(view_ring_script()):
        ci.add_view({"camera0": {"Color": color, "Mask": mask, "Intrinsic": K, "Extrinsic": E}}, pose * 0.1)
        ci.accumulate_steps += 1
        if t >= 5:            # the controller resets its queues before the step counter wraps; stay inside one episode
            break
        obs = ci.get_observation()
        rec[f"s{t}_obs"] = obs.numpy()
        rec[f"s{t}_act"] = ac.act_inference(obs).numpy()
    np.savez_compressed(os.path.join(out_dir, "actor.npz"), **rec)
    print("actor golden:", len(rec), "arrays; obs", obs.shape)


def golden_preprocess(interface_v5, utils, out_dir):
    est, cfg, _ = build_reference_estimator(interface_v5)
    rng = np.random.default_rng(7)
    K = synth.intrinsics()
    out = {}
    n = 0
    for i in range(40):
        dt = np.float64 if i % 3 == 0 else np.float32
        rgb = rng.random((480, 640, 3)).astype(dt)
        cx, cy = rng.uniform(0, 640), rng.uniform(0, 480)
        ax, ay = rng.uniform(2, 260), rng.uniform(2, 200)
        mask = synth._ellipse_mask(cx, cy, ax, ay)
        if i % 5 == 0:
            mask = mask.astype(np.float64)
        if mask.sum() == 0:
            continue
        np.random.seed(100 + i)
        v, ch, pts, Kp = est.prepare_model_input(rgb, mask, K, 224)
        ys, xs = np.nonzero(mask)
        win = utils.get_bbox([ys.min(), xs.min(), ys.max(), xs.max()])
        out[f"c{n}_params"] = np.array([cx, cy, ax, ay, i], np.float64)
        out[f"c{n}_window"] = np.array(win, np.int64)
        out[f"c{n}_choose"] = ch
        out[f"c{n}_pts2d_sub"] = pts[::16]
        out[f"c{n}_K"] = Kp
        out[f"c{n}_rgb_sub"] = v.numpy()[:, 3::8, 5::8].astype(np.float32)
        out[f"c{n}_rgb_sum"] = np.array([float(v.double().sum()), float(v.double().abs().sum())])
        n += 1
    out["count"] = np.array(n)
    # nearest-neighbour index tables for every window size the reference can produce
    import cv2
    for ws in range(40, 441, 40):
        ramp = np.arange(ws, dtype=np.float32)[None].repeat(ws, 0)
        out[f"nn_{ws}"] = cv2.resize(ramp, (224, 224), interpolation=cv2.INTER_NEAREST)[0].astype(np.int64)
    np.savez_compressed(os.path.join(out_dir, "preprocess.npz"), **out)
    print("preprocess cases:", n)


def unit_inputs():
    """Seeded inputs of the unit goldens (numpy PCG64: identical on every machine; not stored)."""
    rng = np.random.default_rng(11)
    f32 = lambda *s: rng.standard_normal(s, dtype=np.float32)
    return {"warp_src": f32(1, 4, 224, 224), "cr_in": f32(1, 32, 8, 16, 24), "psp_in": f32(1, 3, 64, 96),
            "r6": f32(5, 6)}


def golden_units(network_v5, rotation_utils, utils, align, out_dir):
    ui = {k: torch.from_numpy(v) for k, v in unit_inputs().items()}
    out = {}
    net = network_v5.StereoPoseNet_with_depth(n_cat=1, nv_pts=1024, regress_pose=True)
    sd = weights.init_state_dict(1)
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    net.eval()
    with torch.no_grad():
        # homo_warping on a small map with a realistic projection pair
        b = synth.make_batch(1, seed=21, special=False)
        from oracle import adapose_oracle as O
        _, _, _, K1 = O.prepare_model_input(b.rgb1[0], b.mask1[0], b.K[0])
        _, _, _, K2 = O.prepare_model_input(b.rgb2[0], b.mask2[0], b.K[0])
        P1 = torch.from_numpy(O.projection(K1, b.E1[0])).float()[None]
        P2 = torch.from_numpy(O.projection(K2, b.E2[0])).float()[None]
        src = ui["warp_src"]
        dv = torch.from_numpy(O.depth_hypotheses())[None]
        warped = net.homo_warping(src, P2, P1, dv)
        out.update(warp_P1=P1.numpy(), warp_P2=P2.numpy(), warp_out_sub=warped.numpy()[0, :, :, ::4, ::4])
        # CostRegNet on a small volume
        out.update(cr_out=net.cost_regularization(ui["cr_in"]).numpy())
        # backbone on a small image
        out.update(psp_out=net.img_extractor(ui["psp_in"]).numpy())
        # 6-D -> rotation
        r6 = ui["r6"]
        out.update(r6_mat=rotation_utils.Ortho6d2Mat(r6[:, :3].contiguous(), r6[:, 3:].contiguous()).numpy())
    # fit functions
    rng = np.random.default_rng(5)
    nocs = (rng.random((1024, 3)).astype(np.float32) - 0.5)
    Rt = np.linalg.qr(rng.standard_normal((3, 3)))[0]
    if np.linalg.det(Rt) < 0:
        Rt[:, 0] *= -1
    cam = (0.23 * (Rt @ nocs.T.astype(np.float64)).T + np.array([0.05, -0.02, 0.8])
           + 0.004 * rng.standard_normal((1024, 3)))
    out.update(fit_nocs=nocs, fit_cam=cam, fit_scale=np.array(utils.compute_scale(cam, nocs)))
    depth = cam[:, 2].astype(np.float32)
    choose = np.sort(rng.choice(224 * 224, 1024, replace=False))
    Kp = np.array([[800.0, 0, 100.5], [0, 800.0, 120.25], [0, 0, 1]])
    t, s = utils.compute_scale_and_translation(depth, nocs, choose, Kp, 224, Rt.astype(np.float32))
    out.update(fit2_depth=depth, fit2_choose=choose, fit2_K=Kp, fit2_R=Rt.astype(np.float32), fit2_t=t,
               fit2_s=np.array(s))
    np.random.seed(9)
    sc, R, tr, T = align.estimateSimilarityTransform(nocs, cam)
    out.update(um_scale=np.array(sc), um_R=R, um_t=tr, um_T=T)
    size = np.array([0.2, 0.4, 0.6])
    out.update(bbox_size=size, bbox=utils.get_3d_bbox(size))
    np.savez_compressed(os.path.join(out_dir, "units.npz"), **out)
    print("units written")


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    interface_v5, network_v5, rotation_utils, utils, align = import_reference()
    torch.set_num_threads(os.cpu_count())
    what = sys.argv[1:] or ["units", "preprocess", "e2e", "branch_b", "view_ring", "actor", "e2e_seed1", "branch_c_units", "branch_c", "baseline"]
    if "units" in what:
        golden_units(network_v5, rotation_utils, utils, align, out_dir)
    if "preprocess" in what:
        golden_preprocess(interface_v5, utils, out_dir)
    if "e2e" in what:
        golden_e2e(interface_v5, out_dir)
    if "branch_b" in what:
        golden_branch_b(interface_v5, out_dir)
    if "branch_c_units" in what:
        golden_branch_c_units(utils, align, out_dir)
    if "branch_c" in what:
        golden_branch_c(interface_v5, out_dir)
    if "baseline" in what:
        golden_baseline(out_dir)
    if "e2e_seed1" in what:
        golden_e2e_seed1(interface_v5, out_dir)
    if "view_ring" in what:
        golden_view_ring(out_dir)
    if "actor" in what:
        golden_actor(out_dir)


if __name__ == "__main__":
    main()
