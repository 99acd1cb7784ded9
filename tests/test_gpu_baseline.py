"""The transformer variant (name: adapose_baseline, train.py:242-244; lib/network_baseline.py:523-669, lib/fusion.py:11-82) on the
B200: the cross-view attention kernel against the oracle, the estimator against boxes of the unmodified reference."""
import os

import numpy as np
import pytest
import torch

from oracle import adapose_oracle as O
from oracle.make_golden import BASELINE_INIT
from rgbmanip_b200 import _lib as L
from rgbmanip_b200 import synth, weights

pytestmark = [pytest.mark.gpu, pytest.mark.filterwarnings("ignore")]


@pytest.fixture(scope="module")
def G():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import importlib
    L.load()
    return importlib.import_module("gpu_util")


def _pack(sd):
    blocks = []
    for b in range(weights.FUSION_DEPTH):
        for f in ("fusion1", "fusion2"):
            for l in range(4):
                nm = f"view_fusion.blocks.{b}.{f}.linears.{l}"
                blocks += [sd[f"{nm}.weight"].reshape(-1), sd[f"{nm}.bias"]]
    dw = np.concatenate([sd["depth_head.0.weight"].reshape(-1), sd["depth_head.0.bias"], sd["depth_head.2.weight"].reshape(32, 64).T.reshape(-1),
                         sd["depth_head.2.bias"], sd["depth_head.4.weight"].reshape(-1), sd["depth_head.4.bias"]])
    return np.concatenate(blocks).astype(np.float32), dw.astype(np.float32)


@pytest.mark.parametrize("gain", [1.0, 6.0])
def test_view_fusion_kernel_matches_oracle(G, gain):
    """adp_view_fusion on random feature maps / pixel subsets, 3 envs (one invalid): tokens after 4 blocks, both depth outputs
    and the bf16 hi/lo planes handed to the pose MLP, fp32 against the oracle (torch fp32 on the CPU)."""
    lib, dev = L.load(), G.DEV
    rng = np.random.default_rng(3)
    sd = weights.init_state_dict(2, arch="baseline", attn_gain=gain)
    B, S, P = 3, 24, 1024
    feat = [rng.standard_normal((B, S * S, 32)).astype(np.float32) * 1.5 for _ in range(2)]
    choose = [rng.integers(0, S * S, (B, P)).astype(np.int32) for _ in range(2)]
    valid = np.array([1, 1, 0], np.uint8)
    bw, dw = _pack(sd)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d = dict(f1=t(feat[0]), f2=t(feat[1]), c1=t(choose[0]), c2=t(choose[1]), valid=t(valid), bw=t(bw), dw=t(dw))
    scratch = torch.zeros((4, B, P, 32), device=dev)
    depth1 = torch.full((B, P), -1.0, device=dev)
    depth2 = torch.full((B, P), -1.0, device=dev)
    fused1 = torch.full((B, P, 32), float("nan"), device=dev)
    fused2 = torch.full((B, P, 32), float("nan"), device=dev)
    xh = torch.zeros((B, P, 96), dtype=torch.bfloat16, device=dev)
    xl = torch.zeros((B, P, 96), dtype=torch.bfloat16, device=dev)
    L.check(lib.adp_view_fusion(L.ptr(d["f1"]), L.ptr(d["f2"]), L.ptr(d["c1"]), L.ptr(d["c2"]), L.ptr(d["valid"]), L.ptr(d["bw"]),
                                L.ptr(d["dw"]), L.ptr(scratch), L.ptr(depth1), L.ptr(depth2), L.ptr(xh), L.ptr(xl), L.ptr(fused1),
                                L.ptr(fused2), B, S, P, weights.FUSION_DEPTH, G.stream()), "view_fusion")
    torch.cuda.synchronize()
    tok = [torch.from_numpy(np.take_along_axis(feat[v], choose[v][:, :, None].astype(np.int64), axis=1)).permute(0, 2, 1) for v in range(2)]
    with torch.no_grad():
        r1, r2 = O.view_fusion(sd, tok[0], tok[1])
        dref = [O._mlp1d(sd, "depth_head", (0, 2, 4), r).squeeze(1).numpy() for r in (r1, r2)]
    r1, r2 = r1.permute(0, 2, 1).numpy(), r2.permute(0, 2, 1).numpy()
    scale = np.abs(r1).max()
    for got, ref in ((fused1, r1), (fused2, r2)):
        np.testing.assert_allclose(got.cpu().numpy()[:2], ref[:2], rtol=0, atol=5e-5 * scale)   # 4 stacked softmax blocks, fp32 both sides
        assert (got[2] == 0).all()
    np.testing.assert_allclose(depth1.cpu().numpy()[:2], dref[0][:2], rtol=0, atol=5e-5)
    np.testing.assert_allclose(depth2.cpu().numpy()[:2], dref[1][:2], rtol=0, atol=5e-5)
    assert (depth1[2] == 0).all()
    planes = (xh.float() + xl.float()).cpu().numpy()
    np.testing.assert_allclose(planes[:2, :, :32], r1[:2], rtol=2e-5, atol=5e-5 * scale)
    assert (planes[:, :, 32:] == 0).all()
    # the last block's view-2 direction is skipped when nobody asks for it: view-1 results must not change
    depth1b = torch.zeros_like(depth1)
    L.check(lib.adp_view_fusion(L.ptr(d["f1"]), L.ptr(d["f2"]), L.ptr(d["c1"]), L.ptr(d["c2"]), L.ptr(d["valid"]), L.ptr(d["bw"]),
                                L.ptr(d["dw"]), L.ptr(scratch), L.ptr(depth1b), None, None, None, None, None, B, S, P,
                                weights.FUSION_DEPTH, G.stream()), "view_fusion")
    torch.cuda.synchronize()
    assert torch.equal(depth1b, depth1)


@pytest.mark.parametrize("precision", ["fp16f8", "bf16x3"])
def test_baseline_estimator_matches_reference(G, golden_dir, precision):
    """AdaPoseEstimator_baseline on the 4 golden scenes (pixel subsets replayed) against the reference's boxes, NOCS, depth and
    rotation (tests/golden/baseline.npz): north_star tolerances 0.5 px / 0.5 deg / 1 mm."""
    from rgbmanip_b200.estimator import AdaPoseEstimator_baseline
    g = np.load(os.path.join(golden_dir, "baseline.npz"))
    sd = weights.init_state_dict(0, arch="baseline", **BASELINE_INIT)
    cfg = {"img_size": 224, "direct_regression": True, "use_depth": True, "load": False, "name": "adapose_baseline"}
    est = AdaPoseEstimator_baseline(None, cfg, None, state_dict=sd, device=G.DEV, max_envs=3, precision=precision, debug=True)
    batch = synth.make_batch(4, seed=9, special=False)
    boxes = est.estimate(*batch.args(), choose=(g["choose1"], g["choose2"]))
    eng = est.estimator
    # 4 envs in chunks of at most 3 = two equal chunks: the last one (envs 2, 3) is still in the engine's buffers
    for slot, e in ((0, 2), (1, 3)):
        assert np.abs(eng.nocs[slot].cpu().numpy() - g["view1_nocs"][e]).max() < 3e-3
        assert np.abs(eng.depth[slot].cpu().numpy() - g["view1_depth"][e]).max() < 2e-3
        assert np.abs(eng.fused1[slot].cpu().numpy().T - g["fused1"][e]).max() < 2e-2 * np.abs(g["fused1"][e]).max()
    worst = np.zeros(4)
    for e in range(4):
        worst = np.maximum(worst, O.parity_errors(boxes[e], g["boxes"][e], batch.K[e], batch.E1[e], min_z=0.5))
    assert worst[0] < 0.5 and worst[1] < 0.5 and worst[2] < 1.0, worst


def test_baseline_sentinel_and_branch_b(G):
    """Empty mask -> sentinel; direct_regression = False (RANSAC + Umeyama on the attention depth) runs through the same class."""
    from rgbmanip_b200.estimator import AdaPoseEstimator_baseline
    sd = weights.init_state_dict(0, regress_pose=False, arch="baseline", **BASELINE_INIT)
    cfg = {"img_size": 224, "direct_regression": False, "use_depth": True, "load": False}
    est = AdaPoseEstimator_baseline(None, cfg, None, state_dict=sd, device=G.DEV, max_envs=4)
    batch = synth.make_batch(4, seed=2, special=False)
    batch.mask1[1] = False
    boxes = est.estimate(*batch.args())
    np.testing.assert_array_equal(boxes[1], O.DEFAULT_BBOX)
    assert np.isfinite(boxes).all()
