"""End-to-end parity on the B200: AdaPoseEstimator_v5.estimate (rgbmanip_b200, CUDA through the C ABI) against
 (i) the golden vectors produced by the reference itself (tests/golden/e2e.npz, see oracle/make_golden.py) and
 (ii) the CPU oracle, with the tolerances of BASELINE.json's north_star: keypoints 0.5 px, rotation 0.5 deg,
 translation 1 mm.  The reference's own random pixel subset is replayed (choose=...) so that both sides decode the
 same 1024 pixels; the device sampler is covered in test_gpu_stages.py."""
import os

import numpy as np
import pytest
import torch

from oracle import adapose_oracle as O
from rgbmanip_b200 import synth, weights

pytestmark = [pytest.mark.gpu, pytest.mark.filterwarnings("ignore")]

TOL_PX, TOL_DEG, TOL_MM = 0.5, 0.5, 1.0
# Keypoints = the 8 box corners + centre projected into view 1.  With random-init weights the boxes are ~1 m wide and
# reach to within centimetres of the camera plane, where d(pixel)/d(metre) = f/z diverges; corners closer than 0.5 m
# (nearer than any object the on-hand camera looks at: depth planes start at 0.1 m, handles sit at 0.4-1.2 m) are
# compared in millimetres only (they still count for the 1 mm corner bound).
MIN_Z = 0.5


@pytest.fixture(scope="module")
def golden(golden_dir):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return np.load(os.path.join(golden_dir, "e2e.npz"))


def _golden_choose(g, n):
    c1 = np.zeros((n, 1024), np.int32)
    c2 = np.zeros((n, 1024), np.int32)
    for e in range(n):
        if g["valid"][e]:
            c1[e], c2[e] = g[f"env{e}_choose1"], g[f"env{e}_choose2"]
    return c1, c2


def _make(cfg_extra=None, **kw):
    from rgbmanip_b200.estimator import AdaPoseEstimator_v5
    cfg = {"name": "adapose_v5", "task_name": "one_drawer_cabinet", "load": False, "img_size": 224, "use_depth": True,
           "n_pts": 1024, "direct_regression": True, "real_world": False}
    cfg.update(cfg_extra or {})
    return AdaPoseEstimator_v5(None, cfg, None, state_dict=weights.init_state_dict(0), **kw)


@pytest.mark.parametrize("precision", ["fp16f8", "fp16x2", "bf16x3"])
@pytest.mark.parametrize("max_envs", [8, 3])
def test_estimate_matches_reference_golden(golden, max_envs, precision):
    g = golden
    batch = synth.make_batch(8, seed=0)
    est = _make(max_envs=max_envs, precision=precision, debug=True)
    boxes = est.estimate(*batch.args(), choose=_golden_choose(g, 8))
    assert boxes.shape == (8, 8, 3) and boxes.dtype == np.float64
    worst = np.zeros(3)
    compared = total = 0
    for e in range(8):
        if not g["valid"][e]:
            np.testing.assert_array_equal(boxes[e], O.DEFAULT_BBOX)      # sentinel: bit exact
            continue
        px, deg, mm, cmm = O.parity_errors(boxes[e], g["boxes"][e], batch.K[e], batch.E1[e], min_z=MIN_Z)
        worst = np.maximum(worst, (px, deg, mm))
        assert px < TOL_PX and deg < TOL_DEG and mm < TOL_MM and cmm < TOL_MM, (e, px, deg, mm, cmm)
        k, t = O.keypoints_compared(boxes[e], g["boxes"][e], batch.K[e], batch.E1[e], min_z=MIN_Z)
        compared, total = compared + k, total + t
    # the MIN_Z exemption (corners nearer than 0.5 m to the camera plane are held to 1 mm instead of 0.5 px) must stay the
    # exception: every keypoint is held to the millimetre bound, and most of them to the pixel bound as well
    print(f"{precision} worst (px, deg, mm):", worst, f"keypoints in the pixel metric: {compared} of {total}")
    assert compared >= 0.6 * total, (compared, total)
    est.estimator.close()


@pytest.mark.parametrize("precision", ["fp16f8", "fp16x2", "bf16x3"])
def test_estimate_matches_second_reference_fixture(golden_dir, precision):
    """Independent fixture: other random weights (seed 1), other scenes (seed 5) -- tests/golden/e2e_seed1.npz."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from rgbmanip_b200.estimator import AdaPoseEstimator_v5
    g = np.load(os.path.join(golden_dir, "e2e_seed1.npz"))
    cfg = {"name": "adapose_v5", "task_name": "one_drawer_cabinet", "load": False, "img_size": 224, "use_depth": True,
           "n_pts": 1024, "direct_regression": True, "real_world": False}
    batch = synth.make_batch(8, seed=5, special=False)
    est = AdaPoseEstimator_v5(None, cfg, None, state_dict=weights.init_state_dict(1), max_envs=8, precision=precision)
    boxes = est.estimate(*batch.args(), choose=(g["choose1"], g["choose2"]))
    worst = np.zeros(4)
    for e in range(8):
        worst = np.maximum(worst, O.parity_errors(boxes[e], g["boxes"][e], batch.K[e], batch.E1[e], min_z=MIN_Z))
    print(f"{precision} second fixture worst (px, deg, mm, corner-mm):", worst)
    assert worst[0] < TOL_PX and worst[1] < TOL_DEG and worst[2] < TOL_MM and worst[3] < TOL_MM, worst
    est.estimator.close()


@pytest.mark.parametrize("precision", ["fp16f8", "fp16x2", "bf16x3"])
def test_network_outputs_match_reference_golden(golden, precision):
    """NOCS / depth / rotation of the last processed chunk against the reference's own tensors."""
    g = golden
    batch = synth.make_batch(8, seed=0)
    est = _make(max_envs=8, precision=precision)
    est.estimate(*batch.args(), choose=_golden_choose(g, 8))
    eng = est.estimator
    for e in range(8):
        if not g["valid"][e]:
            continue
        # NOCS coordinates (unit cube): fp16f8 carries the e4m3 rounding of the low-order weight term on top of fp16x2
        assert float(np.abs(eng.nocs[e].cpu().numpy() - g[f"env{e}_view1_nocs"]).max()) < (3e-3 if precision == "fp16f8" else 2e-3)
        d = np.abs(eng.depth[e].cpu().numpy() - g[f"env{e}_view1_depth"])
        assert d.mean() < 1e-3 and d.max() < 8e-3           # metres; per-pixel soft-argmax through the bf16 U-Net
        assert O.rotation_angle_deg(eng.R[e].cpu().numpy().reshape(3, 3), g[f"env{e}_view1_r"]) < 0.1
        f = eng.feat[e].cpu().numpy().transpose(2, 0, 1)[:, ::8, ::8]
        np.testing.assert_allclose(f, g[f"env{e}_feat1_sub"], rtol=2e-3, atol=5e-3)
    eng.close()


def test_single_pass_bf16_error_is_bounded_and_recorded(golden):
    """Plain bf16 operands (one MMA pass) cannot meet 1 mm through a 36-layer BN-free backbone (DESIGN.md,
    'precision policy'); the mode exists for throughput and its error is bounded here, not hidden."""
    g = golden
    batch = synth.make_batch(8, seed=0)
    est = _make(max_envs=8, precision="bf16")
    boxes = est.estimate(*batch.args(), choose=_golden_choose(g, 8))
    errs = [O.parity_errors(boxes[e], g["boxes"][e], batch.K[e], batch.E1[e]) for e in range(8) if g["valid"][e]]
    errs = np.array(errs)
    print("bf16 single-pass worst (px, deg, mm, corner-mm):", errs.max(0))
    assert errs[:, 1].max() < TOL_DEG            # rotation stays inside the tolerance
    assert errs[:, 3].max() < 80.0               # corners: a few cm on ~1 m boxes (1-2 % of the box size)
    est.estimator.close()


def test_predict_and_dtypes(golden):
    """predict() = one env; float64 frames + float64 0/1 masks (what rl_pose.py passes) give the same box as float32/bool."""
    g = golden
    batch = synth.make_batch(8, seed=0)
    est = _make(max_envs=2)
    e = 1
    ch = (g[f"env{e}_choose1"][None].astype(np.int32), g[f"env{e}_choose2"][None].astype(np.int32))
    sl = batch.slice(e, e + 1)
    a = est.estimate(*sl.args(), choose=ch)[0]
    b = est.estimate(sl.K, sl.rgb1.astype(np.float64), sl.mask1.astype(np.float64), sl.E1,
                     sl.rgb2.astype(np.float64), sl.mask2.astype(np.float64), sl.E2, choose=ch)[0]
    px, deg, mm, cmm = O.parity_errors(a, b, batch.K[e], batch.E1[e])
    assert px < 0.4 and mm < 0.3, (px, deg, mm)      # fp64 vs fp32 frames: ulp-level crop differences through the fp16 backbone
    c = est.predict(sl.K[0], sl.rgb1[0], sl.mask1[0], sl.E1[0], sl.rgb2[0], sl.mask2[0], sl.E2[0])
    assert c.shape == (8, 3) and np.isfinite(c).all()
    est.estimator.close()


def test_checkpoint_path_loading_equals_state_dict(tmp_path):
    """cfg['load'] = True + cfg['checkpoint_path'] (interface_v5.py:55-56): a DataParallel-prefixed .pth gives the same boxes."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from rgbmanip_b200.estimator import AdaPoseEstimator_v5
    sd = weights.init_state_dict(0)
    path = str(tmp_path / "adapose_drawer.pth")
    torch.save({"module." + k: torch.from_numpy(np.asarray(v).copy()) for k, v in sd.items()}, path)
    batch = synth.make_batch(2, seed=6, special=False)
    a = _make(max_envs=2)
    cfg = dict(a.cfg, load=True, checkpoint_path=path)
    b = AdaPoseEstimator_v5(None, cfg, None, max_envs=2)
    np.testing.assert_allclose(a.estimate(*batch.args()), b.estimate(*batch.args()), rtol=0, atol=1e-5)
    a.estimator.close(); b.estimator.close()
    with pytest.raises(FileNotFoundError):
        AdaPoseEstimator_v5(None, dict(cfg, checkpoint_path=str(tmp_path / "missing.pth")), None, max_envs=2)


def test_empty_batch_and_all_blind_environments():
    """num_envs = 0 returns an empty [0,8,3]; environments whose masks are empty in either view get the sentinel box
    (interface_v5.py:232-241,256-257) bit for bit, whatever the other environments hold."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    est = _make(max_envs=4)
    z = lambda *s, dt=np.float32: np.zeros(s, dt)
    out = est.estimate(z(0, 3, 3, dt=np.float64), z(0, 480, 640, 3), z(0, 480, 640, dt=bool), z(0, 4, 4, dt=np.float64),
                       z(0, 480, 640, 3), z(0, 480, 640, dt=bool), z(0, 4, 4, dt=np.float64))
    assert out.shape == (0, 8, 3) and out.dtype == np.float64
    b = synth.make_batch(5, seed=9, special=False)
    m1, m2 = b.mask1.copy(), b.mask2.copy()
    m1[0] = 0                  # blind in view 1
    m2[3] = 0                  # blind in view 2
    m1[4] = 0; m2[4] = 0       # blind in both (and the partial second chunk)
    out = est.estimate(b.K, b.rgb1, m1, b.E1, b.rgb2, m2, b.E2)
    for e in (0, 3, 4):
        np.testing.assert_array_equal(out[e], O.DEFAULT_BBOX)
    for e in (1, 2):
        assert np.isfinite(out[e]).all() and not np.array_equal(out[e], O.DEFAULT_BBOX)
    blind = est.estimate(b.K, b.rgb1, np.zeros_like(m1), b.E1, b.rgb2, np.zeros_like(m2), b.E2)
    np.testing.assert_array_equal(blind, np.broadcast_to(O.DEFAULT_BBOX, (5, 8, 3)))
    est.estimator.close()


def test_fp16_range_overflow_is_reported_not_silently_turned_into_sentinels():
    """fp16x2 stores activations in IEEE half; a checkpoint whose activations leave that range must fail loudly on the
    first batch (and run under bf16x3)."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from rgbmanip_b200._lib import AdpError
    from rgbmanip_b200.estimator import AdaPoseEstimator_v5
    sd = dict(weights.init_state_dict(0))
    sd["img_extractor.feats.conv1.weight"] = sd["img_extractor.feats.conv1.weight"] * 3e4      # blows the stem past 65504
    cfg = {"name": "adapose_v5", "task_name": "one_drawer_cabinet", "load": False, "img_size": 224, "use_depth": True,
           "n_pts": 1024, "direct_regression": True, "real_world": False}
    batch = synth.make_batch(2, seed=4, special=False)
    est = AdaPoseEstimator_v5(None, cfg, None, state_dict=sd, max_envs=2, precision="fp16x2")
    with pytest.raises(AdpError, match="fp16 range"):
        est.estimate(*batch.args())
    est.estimator.close()
    est = AdaPoseEstimator_v5(None, cfg, None, state_dict=sd, max_envs=2, precision="bf16x3")
    assert est.estimate(*batch.args()).shape == (2, 8, 3)          # bf16's range takes it
    est.estimator.close()


def test_four_task_configs_share_the_path():
    """cabinet / drawer / mug / pot yamls differ only in task_name and checkpoint path (cfg/pose_estimator/adapose_*.yaml)."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    batch = synth.make_batch(2, seed=4, special=False)
    outs = []
    for task in ("one_door_cabinet", "one_drawer_cabinet", "mugs", "pots"):
        est = _make({"task_name": task}, max_envs=2)
        assert est.cfg["task_name"] == task
        outs.append(est.estimate(*batch.args()))
        est.estimator.close()
    for o in outs[1:]:
        np.testing.assert_allclose(o, outs[0], rtol=0, atol=1e-5)   # same weights/seed; only atomicAdd order differs


@pytest.mark.parametrize("precision,mm_bound,cmm_bound", [("bf16x3", 5.0, 8.0), ("fp16x2", 10.0, 20.0)])
def test_branch_b_matches_reference_golden(golden_dir, precision, mm_bound, cmm_bound):
    """direct_regression=False, use_depth=True (RANSAC + Umeyama fit on the device) against the reference's boxes.
    The reference consumes the global numpy stream (pixel subsets and RANSAC draws interleaved, early exits included);
    the CPU oracle replays it here to recover the exact draws, which are then handed to the device path."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from rgbmanip_b200.estimator import AdaPoseEstimator_v5
    g = np.load(os.path.join(golden_dir, "branch_b.npz"))
    sd = weights.init_state_dict(0, regress_pose=False)
    cfg = {"img_size": 224, "direct_regression": False, "use_depth": True, "load": False}
    batch = synth.make_batch(2, seed=3, special=False)

    class Recorder:                      # np.random facade that logs the randint draws
        def __init__(self):
            self.draws = []
        def shuffle(self, a):
            np.random.shuffle(a)
        def randint(self, n, size):
            r = np.random.randint(n, size=size)
            self.draws.append(r)
            return r

    np.random.seed(5)
    chooses, tables = [], []
    for e in range(2):
        rec = Recorder()
        d = {}
        O.predict(sd, cfg, batch.K[e], batch.rgb1[e], batch.mask1[e], batch.E1[e], batch.rgb2[e], batch.mask2[e], batch.E2[e],
                  rng=rec, details=d)
        tab = np.zeros((128, 5), np.int32)
        tab[:len(rec.draws)] = np.stack(rec.draws)
        chooses.append((d["choose1"], d["choose2"]))
        tables.append(tab)
    est = AdaPoseEstimator_v5(None, cfg, None, state_dict=sd, max_envs=2, precision=precision)
    choose = (np.stack([c[0] for c in chooses]).astype(np.int32), np.stack([c[1] for c in chooses]).astype(np.int32))
    boxes = est.estimate(*batch.args(), choose=choose, ransac_idx=np.stack(tables))
    eng = est.estimator
    for e in range(2):
        # (i) the fit itself is exact: the oracle's RANSAC + Umeyama on the DEVICE's NOCS / depth with the same sample table
        nocs_d, depth_d = eng.nocs[e].cpu().numpy(), eng.depth[e].cpu().numpy()
        K1 = eng.Kp[e].cpu().numpy().reshape(3, 3)
        cam = O.back_project(depth_d.flatten(), chooses[e][0], K1)
        s_o, R_o, t_o, _ = O.similarity_ransac(nocs_d, cam, rand_idx=tables[e])
        want = O.box_from_fit(nocs_d, s_o, R_o, t_o, batch.E1[e])
        np.testing.assert_allclose(boxes[e], want, rtol=0, atol=1e-6)
        # (ii) against the reference's own box.  RANSAC is discontinuous: a 1e-4 change of one residual can flip an
        # inlier or the early-exit iteration (align.py:78-87), so two numerically different but correct pipelines agree
        # to the inlier-set granularity, not to 1 mm; the bounds are that granularity on these ~1 m boxes (which inlier
        # flips depends on the rounding pattern, hence one bound per operand format).
        px, deg, mm, cmm = O.parity_errors(boxes[e], g["boxes"][e], batch.K[e], batch.E1[e], min_z=MIN_Z)
        assert deg < TOL_DEG and mm < mm_bound and cmm < cmm_bound, (precision, e, px, deg, mm, cmm)
    est.estimator.close()


def test_size_independent_properties_at_scale(golden):
    """num_envs = 256 (BASELINE configs[2] size): properties that do not need the oracle at that size --
    (i) tiling the 8 golden envs gives 32 identical copies of the 8 golden boxes (independence of environments, chunk
    boundaries and buffer reuse), (ii) a permutation of the environments permutes the boxes, (iii) the chunk size does not
    matter, (iv) sentinel envs stay sentinels.  Tolerance 2e-5 m: only the atomicAdd order of the per-env means differs."""
    g = golden
    base = synth.make_batch(8, seed=0)
    N = 256
    idx = np.arange(N) % 8
    args = [a[idx] for a in base.args()]
    c1, c2 = _golden_choose(g, 8)
    choose = (c1[idx], c2[idx])
    est = _make(max_envs=48)          # 256 = 5 * 48 + 16: exercises the partial last chunk
    boxes = est.estimate(*args, choose=choose)
    assert boxes.shape == (N, 8, 3)
    for e in range(8):
        same = boxes[idx == e]
        np.testing.assert_allclose(same, np.broadcast_to(same[0], same.shape), rtol=0, atol=2e-5)
        if not g["valid"][e]:
            np.testing.assert_array_equal(same[0], O.DEFAULT_BBOX)
        else:
            px, deg, mm, cmm = O.parity_errors(same[0], g["boxes"][e], base.K[e], base.E1[e], min_z=MIN_Z)
            assert px < TOL_PX and deg < TOL_DEG and mm < TOL_MM and cmm < TOL_MM
    perm = np.random.default_rng(3).permutation(N)
    boxes_p = est.estimate(*[a[perm] for a in args], choose=(choose[0][perm], choose[1][perm]))
    np.testing.assert_allclose(boxes_p, boxes[perm], rtol=0, atol=2e-5)
    est.estimator.close()
    est2 = _make(max_envs=7)
    boxes_c = est2.estimate(*[a[:40] for a in args], choose=(choose[0][:40], choose[1][:40]))
    np.testing.assert_allclose(boxes_c, boxes[:40], rtol=0, atol=2e-5)
    est2.estimator.close()


def test_single_view_nocs_matches_reference_golden(golden):
    """BASELINE configs[0..1] (one view per env: backbone + NOCS head) against the reference's own per-view NOCS maps, for the
    frames of both views of the golden environments."""
    g = golden
    batch = synth.make_batch(8, seed=0)
    est = _make(max_envs=3)                     # chunks of 3, 3, 2
    for view, (rgb, mask) in enumerate(((batch.rgb1, batch.mask1), (batch.rgb2, batch.mask2)), 1):
        ch = np.zeros((8, 1024), np.int32)
        for e in range(8):
            if g["valid"][e]:
                ch[e] = g[f"env{e}_choose{view}"]
        nocs, choose, valid = est.estimate_nocs_single_view(batch.K, rgb, mask, choose=ch)
        assert nocs.shape == (8, 1024, 3) and nocs.dtype == np.float32
        for e in range(8):
            if g["valid"][e]:
                assert valid[e]
                np.testing.assert_array_equal(choose[e], ch[e])
                assert float(np.abs(nocs[e] - g[f"env{e}_view{view}_nocs"]).max()) < 3e-3, (view, e)
    est.estimator.close()


def test_plain_constructor_init_stage_level():
    """weights.init_state_dict conditions the random init (NOCS x16, depth logits x4) so that box-level parity measures the
    kernels and not a division by ~0.  This test drops the conditioning (gains 1 = the distributions of the reference
    constructor) and checks what is well defined there: the network outputs (NOCS, soft-argmax depth, rotation) of one
    environment against the oracle."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from rgbmanip_b200.estimator import AdaPoseEstimator_v5
    sd = weights.init_state_dict(3, nocs_gain=1.0, prob_gain=1.0)
    cfg = {"name": "adapose_v5", "load": False, "img_size": 224, "use_depth": True, "n_pts": 1024, "direct_regression": True}
    batch = synth.make_batch(1, seed=21, special=False)
    np.random.seed(3)
    d = {}
    O.predict(sd, cfg, batch.K[0], batch.rgb1[0], batch.mask1[0], batch.E1[0], batch.rgb2[0], batch.mask2[0], batch.E2[0],
              both_views=False, details=d)
    est = AdaPoseEstimator_v5(None, cfg, None, state_dict=sd, max_envs=1, precision="fp16f8")
    est.estimate(*batch.args(), choose=(d["choose1"][None].astype(np.int32), d["choose2"][None].astype(np.int32)))
    eng = est.estimator
    nocs, depth, R = eng.nocs[0].cpu().numpy(), eng.depth[0].cpu().numpy(), eng.R[0].cpu().numpy().reshape(3, 3)
    assert np.abs(d["nocs"]).std() < 0.05                              # the unconditioned NOCS map really is nearly constant
    assert np.abs(nocs - d["nocs"]).max() < 5e-4, np.abs(nocs - d["nocs"]).max()
    assert np.abs(depth - d["depth"]).max() < 1e-3, np.abs(depth - d["depth"]).max()
    assert O.rotation_angle_deg(R, d["R"]) < 0.1
    est.estimator.close()
