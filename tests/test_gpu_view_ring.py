"""Device view ring + feature cache (rgbmanip_b200/view_ring.py) on the B200 against
 (i) the reference controller's own queue state (tests/golden/view_ring.npz) and the queue oracle, and
 (ii) the estimator fed with the oracle's paired host frames: same boxes without re-running the backbone."""
import os

import numpy as np
import pytest
import torch

from oracle import adapose_oracle as O
from oracle import view_ring_oracle as V
from rgbmanip_b200 import synth, weights

pytestmark = [pytest.mark.gpu, pytest.mark.filterwarnings("ignore")]


def _estimator(max_envs, task="one_drawer_cabinet", direct=True):
    from rgbmanip_b200.estimator import AdaPoseEstimator_v5
    cfg = {"name": "adapose_v5", "task_name": task, "load": False, "img_size": 224, "use_depth": True, "n_pts": 1024,
           "direct_regression": direct, "real_world": False}
    return AdaPoseEstimator_v5(None, cfg, None, state_dict=weights.init_state_dict(0, regress_pose=direct), max_envs=max_envs)


def test_queue_state_matches_reference_controller(golden_dir):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from rgbmanip_b200.view_ring import ViewRing
    g = np.load(os.path.join(golden_dir, "view_ring.npz"))
    est = _estimator(2)                                   # chunk of 2 < 3 envs: the ring chunks like the estimator
    ring = ViewRing(est, 3, 5)
    q = V.ViewQueues(3, 5)
    for t, (color, mask, K, E, pose) in enumerate(V.view_ring_script()):
        # the script's frames are constant images whose value (10 .. 92) encodes (step, env); the queue state does not depend
        # on them, and as images they would leave the fp16 range of the backbone (the per-call range guard reports that)
        ring.add_view({"camera0": {"Color": color / 100.0, "Mask": mask, "Intrinsic": K, "Extrinsic": E}}, pose)
        ring.accumulate_steps += 1
        q.add_view(color, mask, K, E, pose)
        q.accumulate_steps += 1
        np.testing.assert_array_equal(ring.available, g[f"s{t}_available"])
        np.testing.assert_array_equal(ring.available_num, g[f"s{t}_available_num"])
        np.testing.assert_array_equal(ring.bbox_queue, g[f"s{t}_bbox_queue"])
        np.testing.assert_array_equal(ring.pose_queue, g[f"s{t}_pose_queue"])
        np.testing.assert_array_equal(ring.pair_slots().cpu().numpy(), q.pair_slots())
        # which frames the pairing selects: the extrinsics carry the (step, env) code of the script
        sl = ring.pair_slots().cpu().numpy()
        for v, key in ((0, "E1"), (1, "E2")):
            want = g[f"s{t}_{key}"]
            got = np.array([ring.extrinsic[sl[v, e], e, 0].item() if sl[v, e] >= 0 else 0.0 for e in range(3)])
            np.testing.assert_array_equal(got, want)
    est.estimator.close()


def test_ring_estimation_equals_estimator_on_paired_frames():
    """Controller-like sequence on synthetic scenes: after every step the ring's boxes (cached features, stereo head only)
    equal estimate() on the host frames the reference would have paired (backbone re-run for both views)."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from rgbmanip_b200.view_ring import ViewRing
    N, T = 4, 3
    est = _estimator(4)
    ring = ViewRing(est, N, T)
    q = V.ViewQueues(N, T)
    scenes = [synth.make_batch(N, seed=20 + s, special=False) for s in range(3)]
    frames = []
    for s, b in enumerate(scenes):                        # 6 camera steps: both views of three scenes
        frames.append((b.rgb1.astype(np.float64), b.mask1.astype(np.float64), b.K, b.E1))     # float64, as the controller's queues hold them
        frames.append((b.rgb2.astype(np.float64), b.mask2.astype(np.float64), b.K, b.E2))
    frames[2] = (frames[2][0], np.where(np.arange(N)[:, None, None] == 1, 0.0, frames[2][1]), frames[2][2], frames[2][3])   # env 1 blind once
    for t, (rgb, mask, K, E) in enumerate(frames):
        pose = np.full((N, 7), float(t))
        ring.add_view({"camera0": {"Color": rgb, "Mask": mask, "Intrinsic": K, "Extrinsic": E}}, pose)
        ring.accumulate_steps += 1
        q.add_view(rgb, mask, K, E, pose)
        q.accumulate_steps += 1
        got = ring.get_estimation()
        sl = q.pair_slots()
        Kb, rgb1, m1, E1, rgb2, m2, E2 = q.estimation_inputs()
        ch = ring.choose.cpu().numpy()
        c1 = np.stack([ch[max(sl[0, e], 0), e] for e in range(N)]).astype(np.int32)
        c2 = np.stack([ch[max(sl[1, e], 0), e] for e in range(N)]).astype(np.int32)
        want = est.estimate(Kb, rgb1, m1, E1, rgb2, m2, E2, choose=(c1, c2))
        for e in range(N):
            if sl[0, e] < 0 or sl[1, e] < 0 or m1[e].sum() == 0 or m2[e].sum() == 0:
                np.testing.assert_array_equal(got[e], O.DEFAULT_BBOX)
                np.testing.assert_array_equal(want[e], O.DEFAULT_BBOX)
            else:
                np.testing.assert_allclose(got[e], want[e], rtol=0, atol=2e-5)     # same kernels, same inputs (atomics order only)
    est.estimator.close()


def test_ring_branch_b_equals_estimator_on_paired_frames():
    """The ring with the RANSAC + Umeyama fit (direct_regression=False, use_depth=True; interface_v5.py:322-338): the same
    boxes as estimate() on the paired frames, given the same pixel subsets and the same RANSAC draws."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from rgbmanip_b200.view_ring import ViewRing
    N = 3
    est = _estimator(2, direct=False)                     # chunk of 2 < 3 envs
    ring = ViewRing(est, N, 4)
    b = synth.make_batch(N, seed=31, special=False)
    for rgb, m, E in ((b.rgb1, b.mask1, b.E1), (b.rgb2, b.mask2, b.E2)):
        ring.add_view({"camera0": {"Color": rgb, "Mask": m, "Intrinsic": b.K, "Extrinsic": E}}, np.zeros((N, 7)))
        ring.accumulate_steps += 1
    ridx = np.random.default_rng(3).integers(0, 1024, size=(N, 128, 5)).astype(np.int32)
    got = ring.get_estimation(ransac_idx=ridx)
    ch = ring.choose.cpu().numpy()
    want = est.estimate(b.K, b.rgb1, b.mask1, b.E1, b.rgb2, b.mask2, b.E2, choose=(ch[0], ch[1]), ransac_idx=ridx)
    assert not np.array_equal(want[0], O.DEFAULT_BBOX)
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-5)
    assert ring.get_estimation().shape == (N, 8, 3)       # device-drawn samples
    est.estimator.close()


def test_mug_corner_order():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from rgbmanip_b200.view_ring import MUG_CORNER_ORDER, ViewRing
    b = synth.make_batch(2, seed=4, special=False)
    outs = []
    for task in ("one_drawer_cabinet", "mugs"):
        est = _estimator(2, task)
        ring = ViewRing(est, 2, 5)
        for rgb, m, E in ((b.rgb1, b.mask1, b.E1), (b.rgb2, b.mask2, b.E2)):
            ring.add_view({"camera0": {"Color": rgb, "Mask": m, "Intrinsic": b.K, "Extrinsic": E}}, np.zeros((2, 7)))
            ring.accumulate_steps += 1
        outs.append(ring.get_estimation())
        est.estimator.close()
    np.testing.assert_allclose(outs[1], outs[0][:, MUG_CORNER_ORDER], rtol=0, atol=2e-5)


def test_device_actor_matches_reference(golden_dir):
    """adp_actor_forward on the ring's device queues against the reference's get_observation + act_inference (golden)."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from rgbmanip_b200.actor import DeviceActor
    from rgbmanip_b200.view_ring import ViewRing
    g = np.load(os.path.join(golden_dir, "actor.npz"))
    sd = {k: g[k] for k in g.files if k.startswith("actor.")}
    est = _estimator(2)
    ring = ViewRing(est, 3, 5)
    actor = DeviceActor(sd, ring)
    assert actor.dims == [60, 96, 96, 32, 12]
    for t, (color, mask, K, E, pose) in enumerate(V.view_ring_script()):
        if t >= 5:
            break
        ring.add_view({"camera0": {"Color": color, "Mask": mask, "Intrinsic": K, "Extrinsic": E}}, pose * 0.1)
        ring.accumulate_steps += 1
        act, obs = actor.act_inference()
        np.testing.assert_array_equal(obs.cpu().numpy(), g[f"s{t}_obs"])
        np.testing.assert_allclose(act.cpu().numpy(), g[f"s{t}_act"], rtol=0, atol=2e-5)
    with pytest.raises(IndexError):
        actor.act_inference(accumulate_steps=6)
    est.estimator.close()


def test_attach_rebinds_a_live_controller():
    """view_ring.attach swaps reset_queue / add_view / get_estimation of a ControlInterface-like object and keeps the numpy
    queue attributes that get_observation / get_reward read (rl_pose.py:156-187) in sync."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from rgbmanip_b200 import view_ring

    class FakeController:                     # the attributes ControlInterface.__init__ sets (rl_pose.py:22-44)
        pass

    ci = FakeController()
    ci.estimator = _estimator(2)
    ci.num_envs, ci.max_steps, ci.accumulate_steps = 2, 5, 3
    ring = view_ring.attach(ci)
    assert ci.accumulate_steps == 0 and ci.pose_queue.shape == (5, 2, 7) and ci.available.sum() == 0
    b = synth.make_batch(2, seed=4, special=False)
    for rgb, m, E in ((b.rgb1, b.mask1, b.E1), (b.rgb2, b.mask2, b.E2)):
        ci.add_view({"camera0": {"Color": rgb, "Mask": m, "Intrinsic": b.K, "Extrinsic": E}}, np.full((2, 7), 0.5))
        ci.accumulate_steps += 1
    assert ci.available[:2].sum() == 4 and list(ci.available_num) == [2, 2]
    assert np.all(ci.pose_queue[:2] == 0.5) and np.all(ci.bbox_queue[:2, :, 2] > ci.bbox_queue[:2, :, 0])
    box = ci.get_estimation()
    want = ci.estimator.estimate(b.K, b.rgb1, b.mask1, b.E1, b.rgb2, b.mask2, b.E2,
                                 choose=(ring.choose[0].cpu().numpy(), ring.choose[1].cpu().numpy()))
    np.testing.assert_allclose(box, want, rtol=0, atol=2e-5)
    ci.reset_queue()
    assert ci.available.sum() == 0 and ci.accumulate_steps == 0
    ci.estimator.estimator.close()
