"""Input handling and launch paths of AdaPoseEstimator_v5.estimate on the B200: uint8 / float64 / pageable host frames,
integer segmentation masks, argument checks, CUDA-graph replay per chunk size, the per-call fp16 range guard, one engine
per device."""
import numpy as np
import pytest
import torch

from oracle import adapose_oracle as O
from rgbmanip_b200 import synth, weights

pytestmark = [pytest.mark.gpu, pytest.mark.filterwarnings("ignore")]

ATOL = 2e-5      # metres: identical inputs still differ by the atomicAdd order of the per-env means
CFG = {"name": "adapose_v5", "task_name": "one_drawer_cabinet", "load": False, "img_size": 224, "use_depth": True,
       "n_pts": 1024, "direct_regression": True, "real_world": False}


def _make(cfg_extra=None, sd=None, **kw):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from rgbmanip_b200.estimator import AdaPoseEstimator_v5
    return AdaPoseEstimator_v5(None, dict(CFG, **(cfg_extra or {})), None, state_dict=sd or weights.init_state_dict(0), **kw)


def _choose_for(batch, n, seed=0):
    """Pixel subsets drawn by the oracle's preprocessing (so that every variant decodes the same pixels)."""
    np.random.seed(seed)
    c1, c2 = [], []
    for e in range(n):
        c1.append(O.prepare_model_input(batch.rgb1[e], batch.mask1[e], batch.K[e])[1])
        c2.append(O.prepare_model_input(batch.rgb2[e], batch.mask2[e], batch.K[e])[1])
    return np.stack(c1).astype(np.int32), np.stack(c2).astype(np.int32)


def test_uint8_frames_mean_value_over_255():
    """uint8 RGB (a 4x smaller upload) is defined as the float32 path on rgb.float() / 255 -- torchvision ToTensor semantics
    (interface_v5.py:52,149): bit-identical boxes."""
    b = synth.make_batch(3, seed=12, special=False)
    ch = _choose_for(b, 3)
    u1, u2 = (np.clip(b.rgb1, 0, 1) * 255).round().astype(np.uint8), (np.clip(b.rgb2, 0, 1) * 255).round().astype(np.uint8)
    f1 = (torch.from_numpy(u1).float() / 255).numpy()
    f2 = (torch.from_numpy(u2).float() / 255).numpy()
    est = _make(max_envs=4)
    c = est.estimate(b.K, f1, b.mask1, b.E1, f2, b.mask2, b.E2, choose=ch)
    crops_f32 = est.estimator.crops[:3].cpu().numpy()
    a = est.estimate(b.K, u1, b.mask1, b.E1, u2, b.mask2, b.E2, choose=ch)
    np.testing.assert_array_equal(est.estimator.crops[:3].cpu().numpy(), crops_f32)      # the crops themselves: bit for bit
    np.testing.assert_allclose(a, c, rtol=0, atol=ATOL)
    d = est.estimate(b.K, torch.from_numpy(u1).cuda(), b.mask1, b.E1, torch.from_numpy(u2).cuda(), b.mask2, b.E2, choose=ch)   # zero-copy
    np.testing.assert_allclose(a, d, rtol=0, atol=ATOL)
    assert np.isfinite(a).all()
    with pytest.raises(TypeError):
        est.estimate(b.K, u1.astype(np.int32), b.mask1, b.E1, u2.astype(np.int32), b.mask2, b.E2, choose=ch)
    est.estimator.close()


def test_integer_segmentation_masks_use_nonzero_not_a_wrapping_cast():
    """A segmentation id of 256 is foreground (a uint8 cast would wrap it to 0 and blind the env)."""
    b = synth.make_batch(2, seed=13, special=False)
    ch = _choose_for(b, 2)
    est = _make(max_envs=2)
    ref = est.estimate(*b.args(), choose=ch)
    ids1, ids2 = b.mask1.astype(np.int32) * 256, b.mask2.astype(np.int64) * 512
    got = est.estimate(b.K, b.rgb1, ids1, b.E1, b.rgb2, ids2, b.E2, choose=ch)
    np.testing.assert_allclose(got, ref, rtol=0, atol=ATOL)
    assert not np.array_equal(got[0], O.DEFAULT_BBOX)
    est.estimator.close()


def test_pageable_float64_frames_equal_pinned_float32():
    """What rl_pose.py:194-218 passes: pageable float64 numpy holding float32 renders.  Staged through pinned buffers and
    demoted to float32 on the way (lossless here): bit-identical to pinned float32 tensors, over several chunks."""
    b = synth.make_batch(5, seed=14, special=False)
    ch = _choose_for(b, 5)
    est = _make(max_envs=2, cfg_extra={"first_chunk_envs": 1})
    pinned = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in b.args()]
    a = est.estimate(*pinned, choose=ch)
    c = est.estimate(b.K, b.rgb1.astype(np.float64), b.mask1.astype(np.float64), b.E1, b.rgb2.astype(np.float64),
                     b.mask2.astype(np.float64), b.E2, choose=ch)
    np.testing.assert_allclose(a, c, rtol=0, atol=ATOL)
    # keep_float64: the crop is interpolated in double like cv2 does (interface_v5.py:148).  The crops differ at the ulp level;
    # through the fp16 pipeline that moves a box by up to the pipeline's own rounding noise (two runs that are each within
    # 0.5 px / 1 mm of the reference)
    est64 = _make(max_envs=2, cfg_extra={"keep_float64": True})
    d = est64.estimate(b.K, b.rgb1.astype(np.float64), b.mask1, b.E1, b.rgb2.astype(np.float64), b.mask2, b.E2, choose=ch)
    for e in range(5):
        px, deg, mm, cmm = O.parity_errors(d[e], a[e], b.K[e], b.E1[e], min_z=0.5)
        assert px < 1.0 and mm < 1.0, (e, px, mm)
    est.estimator.close(); est64.estimator.close()


def test_argument_checks():
    b = synth.make_batch(2, seed=15, special=False)
    est = _make(max_envs=2)
    from rgbmanip_b200._lib import AdpError
    with pytest.raises(ValueError):                      # mask / rgb shape mismatch
        est.estimate(b.K, b.rgb1, b.mask1[:, :400], b.E1, b.rgb2, b.mask2, b.E2)
    with pytest.raises(ValueError):                      # pixel indices outside the 224 x 224 crop
        bad = np.full((2, 1024), 224 * 224, np.int32)
        est.estimate(*b.args(), choose=(bad, bad))
    with pytest.raises(AdpError, match="440"):           # a frame that cannot hold the 440-pixel crop window
        est.estimate(b.K, b.rgb1[:, :400, :400], b.mask1[:, :400, :400], b.E1, b.rgb2[:, :400, :400], b.mask2[:, :400, :400], b.E2)
    assert est.estimate(*b.args()).shape == (2, 8, 3)    # and the estimator is still usable
    est.estimator.close()


def test_cuda_graph_replay_per_chunk_size_equals_eager_launches():
    """Every chunk size is captured on its second appearance and replayed afterwards (engine.run_chunk): the boxes of
    replayed chunks equal the eager ones bit for bit up to the atomicAdd order of the per-env means, for full, first and
    tail chunk sizes alike."""
    b = synth.make_batch(11, seed=16, special=True)
    est = _make(max_envs=4, cfg_extra={"first_chunk_envs": 2})          # host inputs: chunks of 2, 3, 3, 3
    eng = est.estimator
    eager = est.estimate(*b.args())
    assert eng._graphs == {} or set(eng._graphs) <= {3}
    second = est.estimate(*b.args())
    third = est.estimate(*b.args())
    assert set(eng._graphs) == {2, 3}, eng._graphs.keys()
    l0 = eng.lib.adp_launch_count()
    fourth = est.estimate(*b.args())
    assert eng.lib.adp_launch_count() - l0 > 4 * 80        # replays are counted as the launches they contain
    # est._calls changes the sampling seed from call to call, so compare with a fresh estimator replaying the same call numbers
    est2 = _make(max_envs=4, cfg_extra={"first_chunk_envs": 2}, use_graph=False)
    for want in (eager, second, third, fourth):
        got = est2.estimate(*b.args())
        np.testing.assert_allclose(got, want, rtol=0, atol=2e-5)
    assert est2.estimator._graphs == {}
    est.estimator.close(); est2.estimator.close()


def test_fp16_range_guard_fires_on_every_call_not_only_the_first():
    """a later batch that leaves the fp16 range must not turn into silent sentinel boxes."""
    from rgbmanip_b200._lib import AdpError
    b = synth.make_batch(2, seed=4, special=False)
    est = _make(max_envs=2)
    ok = est.estimate(*b.args())
    assert np.isfinite(ok).all()
    hot1, hot2 = b.rgb1 * 3e5, b.rgb2 * 3e5                   # activations far beyond 65504
    with pytest.raises(AdpError, match="fp16 range"):
        est.estimate(b.K, hot1, b.mask1, b.E1, hot2, b.mask2, b.E2)
    again = est.estimate(*b.args())                           # the flag is cleared once reported
    assert np.isfinite(again).all()
    # tensor-returning path: reported by the next call / an explicit check instead of a synchronisation
    est.estimate(b.K, hot1, b.mask1, b.E1, hot2, b.mask2, b.E2, return_tensor=True)
    with pytest.raises(AdpError, match="fp16 range"):
        est.check_error_flag()
    est.estimator.close()


def test_one_engine_per_device_in_one_process():
    """Kernel attributes (dynamic shared memory) are per device: a second engine on another GPU of the same process works."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    b = synth.make_batch(2, seed=17, special=False)
    ch = _choose_for(b, 2)
    a = _make(max_envs=2, device="cuda:0").estimate(*b.args(), choose=ch)
    c = _make(max_envs=2, device="cuda:1").estimate(*b.args(), choose=ch)
    np.testing.assert_allclose(a, c, rtol=0, atol=2e-5)


def test_single_process_multi_gpu_equals_one_gpu():
    """cfg["devices"] = [0, 1]: one estimator object in one process (what train.py:238-240 builds) drives one engine per GPU from
    its own host thread; the boxes equal the one-GPU result of the same call bit for bit."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    b = synth.make_batch(9, seed=18, special=True)
    one = _make(max_envs=2)
    multi = _make(max_envs=2, cfg_extra={"devices": list(range(min(torch.cuda.device_count(), 4)))})
    for call in range(3):                     # eager, graph capture, graph replay
        a = one.estimate(*b.args(), sample_seed=77 + call)
        c = multi.estimate(*b.args(), sample_seed=77 + call)
        np.testing.assert_array_equal(a, c)
    t = multi.estimate(*b.args(), sample_seed=79, return_tensor=True)
    assert t.device == multi.device and tuple(t.shape) == (9, 8, 3)
    np.testing.assert_array_equal(t.cpu().numpy(), a)
    multi.check_error_flag()


def test_window_row_upload_equals_whole_frame_upload():
    """Host frames: only the rows under each frame's crop window are uploaded (cfg['window_upload'], default on).  Same boxes as
    uploading the frames whole, bit for bit, over several chunks and with blind / border / small-mask environments in the batch;
    the bytes copied drop by the ratio of window rows to frame rows."""
    b = synth.make_batch(12, seed=19, special=True)
    m1 = b.mask1.copy(); m1[5] = 0
    a_est = _make(max_envs=4, cfg_extra={"first_chunk_envs": 2})
    w_est = _make(max_envs=4, cfg_extra={"first_chunk_envs": 2, "window_upload": False})
    args = (b.K, b.rgb1, m1, b.E1, b.rgb2, b.mask2, b.E2)
    pinned = [torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in args]
    for call in range(3):
        a = a_est.estimate(*args, sample_seed=5 + call)                 # pageable numpy -> staged rows
        p = a_est.estimate(*pinned, sample_seed=5 + call)               # pinned -> direct row copies
        w = w_est.estimate(*args, sample_seed=5 + call)
        np.testing.assert_array_equal(a, w)
        np.testing.assert_array_equal(p, w)
    np.testing.assert_array_equal(a[5], O.DEFAULT_BBOX)
    assert a_est.h2d_bytes < 0.75 * w_est.h2d_bytes * 2, (a_est.h2d_bytes, w_est.h2d_bytes)     # a_est ran twice as many calls
    a_est.estimator.close(); w_est.estimator.close()


def test_auto_sized_workspace_grows_with_the_batch():
    """Without max_envs / max_envs_per_chunk the estimator sizes itself: 16 environments to start with, grown on demand (up to
    AUTO_CHUNK_MAX per chunk) when a larger batch arrives; results equal those of an explicitly sized estimator."""
    from rgbmanip_b200.estimator import AUTO_CHUNK_MAX
    b = synth.make_batch(20, seed=4, n_unique=5)
    auto = _make()
    assert auto._auto_chunk and auto.estimator.E == 16
    small = auto.estimate(*b.slice(0, 4).args(), sample_seed=3)
    assert auto.estimator.E == 16
    boxes = auto.estimate(*b.args(), sample_seed=3)
    assert auto.estimator.E == 20 <= AUTO_CHUNK_MAX
    fixed = _make(max_envs=20)
    assert not fixed._auto_chunk
    np.testing.assert_array_equal(boxes, fixed.estimate(*b.args(), sample_seed=3))
    np.testing.assert_array_equal(small, boxes[:4])
    assert not _make({"max_envs_per_chunk": 8})._auto_chunk
