"""Host-side logic of the estimator that needs no GPU: the chunk schedule of estimate()."""
import pytest

from rgbmanip_b200.estimator import chunk_bounds


@pytest.mark.parametrize("N,E,first", [(0, 74, 0), (1, 74, 0), (74, 74, 16), (75, 74, 0), (128, 74, 0), (128, 74, 16), (1024, 74, 0),
                                       (1024, 74, 16), (4096, 74, 16), (40, 7, 0), (256, 48, 16), (5, 4, 16)])
def test_chunk_bounds_cover_the_batch_in_equal_chunks(N, E, first):
    b = chunk_bounds(N, E, first)
    assert [lo for lo, _ in b] == [0] * (N > 0) + [hi for _, hi in b[:-1]]          # contiguous, in order
    assert (b[-1][1] if b else 0) == N
    sizes = [hi - lo for lo, hi in b]
    assert all(0 < s <= E for s in sizes)
    ramp = 0
    if first and N > E:      # the ramp: first, 2 first, 4 first ... while as much again remains behind each ramp chunk
        while ramp < len(sizes) and sizes[ramp] == min(first << ramp, E) and sizes[ramp] < E and sum(sizes[ramp + 1:]) >= sizes[ramp]:
            ramp += 1
        assert min(first, E) == E or (ramp >= 1 and sizes[0] == first)
    body = sizes[ramp:]
    assert max(body, default=0) - min(body, default=0) <= 1      # equal chunks: no short tail off the graph / whole-wave path
    assert len(set(sizes)) <= 8                           # CUDA graphs per call (the engine keeps at most 8)
    assert len(body) == -(-sum(body) // E) if body else True     # and no more chunks than necessary


def test_chunk_bounds_examples():
    assert chunk_bounds(128, 74, 0) == [(0, 64), (64, 128)]                        # an 8-GPU shard of 1024 envs, device-resident
    assert chunk_bounds(128, 74, 16) == [(0, 16), (16, 48), (48, 88), (88, 128)]   # the same shard uploaded from the host
    assert [h - l for l, h in chunk_bounds(1024, 74, 16)][:4] == [16, 32, 64, 71]
    assert chunk_bounds(8, 8, 16) == [(0, 8)]


def test_branch_c_host_tail_equals_reference(golden_dir):
    """The host tail of branch C (cv2 PnP + box, estimator.pnp_box_tail) on the reference's own NOCS / scale -> its boxes."""
    import os
    import numpy as np
    from rgbmanip_b200 import synth
    from rgbmanip_b200.estimator import DEFAULT_BBOX, pnp_box_tail
    g = np.load(os.path.join(golden_dir, "branch_c.npz"))
    b = synth.make_batch(4, seed=3, special=False)
    scale = g["left_scale"].copy()
    valid = np.array([1, 1, 1, 0], bool)
    box = pnp_box_tail(g["nocs1"], g["pts2d1"].astype(np.float32), scale, valid, b.K, b.E1)
    np.testing.assert_allclose(box[:3], g["boxes"][:3], rtol=1e-9, atol=1e-9)
    np.testing.assert_array_equal(box[3], DEFAULT_BBOX)
    scale[0] = np.nan
    np.testing.assert_array_equal(pnp_box_tail(g["nocs1"], g["pts2d1"].astype(np.float32), scale, valid, b.K, b.E1)[0], DEFAULT_BBOX)
