"""The view-queue oracle (oracle/view_ring_oracle.py) against the reference's own ControlInterface run unmodified
(tests/golden/view_ring.npz, produced by oracle/make_golden.py view_ring): availability flags incl. the global-nonzero
quirk, mask boxes, and which ring slot feeds which argument of estimate()."""
import os

import numpy as np

from oracle import view_ring_oracle as V


def test_queues_and_pairing_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "view_ring.npz"))
    q = V.ViewQueues(3, 5)
    for t, (color, mask, K, E, pose) in enumerate(V.view_ring_script()):
        q.add_view(color, mask, K, E, pose)
        q.accumulate_steps += 1
        np.testing.assert_array_equal(q.available, g[f"s{t}_available"])
        np.testing.assert_array_equal(q.available_num, g[f"s{t}_available_num"])
        np.testing.assert_array_equal(q.bbox_queue, g[f"s{t}_bbox_queue"])
        np.testing.assert_array_equal(q.pose_queue, g[f"s{t}_pose_queue"])
        Kb, rgb1, m1, E1, rgb2, m2, E2 = q.estimation_inputs()
        np.testing.assert_array_equal(Kb[:, 0, 0], g[f"s{t}_K"])
        np.testing.assert_array_equal(rgb1[:, 0, 0, 0], g[f"s{t}_rgb1"])
        np.testing.assert_array_equal(rgb2[:, 0, 0, 0], g[f"s{t}_rgb2"])
        np.testing.assert_array_equal(m1.sum((1, 2)), g[f"s{t}_m1"])
        np.testing.assert_array_equal(m2.sum((1, 2)), g[f"s{t}_m2"])
        np.testing.assert_array_equal(E1[:, 0, 0], g[f"s{t}_E1"])
        np.testing.assert_array_equal(E2[:, 0, 0], g[f"s{t}_E2"])
    # the script exercises the quirk (an env with an empty mask marked available) and the ring wrap-around
    assert g["s1_available"][1].sum() == 3 and (g["s1_m1"] == 0).any() or (g["s1_m2"] == 0).any()
    box = np.arange(3 * 24, dtype=np.float64).reshape(-1, 8, 3)
    np.testing.assert_array_equal(box[:, V.MUG_CORNER_ORDER], g["mug_box"])


def test_pair_slots_semantics():
    q = V.ViewQueues(2, 5, h=8, w=8)
    q.available[:, 0] = [1, 0, 1, 1, 0]          # available views of env 0 in slot order: 0, 2, 3 -> k = 0, 1, 2
    q.available[:, 1] = [0, 0, 0, 1, 0]
    sl = q.pair_slots()
    assert list(sl[:, 0]) == [3, 2]               # slot 0: last even-indexed (k = 2 -> ring 3); slot 1: last odd-indexed (k = 1 -> ring 2)
    assert list(sl[:, 1]) == [3, -1]
