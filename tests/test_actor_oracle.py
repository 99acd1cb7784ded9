"""oracle/actor_oracle.py against the reference's own ControlInterface.get_observation + ActorCritic.act_inference
(tests/golden/actor.npz from oracle/make_golden.py actor)."""
import os

import numpy as np

from oracle import actor_oracle as A
from oracle import view_ring_oracle as V


def test_observation_and_actor_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "actor.npz"))
    sd = {k: g[k] for k in g.files if k.startswith("actor.")}
    q = V.ViewQueues(3, 5)
    for t, (color, mask, K, E, pose) in enumerate(V.view_ring_script()):
        if t >= 5:
            break
        q.add_view(color, mask, K, E, pose * 0.1)
        q.accumulate_steps += 1
        obs = A.get_observation(q.pose_queue, q.bbox_queue, q.accumulate_steps)
        assert obs.shape == (3, 60) and obs.dtype == np.float32
        np.testing.assert_array_equal(obs, g[f"s{t}_obs"])
        np.testing.assert_allclose(A.act_inference(sd, obs), g[f"s{t}_act"], rtol=0, atol=2e-5)
    assert np.abs(g["s4_act"]).max() > 0.05          # the golden policy output is not degenerate
