"""CPU restatement of the RL controller's observation and actor forward (test infrastructure only).

``get_observation`` follows models/controller/rl_pose.py:173-187; ``act_inference`` follows
algo/ppo/ppo/module.py:24-34,89-91 (Linear / ELU stack; cfg/controller/rl.yaml:28-32: 60 -> 96 -> 96 -> 32 -> 12)."""
from __future__ import annotations

import numpy as np


def get_observation(pose_queue, bbox_queue, accumulate_steps):
    """[T,N,7], [T,N,4] -> float32 [N, T*11 + T]: per-env history in ring order, then one_hot(accumulate_steps - 1, T)."""
    T, N = pose_queue.shape[:2]
    cur = np.concatenate([pose_queue.astype(np.float32), bbox_queue.astype(np.float32)], axis=-1)     # torch.tensor(..).float()
    ret = cur.transpose(1, 0, 2).reshape(N, -1)
    onehot = np.zeros((T,), np.float32)
    onehot[accumulate_steps - 1] = 1.0
    return np.concatenate([ret, np.broadcast_to(onehot[None], (N, T))], axis=-1)


def act_inference(actor_sd, obs):
    """actor_sd: {"actor.0.weight", "actor.0.bias", "actor.2.weight", ...} (nn.Sequential indices of the Linear layers)."""
    idx = sorted({int(k.split(".")[1]) for k in actor_sd if k.startswith("actor.")})
    x = obs.astype(np.float32)
    for n, i in enumerate(idx):
        x = x @ np.asarray(actor_sd[f"actor.{i}.weight"], np.float32).T + np.asarray(actor_sd[f"actor.{i}.bias"], np.float32)
        if n != len(idx) - 1:
            x = np.where(x > 0, x, np.expm1(np.minimum(x, 0))).astype(np.float32)       # ELU, alpha = 1
    return x
