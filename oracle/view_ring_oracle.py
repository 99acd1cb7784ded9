"""CPU restatement of the caller-side view queues of RGBManip's RL controller (test infrastructure only).

Follows ``models/controller/rl_pose.py`` of the reference: ``reset_queue`` (:85-97), ``add_view`` (:118-150) and the
view pairing of ``get_estimation`` (:189-223).  SURVEY.md 8(f)-1.  Nothing under ``rgbmanip_b200/`` imports this file.
"""
from __future__ import annotations

import numpy as np

IMG_H, IMG_W = 480, 640          # CAMERA_INTRINSIC[-1], CAMERA_INTRINSIC[-2] (env/sapien_envs/open_cabinet.py:20)


class ViewQueues:
    def __init__(self, num_envs: int, max_steps: int, h: int = IMG_H, w: int = IMG_W):
        self.num_envs, self.max_steps, self.h, self.w = num_envs, max_steps, h, w
        self.reset_queue()

    def reset_queue(self):
        """rl_pose.py:85-97."""
        n, t = self.num_envs, self.max_steps
        self.image_queue = np.zeros((t, n, self.h, self.w, 3))
        self.mask_queue = np.zeros((t, n, self.h, self.w))
        self.bbox_queue = np.zeros((t, n, 4))
        self.pose_queue = np.zeros((t, n, 7))
        self.intrinsic_queue = np.zeros((t, n, 3, 3))
        self.extrinsic_queue = np.zeros((t, n, 4, 4))
        self.available = np.zeros((t, n))
        self.available_num = np.zeros((n,), dtype=np.int32)
        self.accumulate_steps = 0

    def add_view(self, color, mask, intrinsic, extrinsic, cam_pose):
        """rl_pose.py:118-150.  Note the quirk at :132: ``p_env.shape[0]`` is the number of mask pixels over ALL
        environments, so as soon as any environment sees the object every environment's slot is marked available
        (an environment with an empty mask gets the degenerate box (2, 2, 0, 0) and the estimator's sentinel later)."""
        i_id = self.accumulate_steps % self.max_steps
        self.image_queue[i_id] = color
        self.mask_queue[i_id] = mask
        self.pose_queue[i_id] = cam_pose
        self.intrinsic_queue[i_id] = intrinsic
        self.extrinsic_queue[i_id] = extrinsic
        any_pixel = bool(np.any(mask))
        for e in range(self.num_envs):
            rows, cols = np.nonzero(mask[e])
            if any_pixel:
                x_min = rows.min() if rows.size else self.h * 2
                x_max = rows.max() if rows.size else 0
                y_min = cols.min() if cols.size else self.w * 2
                y_max = cols.max() if cols.size else 0
                self.available[i_id, e] = 1
                self.available_num[e] += 1
            else:
                x_min, x_max, y_min, y_max = self.h * 2, 0, self.w * 2, 0
                self.available[i_id, e] = 0
            self.bbox_queue[i_id, e] = (x_min / self.h, y_min / self.w, x_max / self.h, y_max / self.w)

    def pair_slots(self):
        """rl_pose.py:199-208: walking the ring in slot order, the k-th available view of an environment goes to batch slot
        k % 2, so slot 0 ends up holding its last even-indexed available view and slot 1 the last odd-indexed one.
        Returns int [2, num_envs] ring indices, -1 where the batch slot stays zero-filled."""
        out = -np.ones((2, self.num_envs), dtype=np.int64)
        used = np.zeros(self.num_envs, dtype=np.int64)
        for i in range(self.max_steps):
            for j in range(self.num_envs):
                if self.available[i, j]:
                    out[used[j] % 2, j] = i
                    used[j] += 1
        return out

    def estimation_inputs(self):
        """The seven arguments of ``estimator.estimate`` as built at rl_pose.py:194-218."""
        sl = self.pair_slots()
        n = self.num_envs
        K = np.zeros((2, n, 3, 3)); E = np.zeros((2, n, 4, 4))
        rgb = np.zeros((2, n, self.h, self.w, 3)); m = np.zeros((2, n, self.h, self.w))
        for s in range(2):
            for j in range(n):
                i = sl[s, j]
                if i >= 0:
                    K[s, j], E[s, j], rgb[s, j], m[s, j] = (self.intrinsic_queue[i, j], self.extrinsic_queue[i, j],
                                                             self.image_queue[i, j], self.mask_queue[i, j])
        return K[0], rgb[0], m[0], E[0], rgb[1], m[1], E[1]


MUG_CORNER_ORDER = [0, 2, 4, 6, 1, 3, 5, 7]      # rl_pose.py:220-221


def view_ring_script(num_envs=3, steps=9, seed=11):
    """Deterministic camera-step script shared by the golden generator and the tests: per step (color, mask, K, E, pose).
    Frames are constant images whose value encodes (step, env), so that the pairing can be read off the estimator's inputs."""
    rng = np.random.default_rng(seed)
    H, W = 480, 640
    out = []
    for t in range(steps):
        color = np.zeros((num_envs, H, W, 3))
        mask = np.zeros((num_envs, H, W))
        K = np.zeros((num_envs, 3, 3)); E = np.zeros((num_envs, 4, 4)); pose = np.zeros((num_envs, 7))
        for e in range(num_envs):
            color[e] = (t + 1) * 10 + e
            K[e] = np.eye(3) * ((t + 1) * 100 + e)
            E[e] = np.eye(4) * ((t + 1) * 1000 + e)
            pose[e] = (t + 1) + 0.1 * e
            # step 0: nobody sees the object; step 2: only env 1; otherwise random, with empty masks now and then
            see = (t != 0) and ((t != 2) or e == 1) and (rng.random() > 0.25)
            if see:
                r0, c0 = int(rng.integers(0, H - 40)), int(rng.integers(0, W - 40))
                r1, c1 = r0 + int(rng.integers(1, 40)), c0 + int(rng.integers(1, 40))
                mask[e, r0:r1, c0:c1] = 1
        out.append((color, mask, K, E, pose))
    return out

