"""Seeded synthetic inputs for the AdaPose hot path.

Shapes and value ranges follow what the reference's callers hand to ``estimate``
(models/controller/rl_pose.py:189-218, models/controller/heuristic_pose.py:57-65):
  K    [N,3,3]        pin-hole intrinsics of the 640x480 on-hand camera (base_manipulation.py:20)
  rgb  [N,480,640,3]  float in [0,1]
  mask [N,480,640]    0/1 foreground of the target part
  E    [N,4,4]        world->camera extrinsic, OpenCV axes (base_sapien_env.py:157-158)
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

IMG_H, IMG_W = 480, 640
FX = FY = 240.0 / math.tan(0.5)
CX, CY = 320.0, 240.0


@dataclass
class StereoBatch:
    K: np.ndarray
    rgb1: np.ndarray
    mask1: np.ndarray
    E1: np.ndarray
    rgb2: np.ndarray
    mask2: np.ndarray
    E2: np.ndarray

    def args(self):
        """Positional arguments of ``estimate`` (interface_v5.py:213-214)."""
        return (self.K, self.rgb1, self.mask1, self.E1, self.rgb2, self.mask2, self.E2)

    def __len__(self):
        return self.K.shape[0]

    def slice(self, lo, hi):
        return StereoBatch(*(a[lo:hi] for a in self.args()))


def intrinsics():
    return np.array([[FX, 0.0, CX], [0.0, FY, CY], [0.0, 0.0, 1.0]], dtype=np.float64)


def _texture(rng, dtype):
    """Deterministic low-pass texture: coarse noise upsampled bilinearly + a little fine noise."""
    coarse = rng.random((16, 21, 3), dtype=np.float32)
    ys = np.linspace(0, 14.999, IMG_H, dtype=np.float32)
    xs = np.linspace(0, 19.999, IMG_W, dtype=np.float32)
    y0 = ys.astype(np.int32)
    x0 = xs.astype(np.int32)
    fy = (ys - y0)[:, None, None]
    fx = (xs - x0)[None, :, None]
    a = coarse[y0][:, x0]
    b = coarse[y0][:, x0 + 1]
    c = coarse[y0 + 1][:, x0]
    d = coarse[y0 + 1][:, x0 + 1]
    img = (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy
    img = 0.85 * img + 0.15 * rng.random((IMG_H, IMG_W, 3), dtype=np.float32)
    return np.clip(img, 0.0, 1.0).astype(dtype)


def _ellipse_mask(cx, cy, ax, ay):
    yy, xx = np.mgrid[0:IMG_H, 0:IMG_W]
    return (((xx - cx) / ax) ** 2 + ((yy - cy) / ay) ** 2) <= 1.0


def _mask(rng, kind):
    if kind == "empty":
        return np.zeros((IMG_H, IMG_W), bool)
    if kind == "small":  # < 1024 foreground pixels after the 224x224 resize -> np.pad(...,'wrap') path
        cx, cy = rng.uniform(200, 440), rng.uniform(150, 330)
        return _ellipse_mask(cx, cy, rng.uniform(3, 6), rng.uniform(3, 6))
    if kind == "border":  # crop window has to be shifted back inside the frame (utils.py:22-37)
        side = int(rng.integers(0, 4))
        cx = (8.0, IMG_W - 9.0, rng.uniform(100, 540), rng.uniform(100, 540))[side]
        cy = (rng.uniform(100, 380), rng.uniform(100, 380), 6.0, IMG_H - 7.0)[side]
        return _ellipse_mask(cx, cy, rng.uniform(30, 110), rng.uniform(20, 80))
    cx, cy = rng.uniform(200, 440), rng.uniform(150, 330)
    return _ellipse_mask(cx, cy, rng.uniform(30, 110), rng.uniform(20, 80))


def _look_at(center, target, roll):
    z = target - center
    z /= np.linalg.norm(z)
    up = np.array([0.0, 0.0, 1.0])
    x = np.cross(z, up)
    if np.linalg.norm(x) < 1e-6:
        x = np.array([1.0, 0.0, 0.0])
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    c, s = math.cos(roll), math.sin(roll)
    x, y = c * x + s * y, -s * x + c * y
    R = np.stack([x, y, z])  # rows: camera axes in world
    E = np.eye(4)
    E[:3, :3] = R
    E[:3, 3] = -R @ center
    return E


def _camera_pair(rng):
    lo, hi = np.array([-0.3, -0.3, 0.4]), np.array([0.3, 0.3, 1.0])  # cfg/controller/rl.yaml:6-7
    target = np.array([0.65, 0.0, 0.3]) + rng.uniform(-1, 1, 3) * np.array([0.15, 0.2, 0.2])
    c1 = rng.uniform(lo, hi)
    d = rng.standard_normal(3)
    d /= np.linalg.norm(d)
    c2 = np.clip(c1 + d * rng.uniform(0.1, 0.6), lo - 0.1, hi + 0.1)
    E1 = _look_at(c1, target, rng.uniform(-math.pi / 8, math.pi / 8))
    E2 = _look_at(c2, target, rng.uniform(-math.pi / 8, math.pi / 8))
    return E1, E2


def make_batch(num_envs: int, seed: int = 0, dtype=np.float32, special: bool = True,
               n_unique: int | None = None) -> StereoBatch:
    """Seeded stereo batch. With ``special`` roughly 2% small, 1% empty and 2% border masks are injected
    (always at least one of each when num_envs >= 8) so the sentinel / wrap / window-shift paths run.
    ``n_unique`` < num_envs generates that many distinct envs and tiles them (bench-size batches)."""
    nu = num_envs if n_unique is None else min(n_unique, num_envs)
    rng = np.random.default_rng(seed)
    kinds = ["normal"] * nu
    if special and nu >= 8:
        n_small = max(1, round(0.02 * nu))
        n_empty = max(1, round(0.01 * nu))
        n_border = max(1, round(0.02 * nu))
        slots = rng.permutation(nu)[: n_small + n_empty + n_border]
        for i, s in enumerate(slots):
            kinds[s] = "small" if i < n_small else ("empty" if i < n_small + n_empty else "border")
    K = np.repeat(intrinsics()[None], nu, 0)
    rgb1 = np.empty((nu, IMG_H, IMG_W, 3), dtype)
    rgb2 = np.empty((nu, IMG_H, IMG_W, 3), dtype)
    mask1 = np.empty((nu, IMG_H, IMG_W), bool)
    mask2 = np.empty((nu, IMG_H, IMG_W), bool)
    E1 = np.empty((nu, 4, 4))
    E2 = np.empty((nu, 4, 4))
    for e in range(nu):
        rgb1[e] = _texture(rng, dtype)
        rgb2[e] = _texture(rng, dtype)
        mask1[e] = _mask(rng, kinds[e])
        mask2[e] = _mask(rng, "normal" if kinds[e] == "empty" and rng.random() < 0.5 else kinds[e])
        E1[e], E2[e] = _camera_pair(rng)
    b = StereoBatch(K, rgb1, mask1, E1, rgb2, mask2, E2)
    if nu < num_envs:
        idx = np.arange(num_envs) % nu
        b = StereoBatch(*(a[idx] for a in b.args()))
    return b
