"""Device-resident view ring + per-view feature cache.

Replaces the host queues of the reference's RL controller, ``models/controller/rl_pose.py``:
``reset_queue`` (:85-97), ``add_view`` (:118-150) and ``get_estimation`` (:189-223).  The reference keeps every frame of
every environment as float64 on the host (``[max_steps, N, 480, 640, 3]``: 151 GB at 4096 envs), re-walks the ring in
Python on every step and pushes BOTH views of every environment through the estimator again.  Here

* ``add_view`` uploads only the new frame, runs preprocessing + the backbone ONCE for it and keeps what the stereo head
  needs per (ring slot, env) on the device: the 32 x 224 x 224 feature map (fp32 + the fp16 twin the volume builder
  gathers), K', the 1024 sampled pixels, validity, extrinsic.  The mask box of :130-150 comes out of the preprocess
  kernel's bounding-box pass (including the quirk at :132: once ANY environment sees the object, every environment's
  slot counts as available);
* ``get_estimation`` pairs the views with the reference's rule (the k-th available view of an env goes to batch slot
  k % 2, so slot 0 = last even-indexed, slot 1 = last odd-indexed available view), gathers the two cached feature maps
  of every environment and runs only volume -> U-Net -> decode -> fit.  One backbone pass per controller step instead of
  two (81.8 vs 138.9 GFLOP per estimate).

Differences from the reference, by construction: the random 1024-pixel subset of a view is drawn once, when the view
is added (the reference redraws it at every estimate; same distribution), and frames themselves are not retained.
"""
from __future__ import annotations

import numpy as np
import torch

MUG_CORNER_ORDER = [0, 2, 4, 6, 1, 3, 5, 7]      # rl_pose.py:220-221


class ViewRing:
    def __init__(self, estimator, num_envs: int, max_steps: int, height: int = 480, width: int = 640):
        self.estimator = estimator
        if hasattr(estimator, "_ensure_capacity"):
            estimator._ensure_capacity(num_envs)         # an auto-sized estimator grows its chunk capacity to the ring's batch
        eng = estimator.estimator
        if getattr(eng, "branch_c", False) or getattr(eng, "arch", "v5") != "v5":
            raise NotImplementedError("the view ring drives the cost-volume head with the direct-regression fit (every shipped config) "
                                      "or the RANSAC + Umeyama fit (direct_regression=False, use_depth=True)")
        self.num_envs, self.max_steps, self.h, self.w = int(num_envs), int(max_steps), int(height), int(width)
        self.device = eng.device
        T, N, S, P = self.max_steps, self.num_envs, eng.S, eng.P
        dev = self.device
        self.feat = torch.zeros((T, N, S, S, 32), dtype=torch.float32, device=dev)
        self.feat16 = torch.zeros((T, N, S, S, 32), dtype=torch.float16, device=dev) if eng.feat16 is not None else None
        self.Kp = torch.zeros((T, N, 9), dtype=torch.float64, device=dev)
        self.choose = torch.zeros((T, N, P), dtype=torch.int32, device=dev)
        self.valid = torch.zeros((T, N), dtype=torch.uint8, device=dev)
        self.extrinsic = torch.zeros((T, N, 16), dtype=torch.float64, device=dev)
        self.intrinsic = torch.zeros((T, N, 9), dtype=torch.float64, device=dev)
        self.pose = torch.zeros((T, N, 7), dtype=torch.float64, device=dev)
        self.bbox = torch.zeros((T, N, 4), dtype=torch.float64, device=dev)
        self.avail = torch.zeros((T, N), dtype=torch.float64, device=dev)
        self.avail_num = torch.zeros((N,), dtype=torch.int32, device=dev)
        self._hw = torch.tensor([self.h, self.w, self.h, self.w], dtype=torch.float64, device=dev)
        self.accumulate_steps = 0          # advanced by the caller after add_view, as in the reference (rl_pose.py:116)
        self._adds = 0

    @property
    def eng(self):
        return self.estimator.estimator          # looked up per call: an auto-sized estimator may rebuild its engine

    # ---- queue state in the reference's host format (small arrays: observations / rewards read them) ----
    @property
    def available(self):
        return self.avail.cpu().numpy()

    @property
    def available_num(self):
        return self.avail_num.cpu().numpy()

    @property
    def bbox_queue(self):
        return self.bbox.cpu().numpy()

    @property
    def pose_queue(self):
        return self.pose.cpu().numpy()

    def reset_queue(self):
        """rl_pose.py:85-97."""
        for t in (self.feat, self.feat16, self.Kp, self.choose, self.valid, self.extrinsic, self.intrinsic, self.pose, self.bbox,
                  self.avail, self.avail_num):
            if t is not None:
                t.zero_()
        self.accumulate_steps = 0

    def _dev(self, a, dtype=None):
        t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
        return t.to(self.device, non_blocking=True).contiguous()

    def add_view(self, image, cam_pose):
        """rl_pose.py:118-150: ``image`` is the env's camera dict (``image["camera0"]["Color" | "Mask" | "Intrinsic" |
        "Extrinsic"]``, one entry per environment), ``cam_pose`` [N,7]."""
        cam = image["camera0"]
        color, mask, K, E = cam["Color"], cam["Mask"], cam["Intrinsic"], cam["Extrinsic"]
        eng, N, slot = self.eng, self.num_envs, self.accumulate_steps % self.max_steps
        self._adds += 1
        with torch.cuda.device(self.device):
            any_pixel = torch.zeros((), dtype=torch.bool, device=self.device)
            for lo in range(0, N, eng.F):                 # the frame buffers hold 2 x max_envs frames: full-size backbone launches
                hi = min(N, lo + eng.F)
                n = hi - lo
                c = color[lo:hi]
                c = c if isinstance(c, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(c))
                # same staging policy as estimate(): float64 host frames are demoted to float32, uint8 means value / 255
                rgb_d = self._dev(c, self.estimator._rgb_dtype(c))
                mask_d = self._dev(mask[lo:hi])
                if mask_d.dtype not in (torch.uint8, torch.bool, torch.float32, torch.float64):
                    mask_d = (mask_d != 0).to(torch.uint8)       # segmentation ids: a cast could wrap 256 to 0
                K_d = self._dev(K[lo:hi], torch.float64)
                eng.preprocess(0, rgb_d, mask_d, K_d, n, seed=(self.estimator._seed + 104729 * self._adds + lo))
                eng.run_backbone(n)                       # ONE view: frames [0, n)
                self.feat[slot, lo:hi].copy_(eng.feat[:n])
                if self.feat16 is not None:
                    self.feat16[slot, lo:hi].copy_(eng.feat16[:n])
                self.Kp[slot, lo:hi].copy_(eng.Kp[:n])
                self.choose[slot, lo:hi].copy_(eng.choose[:n])
                self.valid[slot, lo:hi].copy_(eng.valid[:n])
                bb = eng.bbox_ws[:n].to(torch.float64)    # {ymin, xmin, ymax, xmax}; ymax < 0 <=> empty mask
                empty = bb[:, 2] < 0
                any_pixel = any_pixel | (~empty).any()
                box = torch.div(bb, self._hw)             # tensor divisor: a true IEEE division, as numpy's (scalar division multiplies by 1/x)
                box[empty] = torch.tensor([2.0, 2.0, 0.0, 0.0], dtype=torch.float64, device=self.device)   # (2H/H, 2W/W, 0, 0): :141-144
                self.bbox[slot, lo:hi].copy_(box)
            self.extrinsic[slot].copy_(self._dev(E, torch.float64).reshape(N, 16))
            self.intrinsic[slot].copy_(self._dev(K, torch.float64).reshape(N, 9))
            self.pose[slot].copy_(self._dev(cam_pose, torch.float64).reshape(N, 7))
            # :132 tests the pixel count over ALL environments: one visible object marks every environment's slot available
            self.avail[slot] = any_pixel.to(torch.float64)
            self.avail_num += any_pixel.to(torch.int32)
        eng.check_error_flag()

    def pair_slots(self):
        """[2, N] ring indices feeding estimate()'s view 1 / view 2 (-1: that batch slot stays empty); rl_pose.py:199-208."""
        av = self.avail > 0
        k = torch.cumsum(av.to(torch.int64), 0) - 1
        ring = torch.arange(self.max_steps, device=self.device)[:, None]
        out = []
        for par in (0, 1):
            m = av & ((k % 2) == par)
            out.append(torch.where(m, ring, torch.full_like(ring, -1)).max(0).values)
        return torch.stack(out)

    def get_estimation(self, return_tensor: bool = False, ransac_idx=None):
        """rl_pose.py:189-223 -> [N,8,3] float64 world-frame boxes (sentinel where an env has no valid view pair).
        ``ransac_idx`` ([N,128,5] int32, optional) replays given RANSAC draws of the Umeyama fit (direct_regression=False);
        by default they come from the device hash, as in ``estimate()``."""
        eng, N = self.eng, self.num_envs
        E = eng.E
        out = torch.empty((N, 8, 3), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            sl = self.pair_slots()
            for lo in range(0, N, E):
                hi = min(N, lo + E)
                n = hi - lo
                idx = torch.arange(lo, hi, device=self.device)
                ext = []
                for v in (0, 1):
                    s = sl[v, lo:hi]
                    have = s >= 0
                    sc = s.clamp(min=0)
                    o = v * E
                    eng.feat[o:o + n].copy_(self.feat[sc, idx])
                    if self.feat16 is not None:
                        eng.feat16[o:o + n].copy_(self.feat16[sc, idx])
                    eng.Kp[o:o + n].copy_(self.Kp[sc, idx])
                    eng.choose[o:o + n].copy_(self.choose[sc, idx])
                    eng.valid[o:o + n].copy_(self.valid[sc, idx] * have.to(torch.uint8))
                    ext.append(self.extrinsic[sc, idx].contiguous())
                ridx = None
                if ransac_idx is not None and not eng.regress_pose:
                    ridx = torch.as_tensor(np.ascontiguousarray(ransac_idx[lo:hi])).to(torch.int32).to(self.device)
                eng.stereo(n, ext[0], ext[1], ransac_idx=ridx, seed=7919 * self._adds + lo)
                out[lo:hi].copy_(eng.bbox[:n])
            if self.estimator.cfg.get("task_name") == "mugs":
                out = out[:, MUG_CORNER_ORDER]
            if return_tensor:
                return out
            res = out.cpu().numpy()
        eng.check_error_flag()
        return res


def attach(control_interface, estimator=None):
    """Swap the queue methods of a live ``ControlInterface`` (rl_pose.py:14) for a device ring; observations keep reading
    ``pose_queue`` / ``bbox_queue`` / ``available`` as numpy arrays."""
    ci = control_interface
    ring = ViewRing(estimator or ci.estimator, ci.num_envs, ci.max_steps)

    def sync():
        ci.pose_queue, ci.bbox_queue = ring.pose_queue, ring.bbox_queue
        ci.available, ci.available_num = ring.available, ring.available_num

    def reset_queue():
        ring.reset_queue()
        ci.accumulate_steps = 0
        ci.pred_bbox = np.zeros((ci.max_steps, ci.num_envs, 8, 3))
        ci.gt_bbox = np.zeros((ci.max_steps, ci.num_envs, 8, 3))
        sync()

    def add_view(image, cam_pose):
        ring.accumulate_steps = ci.accumulate_steps
        ring.add_view(image, cam_pose)
        sync()

    ci.reset_queue, ci.add_view, ci.get_estimation = reset_queue, add_view, ring.get_estimation
    ci.view_ring = ring
    reset_queue()
    return ring
