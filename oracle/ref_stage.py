"""Run-time copy of the UNMODIFIED reference for the CPU baseline -- TEST / MEASUREMENT INFRASTRUCTURE, NOT THE PRODUCT.

``/root/reference`` exists only in the build container.  ``stage()`` (called by ``__graft_entry__.build()``) copies the
handful of pure-Python reference modules that ``models.pose_estimator.AdaPose.interface_v5`` imports, byte for byte, into
``oracle/_ref/`` -- a git-ignored directory (nothing of the reference enters the history) that still travels to the GPU box
with the snapshot, like the built ``.so``.  ``bench.py --impl reference`` and the ``cpu_baseline`` leg then time the
reference's own ``AdaPoseEstimator_v5.estimate`` (``kind: "reference"``); without the copy they fall back to the oracle port
(``kind: "port"``).  Only ``bench.py``'s CPU legs and ``tests/`` import this module.

``load()`` imports the staged (or the original) reference exactly like ``oracle/make_golden.py`` does: MagicMock stand-ins
for the simulator packages it imports transitively, no-op ``.cuda()`` shims so that it stays on the host cores (that is the
arm being timed), ``eval()`` mode (SURVEY.md finding 0.3-1), weights of ``rgbmanip_b200.weights.init_state_dict``.
"""
from __future__ import annotations

import filecmp
import logging
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref")
ORIGINAL = os.environ.get("RGBMANIP_REFERENCE", "/root/reference")

# what `import models.pose_estimator.AdaPose.interface_v5` loads from the reference tree (+ the four estimator yamls)
FILES = [
    "LICENSE",
    "env/base_sapien_env.py", "env/base_viewer.py", "env/sapien_envs/base_manipulation.py", "env/sapien_envs/impedance_control.py",
    "env/sapien_envs/open_cabinet.py", "env/sapien_envs/osc_planner.py",
    "models/pose_estimator/AdaPose/interface_v5.py", "models/pose_estimator/AdaPose/lib/align.py",
    "models/pose_estimator/AdaPose/lib/network_v5.py", "models/pose_estimator/AdaPose/lib/pspnet.py",
    "models/pose_estimator/AdaPose/lib/rotation_utils.py", "models/pose_estimator/AdaPose/lib/utils.py",
    "models/pose_estimator/base_estimator.py",
    "utils/logger.py", "utils/sapien_utils.py", "utils/tools.py", "utils/transform.py",
    "cfg/pose_estimator/adapose_cabinet.yaml", "cfg/pose_estimator/adapose_drawer.yaml", "cfg/pose_estimator/adapose_mug.yaml",
    "cfg/pose_estimator/adapose_pot.yaml",
]
MOCKED = ["sapien", "sapien.core", "sapien.core.renderer", "sapien.utils", "mplib", "gym", "gym.spaces", "gym.vector",
          "gym.vector.utils", "gym.vector.utils.shared_memory", "ujson", "matplotlib", "matplotlib.pyplot", "trimesh", "transforms3d"]


def stage(verbose=True):
    """Copy the reference modules into oracle/_ref/ (no-op without /root/reference).  Returns the staged root or None."""
    if not os.path.isdir(ORIGINAL):
        return STAGED if available() else None
    for rel in FILES:
        src, dst = os.path.join(ORIGINAL, rel), os.path.join(STAGED, rel)
        if not os.path.exists(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
    if verbose:
        print("reference staged under", STAGED)
    return STAGED


def available():
    return os.path.exists(os.path.join(STAGED, "models/pose_estimator/AdaPose/interface_v5.py"))


def root():
    """The reference tree to import from: the staged copy if present, else the original (build container)."""
    if available():
        return STAGED
    if os.path.isdir(ORIGINAL):
        return ORIGINAL
    return None


def load(seed=0, task="drawer", direct_regression=True):
    """-> (reference AdaPoseEstimator_v5 on the CPU in eval mode with the seeded weights, its cfg)."""
    from unittest.mock import MagicMock

    import numpy as np
    import torch
    import torch.nn as nn
    import yaml
    r = root()
    if r is None:
        raise RuntimeError("no reference tree: run __graft_entry__.build() in the build container first")
    repo = os.path.dirname(HERE)
    for p in (repo, r):
        if p not in sys.path:
            sys.path.insert(0, p)
    for m in MOCKED:
        sys.modules.setdefault(m, MagicMock())
    torch.Tensor.cuda = lambda s, *a, **k: s          # keep the reference arm on the host cores
    nn.Module.cuda = lambda s, *a, **k: s
    from models.pose_estimator.AdaPose import interface_v5
    from rgbmanip_b200 import weights
    cfg = yaml.safe_load(open(os.path.join(r, f"cfg/pose_estimator/adapose_{task}.yaml")))
    cfg["load"] = False
    cfg["direct_regression"] = direct_regression
    est = interface_v5.AdaPoseEstimator_v5(None, cfg, logging.getLogger("reference"))
    sd = weights.init_state_dict(seed, regress_pose=direct_regression)
    est.estimator.load_state_dict({"module." + k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    est.estimator.eval()
    return est, cfg


if __name__ == "__main__":
    stage()
