"""Make an unmodified RGBManip checkout use the B200 estimator.

``train.py:37`` and ``models/controller/heuristic_pose.py:11`` import
``models.pose_estimator.AdaPose.interface_v5.AdaPoseEstimator_v5`` by name and ``heuristic_pose.py:52-56`` gates on
``isinstance``.  ``install()`` registers a module object under that dotted name *before* they are imported, whose
``AdaPoseEstimator_v5`` is :class:`rgbmanip_b200.estimator.AdaPoseEstimator_v5`; the Hydra yamls
(``pose_estimator=adapose_{cabinet,drawer,mug,pot}``, ``name: adapose_v5``) are read unchanged.

    import rgbmanip_b200.overlay as o; o.install()      # first lines of a launcher script, then run train.py's my_app()
"""
from __future__ import annotations

import importlib
import sys
import types

TARGET = "models.pose_estimator.AdaPose.interface_v5"
TARGET_BASELINE = "models.pose_estimator.AdaPose.interface_baseline"      # train.py:38,242-244 (name: adapose_baseline)


def install():
    from . import estimator
    mod = types.ModuleType(TARGET)
    mod.__doc__ = "B200-native replacement installed by rgbmanip_b200.overlay"
    mod.AdaPoseEstimator_v5 = estimator.AdaPoseEstimator_v5
    mod.StereoPoseNet_with_depth = None
    sys.modules[TARGET] = mod
    modb = types.ModuleType(TARGET_BASELINE)
    modb.__doc__ = mod.__doc__
    modb.AdaPoseEstimator_baseline = estimator.AdaPoseEstimator_baseline
    sys.modules[TARGET_BASELINE] = modb
    try:   # make `from models.pose_estimator.AdaPose import interface_v5` resolve to the same object
        pkg = importlib.import_module("models.pose_estimator.AdaPose")
        setattr(pkg, "interface_v5", mod)
        setattr(pkg, "interface_baseline", modb)
    except Exception:
        pass
    return mod


def uninstall():
    mod = sys.modules.pop(TARGET, None)
    modb = sys.modules.pop(TARGET_BASELINE, None)
    pkg = sys.modules.get("models.pose_estimator.AdaPose")
    if pkg is not None and modb is not None and getattr(pkg, "interface_baseline", None) is modb:
        delattr(pkg, "interface_baseline")
    if pkg is not None and mod is not None and getattr(pkg, "interface_v5", None) is mod:
        delattr(pkg, "interface_v5")      # `from models.pose_estimator.AdaPose import interface_v5` resolves to the real module again
