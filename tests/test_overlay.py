"""The import overlay binds the reference's dotted name to the B200 estimator (no GPU needed: nothing is constructed)."""
import importlib
import sys


def test_overlay_rebinds_reference_symbol():
    from rgbmanip_b200 import overlay
    from rgbmanip_b200.estimator import AdaPoseEstimator_v5
    try:
        overlay.install()
        mod = importlib.import_module("models.pose_estimator.AdaPose.interface_v5")
        assert mod.AdaPoseEstimator_v5 is AdaPoseEstimator_v5
        # same public surface as the reference class (interface_v5.py:39,213,229)
        for name in ("estimate", "predict"):
            assert callable(getattr(AdaPoseEstimator_v5, name))
        import inspect
        params = list(inspect.signature(AdaPoseEstimator_v5.estimate).parameters)[1:8]
        assert params == ["camera_intrinsic_batch", "rgb1_batch", "view1_mask_batch", "view1_extrinsic_batch",
                          "rgb2_batch", "view2_mask_batch", "view2_extrinsic_batch"]
        assert list(inspect.signature(AdaPoseEstimator_v5.__init__).parameters)[1:4] == ["env", "cfg", "logger"]
    finally:
        overlay.uninstall()
        sys.modules.pop("models.pose_estimator.AdaPose.interface_v5", None)
