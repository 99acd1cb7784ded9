"""Checkpoint ingestion: the reference stores ``torch.save(DataParallel(model).state_dict())`` and loads it
with ``strict=True`` (interface_v5.py:48,55-56) -- 150 tensors incl. BatchNorm ``num_batches_tracked``."""
import numpy as np
import pytest
import torch

from rgbmanip_b200 import weights as W


def test_param_table_is_the_reference_state_dict_layout():
    table = W.param_table(True)
    assert len(table) == 150
    assert sum(1 for n, _, _ in table if n.endswith("num_batches_tracked")) == 10      # the ten BatchNorm3d of CostRegNet
    assert len(W.param_table(False)) < 150                                             # direct_regression=False drops the pose heads


def test_checkpoint_roundtrip_with_dataparallel_prefix(tmp_path):
    sd = W.init_state_dict(3)
    ckpt = {"module." + k: torch.from_numpy(np.asarray(v).copy()) for k, v in sd.items()}
    path = str(tmp_path / "adapose.pth")
    torch.save(ckpt, path)
    got = W.load_checkpoint(path)
    assert list(got.keys()) == list(sd.keys())
    for k in sd:
        np.testing.assert_array_equal(got[k], sd[k])


def test_strict_loading_rejects_missing_unexpected_and_misshaped(tmp_path):
    sd = W.init_state_dict(0)
    missing = {k: v for k, v in sd.items() if k != "nocs_head.4.bias"}
    with pytest.raises(KeyError):
        W.check_state_dict(missing)
    extra = dict(sd, bogus=np.zeros(1, np.float32))
    with pytest.raises(KeyError):
        W.check_state_dict(extra)
    bad = dict(sd)
    bad["img_extractor.final.weight"] = np.zeros((32, 64, 3, 3), np.float32)
    with pytest.raises(ValueError):
        W.check_state_dict(bad)
    with pytest.raises(FileNotFoundError):        # a missing file fails like the reference's torch.load
        W.load_checkpoint(str(tmp_path / "nope.pth"))


def test_fold_bn_equals_eval_mode_batchnorm3d():
    sd = W.init_state_dict(1, randomize_bn=True)
    name = "cost_regularization.conv3.bn"
    bn = torch.nn.BatchNorm3d(sd[f"{name}.weight"].shape[0]).eval()
    bn.load_state_dict({k: torch.from_numpy(np.asarray(sd[f"{name}.{k}"])) for k in
                        ("weight", "bias", "running_mean", "running_var", "num_batches_tracked")})
    x = torch.randn(2, bn.num_features, 3, 4, 5)
    scale, shift = W.fold_bn(sd, name)
    with torch.no_grad():
        want = bn(x)
    got = x * torch.from_numpy(scale).view(1, -1, 1, 1, 1) + torch.from_numpy(shift).view(1, -1, 1, 1, 1)
    assert float((got - want).abs().max()) < 1e-5
