"""oracle/_ref (the run-time copy of the reference that bench.py's CPU arm times) reproduces the committed golden boxes."""
import os

import numpy as np
import pytest

from oracle import ref_stage
from rgbmanip_b200 import synth


def test_staged_reference_reproduces_golden_boxes(golden_dir):
    if ref_stage.root() is None:
        pytest.skip("no reference tree here (oracle/_ref is staged by __graft_entry__.build() in the build container)")
    est, cfg = ref_stage.load(seed=0)
    assert cfg["name"] == "adapose_v5"
    g = np.load(os.path.join(golden_dir, "e2e.npz"))
    b = synth.make_batch(8, seed=0).slice(0, 2)
    np.random.seed(0)
    out = est.estimate(*b.args())
    np.testing.assert_array_equal(out, g["boxes"][:2])           # the unmodified reference, bit for bit


def test_staged_files_are_byte_copies():
    if not (ref_stage.available() and os.path.isdir(ref_stage.ORIGINAL)):
        pytest.skip("needs both the staged copy and /root/reference")
    import filecmp
    for rel in ref_stage.FILES:
        src = os.path.join(ref_stage.ORIGINAL, rel)
        if os.path.exists(src):
            assert filecmp.cmp(src, os.path.join(ref_stage.STAGED, rel), shallow=False), rel
