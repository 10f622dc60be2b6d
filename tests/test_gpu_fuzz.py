"""
GPU parity soak (scripts/gpu_fuzz.py): random frame shapes, region soups, reference counts, all four flag
combinations and zero-heavy payloads through both host entry points, byte for byte against the oracle.
(A 6000-case run of the same generator - 15 087 frames - is recorded in DESIGN.md.)
"""
import importlib.util
import os
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_random_soak(gpu, po, oracle):
    spec = importlib.util.spec_from_file_location("gpu_fuzz", os.path.join(ROOT, "scripts", "gpu_fuzz.py"))
    fz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fz)
    bad, frames = fz.run(300, 424242, gc=gpu, po=po, orc=oracle, verbose=True)
    assert frames > 500 and bad == 0
