"""The driver's smoke entry point (__graft_entry__.smoke) as a GPU test: the kernel routing it asserts (tensor kernel for C2
and the 4x interpolator, row-pair and single-row FFMA2 kernels, AGC) must follow every change of the per-call kernel choice."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
pytest.importorskip("torch")

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def test_smoke_entry_point(capsys):
    import __graft_entry__ as g
    g.smoke()
    assert "smoke ok" in capsys.readouterr().out
