"""bench.py's roofline accounting as pure functions: the MMA work the split-FP16 tensor kernel issues per output sample
(three products over ceil-padded 64-sample Toeplitz chunks, U = 1 and the polyphase forms) and the choice of the bound."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("afp_bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("n,U,chunks", [(1001, 1, 17), (513, 1, 9), (51, 1, 2), (251, 2, 3), (1001, 2, 9), (1001, 3, 7), (1001, 4, 5),
                                        (251, 4, 2)])
def test_issued_work_follows_the_kernel_chunking(bench, n, U, chunks):
    # tc2_build_cores: D = (n - 1 + 63 U) / (64 U) chunks beyond the first
    assert (n - 1 + 63 * U) // (64 * U) + 1 == chunks
    peaks = {"n64": (1366.0, 1700.0, 48.0), "n128": (2005.0, 1700.0, 64.0), "n256": (2137.0, 1700.0, 128.0)}
    t = bench.tensor_block(n, 1e11, peaks, {"bf16_tflops": 1660.7}, U)
    taps_per_phase = -(-n // U)
    assert t["issued_over_algorithmic"] == pytest.approx(3.0 * chunks * 64 / taps_per_phase, abs=2e-3)
    assert t["issued_tflops"] == pytest.approx(3 * 2 * chunks * 64 * 1e11 / 1e12, rel=1e-3)
    mix = 3.0 / (2.0 / 2005.0 + 1.0 / 1366.0) if U in (1, 3) else 2005.0
    assert t["instruction_mix_roof_tflops"] == pytest.approx(mix, rel=1e-3)
    assert t["frac_of_instruction_mix_roof"] == pytest.approx(t["issued_tflops"] / mix, abs=2e-3)
