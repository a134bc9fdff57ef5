"""AGC with target_level / max_gain < 1e-4: the host then selects the kernel variant that keeps the reference's
`target / (rms + 1e-10)` form instead of the rsqrt form (csrc/afp_abi.cu::launch_agc).  Quiet inputs, so that
the gain is neither clamped to max_gain nor to the 0.1 floor over most of the block."""
import numpy as np
import pytest

from conftest import assert_parity

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from audio_filtering_pipeline_b200.engine import BatchFirEngine  # noqa: E402
from oracle import agc as oagc  # noqa: E402


@pytest.mark.parametrize("n,W", [(1024, 512), (512, 256), (2048, 300)])
def test_exact_gain_variant_matches_oracle(n, W):
    S = 9
    r = np.random.default_rng(n + W)
    env = np.repeat(10.0 ** r.uniform(-5.5, -2.0, (S, 8)), n // 8, axis=1)
    x = (env * r.standard_normal((S, n))).astype(np.float32)
    kw = dict(target_level=1e-3, window_size=W, max_gain=100.0, attack=0.01, release=0.1)
    assert kw["target_level"] / kw["max_gain"] < 1e-4
    with BatchFirEngine([1.0], n_streams=1, max_frames=1) as eng:
        y = eng.agc_block(torch.from_numpy(x).cuda(), clip=0.99, **kw).cpu().numpy()
    unclamped = 0
    for s in range(S):
        ref = np.clip(oagc.apply_agc(x[s], **kw), -0.99, 0.99)
        assert_parity(y[s], ref, f"stream {s}")
        g = np.abs(ref[np.abs(x[s]) > 0] / x[s][np.abs(x[s]) > 0])
        unclamped += int(((g > 0.1001) & (g < 99.9)).sum())
    assert unclamped > S * n // 4          # the comparison is not just the limiter
