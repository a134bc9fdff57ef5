"""afp_spectrum_db (csrc/afp_spectrum.cuh) against the float64 oracle (oracle/spectrum.py, pinned to the
reference's own waterfall rows) and against those rows directly.  The float32 FFT is compared on the linear
magnitude: |X_gpu - X_ref| <= 1e-5 * max|X_ref| per row (the parity contract's relative tolerance, taken
against the row's peak), plus the dB values of every bin within 60 dB of the peak to 0.01 dB."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from audio_filtering_pipeline_b200.engine import BatchFirEngine  # noqa: E402
from oracle import spectrum as ospec  # noqa: E402


@pytest.fixture(scope="module")
def eng():
    with BatchFirEngine([1.0], n_streams=1, max_frames=1) as e:
        yield e


def check(db_gpu, db_ref):
    lin_g, lin_r = 10.0 ** (db_gpu.astype(np.float64) / 20), 10.0 ** (db_ref / 20)
    peak = lin_r.max()
    assert np.abs(lin_g - lin_r).max() <= 1e-5 * max(peak, 1e-3), np.abs(lin_g - lin_r).max() / peak
    strong = db_ref >= db_ref.max() - 60
    assert np.abs(db_gpu[strong] - db_ref[strong]).max() <= 0.01


def test_golden_waterfall_rows(eng):
    g = load_golden("waterfall_spectrum")
    y = eng.spectrum_db(torch.from_numpy(g["outdata"]).cuda()).cpu().numpy()
    assert y.shape == g["spectrum_db"].shape
    for row_gpu, row_ref in zip(y, g["spectrum_db"]):
        check(row_gpu, row_ref)


@pytest.mark.parametrize("S,n", [(1, 64), (3, 128), (37, 256), (5, 512), (300, 1024), (9, 2048), (4, 4096)])
def test_spectrum_matches_oracle(eng, S, n):
    r = np.random.default_rng(S + n)
    t = np.arange(n)
    x = np.stack([0.5 * np.sin(2 * np.pi * r.uniform(0.001, 0.49) * t + r.uniform(0, 6)) +
                  10.0 ** r.uniform(-4, -1) * r.standard_normal(n) for _ in range(S)]).astype(np.float32)
    y = eng.spectrum_db(torch.from_numpy(x).cuda()).cpu().numpy()
    assert y.shape == (S, n // 2 + 1)
    for s in range(S):
        check(y[s], ospec.spectrum_db(x[s]))


def test_silence_and_pitched_rows(eng):
    n = 1024
    z = eng.spectrum_db(torch.zeros((2, n), device="cuda")).cpu().numpy()
    assert np.allclose(z, -240.0, atol=1e-3)                       # 20*log10(1e-12)
    x = (0.3 * np.random.default_rng(1).standard_normal((4, n + 8))).astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    out = torch.full((4, n // 2 + 1 + 6), 7.0, device="cuda")
    eng.spectrum_db(xd[:, 3:3 + n], out=out[:, 2:2 + n // 2 + 1])
    o = out.cpu().numpy()
    assert (o[:, :2] == 7.0).all() and (o[:, 2 + n // 2 + 1:] == 7.0).all()      # nothing written outside the rows
    for s in range(4):
        check(o[s, 2:2 + n // 2 + 1], ospec.spectrum_db(x[s, 3:3 + n]))


def test_rejects_other_lengths(eng):
    for n in (1000, 32, 8192):
        with pytest.raises(Exception):
            eng.spectrum_db(torch.zeros((1, n), device="cuda"))


def test_callback_collects_waterfall_rows():
    """make_audio_callback with show_waterfall on appends one row per block to chain.waterfall_data
    (stream_process_GUI_Waterfall.py:471-476): the row is the spectrum of what was written to outdata."""
    from audio_filtering_pipeline_b200 import compat
    B = 1024
    chain = compat.StreamChain(samplerate=44100, upsample_factor=1, blocksize=B, numtaps=129, cutoff=9000, dither=False,
                               show_waterfall=True)
    cb = compat.make_audio_callback(chain, verbose=False)
    x = (0.2 * np.random.default_rng(2).standard_normal((3 * B, 1))).astype(np.float32)
    for i in range(3):
        out = np.zeros((B, 1), dtype=np.float32)
        cb(x[i * B:(i + 1) * B], out, B, None, None)
        assert len(chain.waterfall_data) == i + 1
        check(chain.waterfall_data[-1], ospec.spectrum_db(out[:, 0]))
    chain.close()
