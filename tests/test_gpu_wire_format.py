"""afp_process_host_interleaved: the [frames, channels] wire format of a sounddevice stream
(stream_process.py:128-136) against the planar entry point - bit-identical, every output format."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from audio_filtering_pipeline_b200 import compat, design  # noqa: E402
from audio_filtering_pipeline_b200.engine import BatchFirEngine  # noqa: E402


def taps(n=101):
    return design.create_fir_filter(cutoff=7000, numtaps=n, samplerate=48000)


@pytest.mark.parametrize("sessions,channels,fmt,dither,U", [
    (1, 1, "f32", "none", 1), (1, 2, "f32", "none", 1), (3, 2, "i16", "tpdf", 1), (2, 3, "i24_packed", "tpdf", 2),
    (5, 2, "i24", "tpdf", 4), (1, 8, "f32", "rpdf", 2), (4, 1, "i16", "none", 1)])
def test_interleaved_equals_planar(sessions, channels, fmt, dither, U):
    S, B, nblk = sessions * channels, 480, 3
    rng = np.random.default_rng(sessions * 10 + channels)
    x = (0.2 * rng.standard_normal((nblk, sessions, B, channels))).astype(np.float32)
    kw = dict(n_streams=S, max_frames=B, upsample=U, chain="dac", out_format=fmt, dither=dither, seed=7)
    with BatchFirEngine(U * taps(), **kw) as a, BatchFirEngine(U * taps(), **kw) as b:
        for i in range(nblk):
            yi = a.process_host_interleaved(x[i])
            planar = np.ascontiguousarray(x[i].transpose(0, 2, 1)).reshape(S, B)       # row = session*channels + channel
            yp = b.process_host(planar)
            n_out = b.out_frames(B)
            if fmt == "i24_packed":
                want = yp.reshape(sessions, channels, n_out, 3).transpose(0, 2, 1, 3)
            else:
                want = yp.reshape(sessions, channels, n_out).transpose(0, 2, 1)
            assert yi.shape == want.shape and yi.dtype == want.dtype
            assert np.array_equal(yi, want), f"block {i}"


def test_single_session_2d_shape_and_errors():
    with BatchFirEngine(taps(), n_streams=2, max_frames=256) as e:
        y = e.process_host_interleaved(np.zeros((256, 2), dtype=np.float32))
        assert y.shape == (256, 2) and y.dtype == np.float32 and not y.any()
        with pytest.raises(ValueError):
            e.process_host_interleaved(np.zeros((256, 3), dtype=np.float32))
        with pytest.raises(Exception):
            e.process_host_interleaved(np.zeros((512, 2), dtype=np.float32))        # frames > max_frames


def test_stereo_callback_writes_both_columns():
    """make_audio_callback on a two-channel stream: each column is the mono chain of that channel."""
    B, nblk = 1024, 3
    rng = np.random.default_rng(3)
    x = (0.1 * rng.standard_normal((nblk * B, 2))).astype(np.float32)
    stereo = compat.StreamChain(samplerate=44100, upsample_factor=2, blocksize=B, numtaps=129, cutoff=9000, channels=2,
                                dither=False)
    monos = [compat.StreamChain(samplerate=44100, upsample_factor=2, blocksize=B, numtaps=129, cutoff=9000, channels=1,
                                dither=False) for _ in range(2)]
    cb = compat.make_audio_callback(stereo, verbose=False)
    cbs = [compat.make_audio_callback(m, verbose=False) for m in monos]
    for i in range(nblk):
        blk = x[i * B:(i + 1) * B]
        out = np.full((B, 2), np.nan, dtype=np.float32)
        cb(blk, out, B, None, None)
        for ch in range(2):
            o1 = np.full((B, 1), np.nan, dtype=np.float32)
            cbs[ch](blk[:, ch:ch + 1], o1, B, None, None)
            assert np.array_equal(out[:, ch], o1[:, 0])
    for c in [stereo] + monos:
        c.close()
