"""AGC kernel (csrc/afp_agc.cuh) against the float64 oracle (oracle/agc.py, pinned to the reference's own
apply_agc by tests/test_oracle_golden.py) and against the golden blocks directly.  Tolerance: 1e-5
relative / -100 dBFS (conftest); the reference's error path (block returned unchanged) is bit-exact."""
import json

import numpy as np
import pytest

from conftest import assert_parity, load_golden

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from audio_filtering_pipeline_b200 import compat, design  # noqa: E402
from audio_filtering_pipeline_b200.engine import BatchFirEngine  # noqa: E402
from oracle import agc as oagc  # noqa: E402
from oracle import reference_chain as rc  # noqa: E402


@pytest.fixture(scope="module")
def eng():
    with BatchFirEngine([1.0], n_streams=1, max_frames=1) as e:
        yield e


def blocks(S, n, seed=3):
    """Streams with level steps between 1e-4 and 1.0 so that every gain regime occurs: max-gain limited,
    tracking, attack and release transients."""
    r = np.random.default_rng(seed)
    out = np.empty((S, n), dtype=np.float32)
    for s in range(S):
        k = int(r.integers(1, 7))
        edges = np.sort(r.integers(0, n, size=k - 1)) if k > 1 else np.array([], dtype=int)
        env = np.empty(n)
        for lo, hi in zip(np.concatenate([[0], edges]), np.concatenate([edges, [n]])):
            env[lo:hi] = 10.0 ** r.uniform(-4, 0)
        out[s] = np.clip(env * r.standard_normal(n), -1, 1)
    return out


def run(eng, x, **kw):
    return eng.agc_block(torch.from_numpy(x).cuda(), **kw).cpu().numpy()


def test_golden_blocks_of_the_reference(eng):
    g = load_golden("agc_blocks")
    meta = json.loads(str(g["meta"]))
    for m in meta:
        x, y_ref = g["x_" + m["block"]], g["y_" + m["key"]]
        y = run(eng, x[None, :], target_level=m["target_level"], window_size=m["window_size"],
                max_gain=m["max_gain"], attack=m["attack"], release=m["release"], clip=0.0)[0]
        if m["errored"]:
            assert np.array_equal(y, x), m["key"]
        else:
            assert_parity(y, y_ref, m["key"])
        yc = run(eng, x[None, :], target_level=m["target_level"], window_size=m["window_size"],
                 max_gain=m["max_gain"], attack=m["attack"], release=m["release"], clip=0.99)[0]
        assert_parity(yc, np.clip(y_ref, -0.99, 0.99), m["key"] + " clipped")
        assert np.abs(yc).max() <= np.float32(0.99)


@pytest.mark.parametrize("S,n,W", [(1, 1024, 512), (37, 1024, 512), (300, 1024, 512), (19, 4096, 512), (5, 512, 256),
                                   (7, 1000, 500), (3, 777, 101), (9, 130, 100), (2, 8192, 512), (33, 256, 128),
                                   (4, 2048, 2048), (3, 100, 100), (130, 64, 32)])
def test_agc_matches_oracle(eng, S, n, W):
    x = blocks(S, n, seed=S + n)
    kw = dict(target_level=0.1, window_size=W, max_gain=10.0, attack=0.01, release=0.1)
    y = run(eng, x, clip=0.99, **kw)
    for s in range(S):
        ref = np.clip(oagc.apply_agc(x[s], **kw), -0.99, 0.99)
        assert_parity(y[s], ref, f"stream {s}")


def test_attack_slower_than_release_takes_the_min_form(eng):
    x = blocks(6, 1024, seed=41)
    kw = dict(target_level=0.2, window_size=200, max_gain=20.0, attack=0.2, release=0.02)
    y = run(eng, x, clip=0.0, **kw)
    for s in range(6):
        assert_parity(y[s], oagc.apply_agc(x[s], **kw), f"stream {s}")


def test_zero_sample_constants_follow_the_reference_error_path(eng):
    """window 64: int(0.01*64) == 0 -> the attack branch raises in the reference.  A stream whose desired
    gain never rises above the smoothed one never takes it and is processed normally; the others come
    back unchanged (bit for bit)."""
    n = 512
    r = np.random.default_rng(8)
    rising = (np.linspace(0.02, 0.5, n) * r.standard_normal(n)).astype(np.float32)       # level up -> gain only falls?
    falling = (np.linspace(0.5, 0.02, n) * r.standard_normal(n)).astype(np.float32)
    x = np.stack([rising, falling, blocks(1, n, seed=5)[0]])
    kw = dict(target_level=0.1, window_size=64, max_gain=10.0, attack=0.01, release=0.1)
    y = run(eng, x, clip=0.0, **kw)
    n_err = 0
    for s in range(3):
        ref = oagc.apply_agc(x[s], **kw)
        if ref.dtype == np.float32:                      # error path: the input itself
            assert np.array_equal(y[s], x[s])
            n_err += 1
        else:
            assert_parity(y[s], ref, f"stream {s}")
    assert n_err >= 1
    # both constants zero (window 8): whatever branch comes first raises
    kw["window_size"] = 8
    y = run(eng, x, clip=0.99, **kw)
    assert np.array_equal(y, np.clip(x, -0.99, 0.99))


def test_window_longer_than_block_and_empty_window_pass_through(eng):
    x = blocks(4, 256, seed=2)
    for W in (257, 512, 0, -3):
        y = run(eng, x, window_size=W, clip=0.99)
        assert np.array_equal(y, np.clip(x, -0.99, 0.99))


def test_unaligned_rows_and_pitches(eng):
    """Rows that do not start on 16 bytes (odd pitch, offset view) take the scalar loads."""
    S, n = 6, 1001
    x = blocks(S, n + 3, seed=12)
    xd = torch.from_numpy(x).cuda()
    view = xd[:, 1:1 + n]                                 # row stride n+3, 4-byte offset
    out = torch.zeros((S, n + 5), dtype=torch.float32, device="cuda")
    eng.agc_block(view, out=out[:, 2:2 + n], window_size=400, clip=0.99)
    y = out.cpu().numpy()
    assert not y[:, :2].any() and not y[:, 2 + n:].any()  # nothing written outside the rows
    for s in range(S):
        ref = np.clip(oagc.apply_agc(x[s, 1:1 + n], window_size=400), -0.99, 0.99)
        assert_parity(y[s, 2:2 + n], ref, f"stream {s}")


def test_agc_ahead_of_the_fir_chain():
    """The stage order of stream_process_AGC.py:131-160: AGC + clip on the base-rate block, then the
    chain.  Oracle: AGC per block, then the streaming FIR over the concatenated gain-controlled blocks."""
    S, B, nblk, N = 5, 1024, 4, 129
    h = design.create_fir_filter(cutoff=6000, numtaps=N, samplerate=44100)
    x = blocks(S, B * nblk, seed=77)
    with BatchFirEngine(h, n_streams=S, max_frames=B, agc=dict(window_size=min(512, B // 2))) as e:
        xd = torch.from_numpy(x).cuda()
        y = torch.cat([e.process(xd[:, i * B:(i + 1) * B]) for i in range(nblk)], dim=1).cpu().numpy()
        yh = np.concatenate([e.process_host(x[:, i * B:(i + 1) * B]) for i in range(1)], axis=1)
        e.set_agc(False)
        e.reset()
        y_off = e.process(xd[:, :B]).cpu().numpy()
    for s in range(S):
        xa = np.concatenate([oagc.agc_stage(x[s, i * B:(i + 1) * B]) for i in range(nblk)]).astype(np.float32)
        assert_parity(y[s], rc.streaming_fir(h, xa), f"stream {s}")
        assert_parity(y_off[s], rc.streaming_fir(h, x[s, :B]), f"stream {s} (AGC off)")
    assert yh.shape == (S, B)


def test_compat_apply_agc_shape_and_values():
    x = blocks(1, 2048, seed=9)[0]
    y = compat.apply_agc(x.reshape(-1, 1), target_level=0.1, window_size=512)
    assert y.shape == (2048, 1) and y.dtype == np.float32
    assert_parity(y[:, 0], oagc.apply_agc(x, window_size=512), "compat.apply_agc")
    y1 = compat.apply_agc(x)                                   # 1-D in -> [n, 1] out (:52-53)
    assert y1.shape == (2048, 1)
    y2 = compat.apply_agc(x[:128], window_size=64)             # error path -> unchanged
    assert np.array_equal(y2[:, 0], x[:128])


def test_full_batch_properties():
    """16384 streams x 1024 frames (the bench batch): a sample of streams against the oracle, and for all
    of them the bounds that hold by construction: |y| <= clip and 0.1 <= y/x <= max_gain."""
    S, n = 16384, 1024
    g = torch.Generator(device="cuda").manual_seed(5)
    env = 10.0 ** (-3.0 * torch.rand((S, 8), device="cuda", generator=g)).repeat_interleave(n // 8, dim=1)
    x = (env * torch.randn((S, n), device="cuda", generator=g)).clamp_(-1, 1)
    with BatchFirEngine([1.0], n_streams=1, max_frames=1) as e:
        y = e.agc_block(x, window_size=512, clip=0.99)
        y_noclip = e.agc_block(x, window_size=512, clip=0.0)
    assert float(y.abs().max()) <= float(np.float32(0.99))
    nz = x.abs() > 1e-6
    ratio = (y_noclip[nz] / x[nz])
    assert float(ratio.min()) >= 0.1 * (1 - 1e-6) and float(ratio.max()) <= 10.0 * (1 + 1e-6)
    xs, ys = x.cpu().numpy(), y.cpu().numpy()
    for s in list(range(0, 64)) + [4095, 8191, 16383]:
        assert_parity(ys[s], oagc.agc_stage(xs[s]), f"stream {s}")
