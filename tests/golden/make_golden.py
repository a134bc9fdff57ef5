#!/usr/bin/env python
"""Generate the golden fixtures in this directory by EXECUTING THE REFERENCE'S OWN CODE.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

What is executed, unmodified, from /root/reference:
  * fir_filter.create_fir_filter / filter_methods / window_types       -> design_*.npz
  * stream_process.audio_callback                                      -> cb_stream_process_*.npz
  * stream_process_EQ.audio_callback                                   -> cb_stream_process_eq_*.npz
  * stream_process_GUI_Presets.OverlapAddFilter / process_audio_block  -> ola_*.npz
  * stream_process_GUI_Waterfall.make_audio_callback                   -> cb_waterfall_*.npz
  * stream_process_EQ_GUI.process_audio (worker thread) + min-phase call sites -> eqgui_*.npz
  * stream_process_AGC.apply_agc (+ the clip of its call site)         -> agc_blocks.npz
  * stream_process_GUI_Waterfall's waterfall row (show_waterfall on)   -> waterfall_spectrum.npz

What is stubbed (absent from the container and outside the hot path): sounddevice, tkinter,
matplotlib, plot_filter, samplerate, and the import-time side effects (sudo cpufreq-set,
sched_setaffinity, RLIMIT_MEMLOCK).  `soxr` is absent too and IS on the path; it is replaced by the
builder-defined streaming polyphase interpolator (oracle.reference_chain.PolyphaseUpsampler), so
every fixture also records the up-sampled blocks that the reference code actually consumed - the
FIR/EQ stages are pinned, the up-sampler is not.  `filter_methods.remez` is injected
from scipy because the reference forgets the import.

Versions are recorded in MANIFEST.json.
"""
from __future__ import annotations

import json
import os
import queue
import sys
import threading
import types

import numpy as np
import scipy
import scipy.signal

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle.reference_chain import PolyphaseUpsampler  # noqa: E402


# ------------------------------------------------------------------ stubs
class _Anything:
    """Callable / attribute sink used for GUI and audio-device objects."""

    def __init__(self, *a, **k):
        self._value = k.get("value")

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()

    def get(self):
        return self._value

    def set(self, v):
        self._value = v


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything


class _Soxr(types.ModuleType):
    """Stand-in for python-soxr: routes to per-(in_rate,out_rate) streaming interpolators that the
    generator installs in `self.table`; equal rates are the identity (float32)."""

    def __init__(self):
        super().__init__("soxr")
        self.table = {}
        self.log = []

    def resample(self, x, in_rate, out_rate, quality="HQ"):
        x = np.asarray(x, dtype=np.float32)
        if int(in_rate) == int(out_rate):
            y = x.copy()
        else:
            y = self.table[(int(in_rate), int(out_rate))].step(x)
        self.log.append(y.copy())
        return y


def install_stubs():
    soxr = _Soxr()
    sys.modules["soxr"] = soxr
    for name in ("sounddevice", "tkinter", "tkinter.ttk", "matplotlib", "matplotlib.pyplot",
                 "matplotlib.backends", "matplotlib.backends.backend_tkagg", "samplerate"):
        sys.modules[name] = _StubModule(name)
    pf = types.ModuleType("plot_filter")
    pf.plot_filter_response = lambda *a, **k: None
    sys.modules["plot_filter"] = pf
    import resource
    os.system = lambda *a, **k: 0
    os.sched_setaffinity = lambda *a, **k: None
    resource.setrlimit = lambda *a, **k: None
    sys.path.insert(0, REF)
    import filter_methods
    filter_methods.remez = scipy.signal.remez      # SURVEY.md F3 shim, reference file untouched
    return soxr


# ------------------------------------------------------------------ inputs
def sweep_and_noise(n, fs, seed):
    t = np.arange(n) / fs
    sweep = 0.1 * scipy.signal.chirp(t, 20.0, n / fs, min(20000.0, 0.45 * fs), method="logarithmic",
                                     phi=90 * (seed % 4))
    noise = 0.1 * np.random.default_rng(1234 + seed).standard_normal(n)
    return (0.5 * sweep + 0.5 * np.clip(noise, -1, 1)).astype(np.float32)


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"  wrote {name}.npz  ({os.path.getsize(path) / 1024:.1f} KiB)")


# ------------------------------------------------------------------ design fixtures
def gen_design():
    from fir_filter import create_fir_filter
    cases = []

    def add(**kw):
        taps = create_fir_filter(**kw)
        cases.append((kw, taps))

    # BASELINE.json configs and the scripts' own filters
    add(method="window", cutoff=[500, 8000], numtaps=251, window_type="kaiser",
        filter_type="bandpass", samplerate=176400)                       # C1 / README.md:189-195
    add(method="window", cutoff=11000, numtaps=301, window_type="hamming",
        filter_type="lowpass", samplerate=176400)                        # stream_process.py:19-40
    add(method="remez", cutoff=20000, numtaps=1001, window_type="hamming",
        filter_type="lowpass", samplerate=176400)                        # C2
    add(method="window", cutoff=1000, numtaps=501, window_type="boxcar",
        filter_type="highpass", samplerate=88200)                        # C4 (gate passes)
    add(method="window", cutoff=16000, numtaps=257, window_type="hamming",
        filter_type="lowpass", samplerate=48000)                         # C3 main
    add(method="window", cutoff=[300, 3000], numtaps=101, window_type="blackman",
        filter_type="bandstop", samplerate=44100)
    add(method="window", cutoff=5000, numtaps=64, window_type="hann",
        filter_type="lowpass", samplerate=44100)                         # even length
    add()                                                                # all defaults
    for n in (51, 129, 501):
        for band in [(20, 60), (60, 120), (120, 250), (250, 500), (500, 1000), (1000, 2000),
                     (2000, 4000), (4000, 8000), (8000, 16000), (16000, 20000)]:
            add(method="window", cutoff=band, numtaps=n, window_type="hamming",
                filter_type="bandpass", samplerate=48000)
    for w in ("boxcar", "hamming", "hann", "blackman", "kaiser", "bartlett", "flattop", "nuttall",
              "blackmanharris", "bohman", "parzen", "cosine", "barthann"):
        add(method="window", cutoff=4000, numtaps=63, window_type=w, filter_type="lowpass",
            samplerate=44100)
    add(method="remez", cutoff=3000, numtaps=101, filter_type="highpass", samplerate=44100)
    # remez band-pass / band-stop: filter_methods.py:86-91 builds 4 band edges for 3 gains, which
    # scipy rejects with ValueError even with the import repaired - recorded as a behaviour.
    raised = []
    for ft in ("bandpass", "bandstop"):
        try:
            create_fir_filter(method="remez", cutoff=[2000, 6000], numtaps=151, filter_type=ft,
                              samplerate=44100)
            raised.append("none")
        except Exception as exc:  # noqa: BLE001
            raised.append(type(exc).__name__)

    arrays = {}
    meta = []
    for i, (kw, taps) in enumerate(cases):
        arrays[f"taps_{i}"] = taps
        meta.append({k: (list(v) if isinstance(v, (tuple, list)) else v) for k, v in kw.items()})
    arrays["meta"] = np.array(json.dumps(meta))
    arrays["remez_band_errors"] = np.array(json.dumps(raised))
    save("design_taps", **arrays)


def gen_min_phase():
    import stream_process_EQ_GUI as g
    from fir_filter import create_fir_filter
    out = {}
    meta = []
    specs = [
        dict(method="window", cutoff=1000, numtaps=501, window_type="boxcar",
             filter_type="highpass", samplerate=88200),                  # C4
        dict(method="window", cutoff=14000, numtaps=129, window_type="hamming",
             filter_type="lowpass", samplerate=88200),                   # gate fails (periodic window)
        dict(method="remez", cutoff=20000, numtaps=301, filter_type="lowpass", samplerate=176400),
    ]
    for i, kw in enumerate(specs):
        h = create_fir_filter(**kw)
        sym = bool(g.is_symmetric(h))
        res = h
        if sym:                                                          # stream_process_EQ_GUI.py:377-379
            res = g.minimum_phase(h, method="hilbert")
            res = g.normalize_filter(res, kw["samplerate"])
        out[f"in_{i}"] = h
        out[f"out_{i}"] = res
        meta.append(dict(kw, symmetric=sym))
    out["meta"] = np.array(json.dumps(meta))
    save("min_phase", **out)


# ------------------------------------------------------------------ callback fixtures
def _run_callback(mod, cb, x, B, seed):
    """Drive a PortAudio-style callback over x (multiple of B frames).  Returns outdata blocks."""
    nblk = len(x) // B
    outs = np.zeros((nblk, B), dtype=np.float32)
    for i in range(nblk):
        indata = x[i * B:(i + 1) * B].reshape(B, 1).copy()
        outdata = np.zeros((B, 1), dtype=np.float32)
        np.random.seed(seed + i)
        cb(indata, outdata, B, None, None)
        outs[i] = outdata[:, 0]
    return outs


def gen_stream_process(soxr):
    import stream_process as sp
    from fir_filter import create_fir_filter
    from audio_filtering_pipeline_b200.design import design_upsampler
    for tag, (N, B, U, fs, nblk, design) in {
        "c1": (251, 1024, 4, 44100, 6, dict(method="window", cutoff=[500, 8000], numtaps=251,
                                              window_type="kaiser", filter_type="bandpass")),
        "default": (301, 4096, 4, 44100, 2, dict(method="window", cutoff=11000, numtaps=301,
                                                   window_type="hamming", filter_type="lowpass")),
        "c2": (1001, 512, 1, 176400, 8, dict(method="remez", cutoff=20000, numtaps=1001,
                                              filter_type="lowpass")),
        "short": (51, 1024, 2, 48000, 4, dict(method="window", cutoff=9000, numtaps=51,
                                               window_type="hamming", filter_type="lowpass")),
    }.items():
        taps = create_fir_filter(samplerate=fs * U, **design)
        up_taps = design_upsampler(U, fs)
        sp.SAMPLERATE, sp.UPSAMPLE_FACTOR, sp.UPSAMPLE_RATE = fs, U, fs * U
        sp.BLOCKSIZE, sp.NUM_TAPS = B, N
        sp.fir_coeff = taps
        sp.input_buffer = np.zeros(N + B * U - 1, dtype=np.float32)      # stream_process.py:45-46
        sp.audio_buffer.clear()
        sp.silence_block = np.zeros(B, dtype=np.float32)
        soxr.table = {(fs, fs * U): PolyphaseUpsampler(up_taps, U)}
        soxr.log = []
        x = sweep_and_noise(nblk * B, fs, seed=len(tag))
        pre = []
        orig_append = sp.audio_buffer.append
        outs = np.zeros((nblk, B), dtype=np.float32)
        for i in range(nblk):
            indata = x[i * B:(i + 1) * B].reshape(B, 1).copy()
            outdata = np.zeros((B, 1), dtype=np.float32)
            np.random.seed(100 + i)
            sp.audio_callback(indata, outdata, B, None, None)
            pre.append(np.array(sp.audio_buffer[-1], dtype=np.float64))  # 'downsampled', :110
            outs[i] = outdata[:, 0]
        del orig_append
        save(f"cb_stream_process_{tag}", x=x, taps=taps, up_taps=up_taps,
             upsampled=np.stack(soxr.log).astype(np.float32), pre_dither=np.stack(pre),
             outdata=outs, params=np.array([N, B, U, fs]))


def gen_stream_process_eq(soxr):
    import stream_process_EQ as sp
    from fir_filter import create_fir_filter
    from audio_filtering_pipeline_b200.design import design_upsampler
    for tag, (N, B, U, fs, nblk) in {"n101_u2": (101, 512, 2, 44100, 5),
                                     "n301_u4": (301, 1024, 4, 44100, 3)}.items():
        taps = create_fir_filter(method="window", cutoff=11000, numtaps=N, window_type="hamming",
                                 filter_type="lowpass", samplerate=fs * U)
        eq = []
        for (band, gain) in sp.EQ_BANDS:                                  # stream_process_EQ.py:33-37
            c = create_fir_filter(method="window", cutoff=band, numtaps=N, window_type="hamming",
                                  filter_type="bandpass", samplerate=fs * U)
            eq.append((c, gain))
        up_taps = design_upsampler(U, fs)
        sp.SAMPLERATE, sp.UPSAMPLE_FACTOR, sp.UPSAMPLE_RATE = fs, U, fs * U
        sp.BLOCKSIZE, sp.NUM_TAPS = B, N
        sp.fir_coeff, sp.eq_filters = taps, eq
        sp.input_buffer = np.zeros(N + B * U - 1, dtype=np.float32)
        sp.audio_buffer.clear()
        sp.silence_block = np.zeros(B, dtype=np.float32)
        soxr.table = {(fs, fs * U): PolyphaseUpsampler(up_taps, U)}
        soxr.log = []
        x = sweep_and_noise(nblk * B, fs, seed=7)
        pre = []
        outs = np.zeros((nblk, B), dtype=np.float32)
        for i in range(nblk):
            indata = x[i * B:(i + 1) * B].reshape(B, 1).copy()
            outdata = np.zeros((B, 1), dtype=np.float32)
            np.random.seed(200 + i)
            sp.audio_callback(indata, outdata, B, None, None)
            pre.append(np.array(sp.audio_buffer[-1], dtype=np.float64))
            outs[i] = outdata[:, 0]
        save(f"cb_stream_process_eq_{tag}", x=x, taps=taps, up_taps=up_taps,
             eq_taps=np.stack([c for c, _ in eq]), eq_gains=np.array([g for _, g in eq]),
             upsampled=np.stack(soxr.log).astype(np.float32), pre_dither=np.stack(pre),
             outdata=outs, params=np.array([N, B, U, fs]))


def gen_stream_process_eq_c3(soxr):
    """BASELINE.json configs[2] at full size through the same callback (stream_process_EQ.audio_callback): 10 bands
    (the GUIs' 9, stream_process_EQ_GUI.py:162-172, + 16-20 kHz) x 501 taps, 501-tap main low-pass, 48 kHz, U = 1,
    1024-frame blocks, stereo = the mono callback run once per channel."""
    import stream_process_EQ as sp
    from fir_filter import create_fir_filter
    N, B, U, fs, nblk = 501, 1024, 1, 48000, 4
    bands = [(20, 60), (60, 120), (120, 250), (250, 500), (500, 1000), (1000, 2000), (2000, 4000), (4000, 8000),
             (8000, 16000), (16000, 20000)]
    gains = [1.2, 1.0, 0.8, 1.0, 1.5, 1.0, 0.5, 1.0, 1.3, 0.9]
    taps = create_fir_filter(method="window", cutoff=16000, numtaps=N, window_type="hamming", filter_type="lowpass",
                             samplerate=fs * U)
    eq = [(create_fir_filter(method="window", cutoff=b, numtaps=N, window_type="hamming", filter_type="bandpass",
                             samplerate=fs * U), g) for b, g in zip(bands, gains)]
    xs, pres, outs_all = [], [], []
    for ch in range(2):
        sp.SAMPLERATE, sp.UPSAMPLE_FACTOR, sp.UPSAMPLE_RATE = fs, U, fs * U
        sp.BLOCKSIZE, sp.NUM_TAPS = B, N
        sp.fir_coeff, sp.eq_filters = taps, eq
        sp.input_buffer = np.zeros(N + B * U - 1, dtype=np.float32)
        sp.audio_buffer.clear()
        sp.silence_block = np.zeros(B, dtype=np.float32)
        soxr.table = {}
        soxr.log = []
        x = sweep_and_noise(nblk * B, fs, seed=20 + ch)
        pre, outs = [], np.zeros((nblk, B), dtype=np.float32)
        for i in range(nblk):
            indata = x[i * B:(i + 1) * B].reshape(B, 1).copy()
            outdata = np.zeros((B, 1), dtype=np.float32)
            np.random.seed(500 + 10 * ch + i)
            sp.audio_callback(indata, outdata, B, None, None)
            pre.append(np.array(sp.audio_buffer[-1], dtype=np.float64))
            outs[i] = outdata[:, 0]
        xs.append(x); pres.append(np.stack(pre)); outs_all.append(outs)
    save("cb_stream_process_eq_n501_10band", x=np.stack(xs), taps=taps, eq_taps=np.stack([c for c, _ in eq]),
         eq_gains=np.array(gains), pre_dither=np.stack(pres), outdata=np.stack(outs_all),
         params=np.array([N, B, U, fs]))


def gen_overlap_add():
    import stream_process_GUI_Presets as pr
    from fir_filter import create_fir_filter
    rng = np.random.default_rng(5)
    # (a) single filter, N-1 <= L (correct regime) and N-1 > L (reference defect)
    for tag, (N, L, nblk) in {"n129_l2048": (129, 2048, 4), "n1001_l512": (1001, 512, 6),
                              "n51_l256": (51, 256, 6)}.items():
        taps = create_fir_filter(method="window", cutoff=9000, numtaps=N, window_type="hamming",
                                 filter_type="lowpass", samplerate=88200)
        flt = pr.OverlapAddFilter(taps, L)
        x = (0.1 * rng.standard_normal(nblk * L)).astype(np.float32)
        y = np.stack([flt.process_block(x[i * L:(i + 1) * L]) for i in range(nblk)])
        # wrong-sized blocks (:71-76)
        flt2 = pr.OverlapAddFilter(taps, L)
        y_short = flt2.process_block(x[: L // 2])
        y_long = flt2.process_block(x[: L + 17])
        save(f"ola_{tag}", x=x, taps=taps, y=y, y_short=y_short, y_long=y_long,
             params=np.array([N, L]))
    # (b) the operator chain: 9 EQ OverlapAddFilters + main, process_audio_block (:598-613)
    N, L, nblk, fsu = 129, 1024, 5, 88200
    bands = [(20, 60), (60, 120), (120, 250), (250, 500), (500, 1000), (1000, 2000),
             (2000, 4000), (4000, 8000), (8000, 16000)]
    gains = [1.2, 1.0, 0.8, 1.0, 1.5, 1.0, 0.5, 1.0, 1.3]
    eq_taps = [create_fir_filter(method="window", cutoff=b, numtaps=N, window_type="hamming",
                                 filter_type="bandpass", samplerate=fsu) for b in bands]
    main = create_fir_filter(method="window", cutoff=14000, numtaps=N, window_type="hamming",
                             filter_type="lowpass", samplerate=fsu)
    gui = types.SimpleNamespace(
        eq_filter_objects=[pr.OverlapAddFilter(c, L) for c in eq_taps],
        eq_gains=[_Anything(value=g) for g in gains],
        fir_filter_object=pr.OverlapAddFilter(main, L))
    x = (0.05 * rng.standard_normal(nblk * L)).astype(np.float32)
    y = np.stack([pr.process_audio_block(gui, x[i * L:(i + 1) * L]) for i in range(nblk)])
    save("ola_chain_9band", x=x, eq_taps=np.stack(eq_taps), eq_gains=np.array(gains), taps=main,
         y=y, params=np.array([N, L]))
    # empty coefficient list -> identity (:45-46)
    ident = pr.OverlapAddFilter([], 64)
    xi = (0.1 * rng.standard_normal(64)).astype(np.float32)
    save("ola_identity", x=xi, y=ident.process_block(xi), coeffs=ident.coeffs)


def gen_waterfall(soxr):
    import stream_process_GUI_Waterfall as wf
    from fir_filter import create_fir_filter
    N, B, U, fs, nblk = 257, 1024, 1, 44100, 6                          # GUI defaults (:103-127)
    bands = [(20, 60), (60, 120), (120, 250), (250, 500), (500, 1000), (1000, 2000),
             (2000, 4000), (4000, 8000), (8000, 16000)]
    gains = [1.0, 1.1, 0.9, 1.0, 1.4, 1.0, 0.6, 1.0, 1.2]
    eq_taps = [create_fir_filter(method="window", cutoff=b, numtaps=N, window_type="hamming",
                                 filter_type="bandpass", samplerate=fs * U) for b in bands]
    main = create_fir_filter(method="window", cutoff=14000, numtaps=N, window_type="hamming",
                             filter_type="lowpass", samplerate=fs * U)
    buf = np.zeros(B * 4 + 512, dtype=np.float32)                        # :132
    gui = types.SimpleNamespace(
        eq_filters=eq_taps, eq_gains=[_Anything(value=g) for g in gains], fir_coeff=main,
        input_buffer=buf, eq_output=np.zeros_like(buf), audio_buffer=[],
        silence_block=np.zeros(B, dtype=np.float32), last_output=np.zeros(B, dtype=np.float32),
        show_waterfall=_Anything(value=False), waterfall_data=[],
        get_gains=lambda: gains, get_dsp_config=lambda: (fs, U, B))
    cb = wf.make_audio_callback(gui)
    x = sweep_and_noise(nblk * B, fs, seed=3)
    pre = []
    outs = np.zeros((nblk, B), dtype=np.float32)
    for i in range(nblk):
        indata = x[i * B:(i + 1) * B].reshape(B, 1).copy()
        outdata = np.zeros((B, 1), dtype=np.float32)
        np.random.seed(300 + i)
        cb(indata, outdata, B, None, None)
        pre.append(np.array(gui.last_output, dtype=np.float64))
        outs[i] = outdata[:, 0]
    save("cb_waterfall_9band", x=x, taps=main, eq_taps=np.stack(eq_taps), eq_gains=np.array(gains),
         pre_dither=np.stack(pre), outdata=outs, params=np.array([N, B, U, fs]))


def gen_waterfall_spectrum(soxr):
    """The waterfall row of the same callback (stream_process_GUI_Waterfall.py:471-476): Hann-windowed
    rfft of the block just written to outdata, 20*log10(|X| + 1e-12).  show_waterfall = True."""
    import stream_process_GUI_Waterfall as wf
    from fir_filter import create_fir_filter
    N, B, U, fs, nblk = 129, 1024, 1, 44100, 4
    bands = [(20, 60), (60, 120), (120, 250), (250, 500), (500, 1000), (1000, 2000),
             (2000, 4000), (4000, 8000), (8000, 16000)]
    gains = [1.0] * 9
    eq_taps = [create_fir_filter(method="window", cutoff=b, numtaps=N, window_type="hamming",
                                 filter_type="bandpass", samplerate=fs * U) for b in bands]
    main = create_fir_filter(method="window", cutoff=14000, numtaps=N, window_type="hamming",
                             filter_type="lowpass", samplerate=fs * U)
    buf = np.zeros(B * 4 + 512, dtype=np.float32)
    gui = types.SimpleNamespace(
        eq_filters=eq_taps, eq_gains=[_Anything(value=g) for g in gains], fir_coeff=main,
        input_buffer=buf, eq_output=np.zeros_like(buf), audio_buffer=[],
        silence_block=np.zeros(B, dtype=np.float32), last_output=np.zeros(B, dtype=np.float32),
        show_waterfall=_Anything(value=True), waterfall_data=[],
        get_gains=lambda: gains, get_dsp_config=lambda: (fs, U, B))
    cb = wf.make_audio_callback(gui)
    x = sweep_and_noise(nblk * B, fs, seed=4)
    outs = np.zeros((nblk, B), dtype=np.float32)
    for i in range(nblk):
        indata = x[i * B:(i + 1) * B].reshape(B, 1).copy()
        outdata = np.zeros((B, 1), dtype=np.float32)
        np.random.seed(400 + i)
        cb(indata, outdata, B, None, None)
        outs[i] = outdata[:, 0]
    save("waterfall_spectrum", outdata=outs, spectrum_db=np.stack(gui.waterfall_data), params=np.array([B, fs]))


def gen_eq_gui_worker():
    """stream_process_EQ_GUI.process_audio is an endless worker loop fed by two queues (:65-113);
    run it on a daemon thread for a handful of blocks at U=1 (resamplers bypassed, :422-423)."""
    import stream_process_EQ_GUI as g
    from fir_filter import create_fir_filter
    N, B, U, fs, nblk = 129, 2048, 1, 44100, 3
    bands = [(20, 60), (60, 120), (120, 250), (250, 500), (500, 1000), (1000, 2000),
             (2000, 4000), (4000, 8000), (8000, 16000)]
    gains = [1.3, 1.0, 0.7, 1.0, 1.1, 1.0, 0.9, 1.0, 1.25]
    eq_taps = [create_fir_filter(method="window", cutoff=b, numtaps=N, window_type="hamming",
                                 filter_type="bandpass", samplerate=fs * U) for b in bands]
    main = create_fir_filter(method="window", cutoff=14000, numtaps=N, window_type="hamming",
                             filter_type="lowpass", samplerate=fs * U)
    gui = types.SimpleNamespace(eq_filters=eq_taps, get_gains=lambda: gains,
                                get_dsp_config=lambda: (fs, U, B))
    g.active_filter_params = main.copy()
    g.active_eq_filters = list(eq_taps)
    g.audio_queue = queue.Queue(maxsize=20)
    g.processed_queue = queue.Queue(maxsize=20)
    th = threading.Thread(target=g.process_audio, args=(gui,), daemon=True)
    th.start()
    x = sweep_and_noise(nblk * B, fs, seed=9)
    ys = []
    for i in range(nblk):
        g.audio_queue.put(x[i * B:(i + 1) * B].reshape(B, 1).copy())
        ys.append(np.array(g.processed_queue.get(timeout=30), dtype=np.float64))
    save("eqgui_worker_stateless", x=x, taps=main, eq_taps=np.stack(eq_taps),
         eq_gains=np.array(gains), y=np.stack(ys), params=np.array([N, B, U, fs]))


def gen_agc():
    """stream_process_AGC.apply_agc (:43-89) on single blocks, called the way the worker loop calls it
    (:131-143: [n, 1] float32 block, window min(512, bs // 2), then np.clip +-0.99), plus direct calls
    with other windows - among them windows short enough that int(attack * window) == 0, where the
    reference divides by zero inside its loop, catches the exception and returns the block unchanged."""
    import contextlib
    import io
    import stream_process_AGC as ag
    rng = np.random.default_rng(11)
    cases = []

    def level_steps(n, levels):
        seg = n // len(levels)
        env = np.repeat(np.asarray(levels, dtype=np.float64), seg)
        env = np.pad(env, (0, n - len(env)), mode="edge")
        return (env * rng.standard_normal(n)).astype(np.float32)

    blocks = {
        "noise_1024": (0.3 * rng.standard_normal(1024)).astype(np.float32),
        "steps_1024": level_steps(1024, [0.5, 0.002, 0.2, 0.02]),
        "quiet_1024": (1e-4 * rng.standard_normal(1024)).astype(np.float32),
        "sweep_2048": sweep_and_noise(2048, 44100, seed=2),
        "steps_4096": level_steps(4096, [0.01, 0.9, 0.05, 0.4, 0.001, 0.3]),
        "odd_777": level_steps(777, [0.2, 0.01, 0.6]),
        "silence_512": np.zeros(512, dtype=np.float32),
        "loud_1024": np.clip(2.5 * rng.standard_normal(1024), -4, 4).astype(np.float32),
    }
    arrays, meta = {}, []
    for name, x in blocks.items():
        n = len(x)
        variants = [dict(target_level=0.1, window_size=min(512, n // 2), max_gain=10.0, attack=0.01, release=0.1),
                    dict(target_level=0.05, window_size=256, max_gain=4.0, attack=0.02, release=0.3),
                    dict(target_level=0.2, window_size=101, max_gain=20.0, attack=0.1, release=0.01),   # odd window, attack slower than release
                    dict(target_level=0.1, window_size=64, max_gain=10.0, attack=0.01, release=0.1),    # int(0.64) == 0 -> error path
                    dict(target_level=0.1, window_size=n, max_gain=10.0, attack=0.01, release=0.1)]     # window == block
        for kw in variants:
            with contextlib.redirect_stdout(io.StringIO()) as log:
                y = ag.apply_agc(x.reshape(-1, 1).copy(), **kw)
            key = f"{name}_{len(meta)}"
            arrays["x_" + name] = x
            arrays["y_" + key] = np.asarray(y)[:, 0]
            meta.append(dict(kw, key=key, block=name, errored=bool(log.getvalue().startswith("AGC error"))))
    arrays["meta"] = np.array(json.dumps(meta))
    save("agc_blocks", **arrays)


def main():
    soxr = install_stubs()
    if "--only-eq-c3" in sys.argv:      # added in round 2: leaves the other fixtures (and the manifest) untouched
        print("stream_process_EQ, C3 size ..."); gen_stream_process_eq_c3(soxr)
        return
    print("design ...");            gen_design()
    print("min phase ...");         gen_min_phase()
    print("stream_process ...");    gen_stream_process(soxr)
    print("stream_process_EQ ..."); gen_stream_process_eq(soxr)
    print("stream_process_EQ, C3 size ..."); gen_stream_process_eq_c3(soxr)
    print("overlap-add ...");       gen_overlap_add()
    print("waterfall ...");         gen_waterfall(soxr)
    print("EQ_GUI worker ...");     gen_eq_gui_worker()
    print("AGC ...");               gen_agc()
    print("waterfall spectrum ..."); gen_waterfall_spectrum(soxr)
    manifest = {
        "generated_by": "tests/golden/make_golden.py",
        "reference": REF,
        "numpy": np.__version__, "scipy": scipy.__version__,
        "python": sys.version.split()[0],
        "note": "outputs of the reference's own functions; soxr replaced by the builder's "
                "streaming polyphase interpolator (up-sampler parity unpinned)",
    }
    with open(os.path.join(HERE, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print("done")


if __name__ == "__main__":
    main()
