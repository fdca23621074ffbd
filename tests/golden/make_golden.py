"""Generate tests/golden/reference_golden.npz from the REAL reference files.

Run in the build container only (``/root/reference`` must exist):

    python tests/golden/make_golden.py

It imports, by path and unmodified, the reference modules on the hot path that are importable
without librosa -- ``aegis_engine_core/vision.py``, ``aegis_engine_core_v2/financial_filters.py``,
``aegis_engine_core_v2/financial_analysis.py`` -- and ``aegis_engine_core/midi_logic.py`` behind a
three-function ``librosa`` shim (``hz_to_midi``, ``amplitude_to_db``, and a ``util.softmask`` that
rejects the ``margin=`` keyword exactly like real librosa does, so ``midi_logic.py:47-49`` takes its
except-branch) and a ``mido`` stub.  Inputs are seeded; inputs and the reference's outputs are stored
side by side so the tests need nothing but the .npz.
"""
from __future__ import annotations

import importlib.util
import io
import contextlib
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import librosa_ref as L  # noqa: E402
import spectrogram_midi_b200  # noqa: E402,F401
from spectrogram_midi_b200 import corpus  # noqa: E402


def _load(name, path, package=None):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    if package:
        mod.__package__ = package
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _install_shims():
    lib = types.ModuleType("librosa")
    lib.hz_to_midi = L.hz_to_midi
    lib.midi_to_hz = L.midi_to_hz
    lib.note_to_hz = L.note_to_hz
    lib.amplitude_to_db = L.amplitude_to_db
    lib.power_to_db = L.power_to_db
    util = types.ModuleType("librosa.util")

    def softmask(X, X_ref, *, power=1, split_zeros=False):  # no `margin` kwarg -> TypeError, as in librosa
        raise AssertionError("softmask should never be reached with a valid signature here")

    util.softmask = softmask
    lib.util = util
    sys.modules["librosa"] = lib
    sys.modules["librosa.util"] = util
    mido = types.ModuleType("mido")
    mido.Message = type("Message", (), {})
    sys.modules["mido"] = mido


def main():
    _install_shims()
    vision = _load("ref_vision", f"{REF}/aegis_engine_core/vision.py")
    pkg = types.ModuleType("refv2")
    pkg.__path__ = [f"{REF}/aegis_engine_core_v2"]
    sys.modules["refv2"] = pkg
    filters = _load("refv2.financial_filters", f"{REF}/aegis_engine_core_v2/financial_filters.py", "refv2")
    analysis = _load("refv2.financial_analysis", f"{REF}/aegis_engine_core_v2/financial_analysis.py", "refv2")
    midi_logic = _load("ref_midi_logic", f"{REF}/aegis_engine_core/midi_logic.py")

    out = {}
    rng = np.random.default_rng(1234)

    # ---- vision.detect_rake_patterns -------------------------------------------------------
    cases = []
    for sr in (22050, 44100):
        y = corpus.test_track(sr, seed=0)
        S_dB = L.load_audio_features(y, sr)
        cases.append((f"track{sr}", S_dB, sr, 0.6))
    y = corpus.benchmark_signal(22050, seed=0)
    cases.append(("bench22050", L.load_audio_features(y, 22050), 22050, 0.5))
    # synthetic dB images with runs of 1, 2, 3 and 4 broadband columns, one still open at the end
    for sr in (22050, 44100):
        img = rng.uniform(-80, -45, size=(128, 64)).astype(np.float32)
        for s, n in [(5, 1), (12, 2), (20, 3), (30, 4), (62, 2)]:
            img[:, s : s + n] = rng.uniform(-12, 0, size=(128, n)).astype(np.float32)
        img[:, 40] = -70.0  # below the -60 gate
        cases.append((f"synthetic{sr}", img, sr, 0.6))
    for name, S_dB, sr, ratio in cases:
        out[f"rake/{name}/S_dB"] = S_dB.astype(np.float32)
        out[f"rake/{name}/args"] = np.array([512, sr, ratio], dtype=np.float64)
        out[f"rake/{name}/mask"] = vision.detect_rake_patterns(S_dB, 512, sr, ratio)

    # ---- financial filters / analysis ----------------------------------------------------
    def series(n, gap_prob, seed):
        r = np.random.default_rng(seed)
        base = 110.0 * 2 ** (np.cumsum(r.normal(0, 0.02, n)) / 12 + r.integers(0, 24) / 12)
        base += r.normal(0, 0.8, n)
        voiced = np.ones(n, bool)
        i = 0
        while i < n:
            if r.random() < gap_prob:
                g = int(r.integers(1, 25))
                voiced[i : i + g] = False
                i += g
            i += int(r.integers(1, 40))
        base[~voiced] = np.nan
        return base

    f0_cases = {
        "dense": series(431, 0.0, 1),
        "gappy": series(431, 0.5, 2),
        "long": series(1292, 0.3, 3),
        "short": series(11, 0.0, 4),         # == window: savgol skipped -> all-NaN (shorter than 10 makes the reference SMA raise)
        "allnan": np.full(40, np.nan),
        "onevalid": np.where(np.arange(30) == 7, 220.0, np.nan),
        "leading_gap": np.concatenate([np.full(15, np.nan), series(100, 0.2, 5)]),
    }
    an = analysis.FinancialPitchAnalyzer(sr=22050, hop_length=512)
    for name, f0 in f0_cases.items():
        k = f"filt/{name}"
        out[f"{k}/f0"] = f0
        F = filters.FinancialNoiseFilters
        out[f"{k}/savgol"] = np.asarray(F.savitzky_golay(f0.copy()), dtype=np.float64)
        out[f"{k}/kalman"] = np.asarray(F.kalman_filter(f0.copy()), dtype=np.float64)
        out[f"{k}/holt"] = np.asarray(F.holt_winters(f0.copy()), dtype=np.float64)
        with contextlib.redirect_stdout(io.StringIO()):
            c, conf = filters.multi_filter_consensus(f0.copy())
        out[f"{k}/consensus"] = np.asarray(c, dtype=np.float64)
        out[f"{k}/consensus_conf"] = np.asarray(conf, dtype=np.float64)
        out[f"{k}/sma5"] = an.simple_moving_average(f0.copy(), 5)
        out[f"{k}/sma10"] = an.simple_moving_average(f0.copy(), 10)
        out[f"{k}/ema5"] = an.exponential_moving_average(f0.copy(), 5)
        ma, up, lo = an.bollinger_bands(f0.copy(), window=10)
        out[f"{k}/boll_ma"], out[f"{k}/boll_up"], out[f"{k}/boll_lo"] = ma, up, lo
        m, s, h = an.macd(f0.copy())
        out[f"{k}/macd"], out[f"{k}/macd_signal"], out[f"{k}/macd_hist"] = m, s, h
        res = an.analyze_pitch_financial(f0.copy(), ~np.isnan(f0))
        out[f"{k}/an_trend"] = np.asarray(res["trend"], dtype=np.float64)
        out[f"{k}/an_conf"] = np.asarray(res["confidence"], dtype=np.float64)
        at, am = F.atr_filter(f0.copy())
        out[f"{k}/atr"], out[f"{k}/atr_mask"] = np.asarray(at, dtype=np.float64), np.asarray(am)
        out[f"{k}/ichimoku"] = np.asarray(F.ichimoku_baseline(f0.copy()), dtype=np.float64)
        out[f"{k}/stochastic"] = np.asarray(F.stochastic_oscillator(f0.copy()), dtype=np.float64)

    # ---- midi_logic.get_midi_events on oracle perception outputs --------------------------
    for name, y, sr in [("track22050", corpus.test_track(22050, 0, 10.0), 22050),
                        ("track44100", corpus.test_track(44100, 0), 44100),
                        ("clip7", corpus.random_clip(7, 12.0, 22050), 22050)]:
        S_dB = L.load_audio_features(y, sr)
        mask = vision.detect_rake_patterns(S_dB, 512, sr, 0.6)
        f0, vf, vp = L.pyin(y, fmin=L.note_to_hz("E2"), fmax=L.note_to_hz("C6"), sr=sr, hop_length=512)
        f0 = np.nan_to_num(f0)
        rms = L.rms(y, hop_length=512)[0]
        with contextlib.redirect_stdout(io.StringIO()):
            ev = midi_logic.get_midi_events(rake_mask=mask, f0=f0, voiced_flag=vf, active_probs=vp, rms=rms,
                                            sr=sr, hop_length=512, confidence_threshold=0.70)
        tech = {None: 0, "vibrato": 1, "bend": 2, "slide": 3, "hammer_on": 4, "pull_off": 5}
        k = f"midi/{name}"
        out[f"{k}/rake_mask"], out[f"{k}/f0"], out[f"{k}/voiced_flag"] = mask, f0, vf
        out[f"{k}/voiced_prob"], out[f"{k}/rms"] = vp, rms
        out[f"{k}/sr"] = np.array([sr])
        out[f"{k}/events"] = np.array(
            [[e["note"], e["start"], e["end"], e["velocity"], int(e["track"] == "main"), tech[e.get("technique")]] for e in ev],
            dtype=np.int64).reshape(-1, 6)
        out[f"{k}/event_float"] = np.array([[e["confidence"], e["rms_energy"], e.get("slope", 0.0)] for e in ev],
                                           dtype=np.float64).reshape(-1, 3)

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_golden.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
