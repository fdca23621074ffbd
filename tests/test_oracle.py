"""CPU tests: the oracle against the committed golden vectors (made from the real reference files)
and against independent implementations / brute force for the librosa subset."""
import itertools
import warnings

import numpy as np
import pytest

from oracle import librosa_ref as L
from oracle import reference_files as R

import spectrogram_midi_b200  # noqa: F401
from spectrogram_midi_b200 import corpus


def _cases(golden, prefix):
    return sorted({k.split("/")[1] for k in golden if k.startswith(prefix + "/")})


# ------------------------------------------------------------------ pinned: reference's own files
def test_rake_mask_matches_reference(golden):
    names = _cases(golden, "rake")
    assert len(names) >= 5
    fired = 0
    for n in names:
        hop, sr, ratio = golden[f"rake/{n}/args"]
        got = R.detect_rake_patterns(golden[f"rake/{n}/S_dB"], int(hop), int(sr), float(ratio))
        np.testing.assert_array_equal(got, golden[f"rake/{n}/mask"], err_msg=n)
        fired += int(got.sum())
    assert fired > 0  # the fixtures actually exercise the positive branch


def test_guitar_filters_match_reference(guitar_golden):
    """oracle restatement of aegis_engine_core_v2/guitar_specific.py against the real file's outputs"""
    g = guitar_golden
    names = _cases(g, "guitar")
    assert len(names) >= 8
    code = {"clean": 0, "light": 1, "heavy": 2}
    mute = added = 0
    seen = set()
    for n in names:
        hop, sr = (int(v) for v in g[f"guitar/{n}/args"])
        res = R.apply_guitar_filters(g[f"guitar/{n}/f0"], g[f"guitar/{n}/voiced"], g[f"guitar/{n}/S_dB"], hop, sr, g[f"guitar/{n}/rake_in"])
        np.testing.assert_array_equal(res["f0"], g[f"guitar/{n}/out_f0"], err_msg=n)
        np.testing.assert_array_equal(res["voiced"], g[f"guitar/{n}/out_voiced"], err_msg=n)
        np.testing.assert_array_equal(res["rake_mask"], g[f"guitar/{n}/out_rake"], err_msg=n)
        np.testing.assert_array_equal(res["mute_mask"], g[f"guitar/{n}/out_mute"], err_msg=n)
        assert code[res["distortion"]] == int(g[f"guitar/{n}/out_distortion"][0]), n
        mute += int(res["mute_mask"].sum())
        added += int((res["rake_mask"] ^ g[f"guitar/{n}/rake_in"]).sum())
        seen.add(code[res["distortion"]])
    assert mute > 0 and added > 0 and seen == {0, 1, 2}  # every branch is exercised


def test_financial_logic_filter_matches_reference(fin_golden):
    """oracle restatement of aegis_engine_core_v2/midi_logic_financial.py (+ financial_analysis / harmonic_analysis)
    against the real files' events: every integer field, the confidences bit for bit, and the key"""
    from oracle import financial_events as FE

    g = fin_golden
    names = [str(n) for n in g["fin/names"]]
    assert len(names) >= 15
    total = keys = 0
    labels = set()
    for n in names:
        frames, sr, kw = FE.golden_case(g, n)
        ev = FE.get_midi_events_financial(*frames, sr, 512, **kw)
        ints, conf, key = FE.events_rows(ev)
        np.testing.assert_array_equal(ints, g[f"fin/{n}/events"], err_msg=n)
        np.testing.assert_array_equal(conf, g[f"fin/{n}/confidence"], err_msg=n)
        np.testing.assert_array_equal(np.array([-1.0, -1.0, 0.0] if key is None else key), g[f"fin/{n}/key"], err_msg=n)
        total += len(ev)
        keys += key is not None
        labels |= set(ints[:, 5].tolist())
    assert total > 200 and keys >= 4 and labels >= {1, 2, 3, 4}   # events, key detections and every label occur


@pytest.mark.parametrize("key,fn", [
    ("savgol", lambda f: R.savitzky_golay(f)),
    ("kalman", lambda f: R.kalman_filter(f)),
    ("holt", lambda f: R.holt_winters(f)),
    ("consensus", lambda f: R.multi_filter_consensus(f)[0]),
    ("consensus_conf", lambda f: R.multi_filter_consensus(f)[1]),
    ("sma5", lambda f: R.simple_moving_average(f, 5)),
    ("sma10", lambda f: R.simple_moving_average(f, 10)),
    ("ema5", lambda f: R.exponential_moving_average(f, 5)),
    ("boll_ma", lambda f: R.bollinger_bands(f, 10)[0]),
    ("boll_up", lambda f: R.bollinger_bands(f, 10)[1]),
    ("boll_lo", lambda f: R.bollinger_bands(f, 10)[2]),
    ("macd", lambda f: R.macd(f)[0]),
    ("macd_signal", lambda f: R.macd(f)[1]),
    ("macd_hist", lambda f: R.macd(f)[2]),
    ("an_conf", lambda f: R.bollinger_confidence(f, 10)),
    ("atr", lambda f: R.atr_filter(f)[0]),
    ("atr_mask", lambda f: R.atr_filter(f)[1]),
    ("ichimoku", lambda f: R.ichimoku_baseline(f)),
    ("stochastic", lambda f: R.stochastic_oscillator(f)),
])
def test_filters_match_reference(golden, key, fn):
    for n in _cases(golden, "filt"):
        f0 = golden[f"filt/{n}/f0"]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            got = np.asarray(fn(f0.copy()))
        np.testing.assert_array_equal(got, golden[f"filt/{n}/{key}"], err_msg=f"{n}/{key}")


def test_midi_events_match_reference(golden):
    tech = {None: 0, "vibrato": 1, "bend": 2, "slide": 3, "hammer_on": 4, "pull_off": 5}
    total = 0
    for n in _cases(golden, "midi"):
        g = {k.split("/")[2]: v for k, v in golden.items() if k.startswith(f"midi/{n}/")}
        ev = R.get_midi_events(g["rake_mask"], g["f0"], g["voiced_flag"], g["voiced_prob"], g["rms"],
                               int(g["sr"][0]), 512, 0.70)
        got = np.array([[e["note"], e["start"], e["end"], e["velocity"], int(e["track"] == "main"),
                         tech[e.get("technique")]] for e in ev], dtype=np.int64).reshape(-1, 6)
        np.testing.assert_array_equal(got, g["events"], err_msg=n)
        fl = np.array([[e["confidence"], e["rms_energy"], e.get("slope", 0.0)] for e in ev]).reshape(-1, 3)
        np.testing.assert_allclose(fl, g["event_float"], rtol=0, atol=0)
        total += len(ev)
    assert total >= 10


# ------------------------------------------------------------------ librosa subset (unpinned upstream)
def test_unit_helpers():
    assert L.note_to_hz("E2") == 82.4068892282175
    assert L.note_to_hz("C6") == 1046.5022612023945
    assert L.note_to_hz("E6") == 1318.5102276514797
    assert L.n_pitch_bins_for(L.note_to_hz("E2"), L.note_to_hz("C6"))[0] == 441
    assert L.n_pitch_bins_for(L.note_to_hz("E2"), L.note_to_hz("E6"))[0] == 480  # floor trap, SURVEY §8 a-5
    assert L.pyin_periods(22050, L.note_to_hz("E2"), L.note_to_hz("C6")) == (21, 268)
    assert L.pyin_periods(44100, L.note_to_hz("E2"), L.note_to_hz("C6")) == (42, 536)
    assert abs(float(L.hz_to_midi(440.0)) - 69.0) < 1e-12


def test_mel_filterbank_against_independent_implementations():
    M = L.mel_filterbank(22050)
    assert M.shape == (128, 1025) and M.dtype == np.float32
    ta = pytest.importorskip("torchaudio")
    M2 = ta.functional.melscale_fbanks(1025, 0.0, 11025.0, 128, 22050, norm="slaney", mel_scale="slaney").T.numpy()
    np.testing.assert_allclose(M, M2, atol=5e-7)
    au = pytest.importorskip("transformers.audio_utils")
    M3 = au.mel_filter_bank(1025, 128, 0.0, 11025.0, 22050, norm="slaney", mel_scale="slaney").T
    np.testing.assert_allclose(M, M3, atol=1e-8)


def test_stft_against_scipy():
    import scipy.signal

    y = corpus.test_track(22050, 0)
    X = L.stft(y)
    assert X.shape == (1025, 1 + len(y) // 512) and X.dtype == np.complex64
    _, _, Z = scipy.signal.stft(y.astype(np.float64), window="hann", nperseg=2048, noverlap=1536,
                                boundary="zeros", padded=False, scaling="spectrum")
    n = min(X.shape[1], Z.shape[1])
    assert np.abs(X[:, :n] - Z[:, :n] * 1024.0).max() <= 1e-6 * np.abs(X).max()


def test_yin_difference_is_the_time_domain_definition():
    rng = np.random.default_rng(0)
    fr = rng.normal(size=(2048, 3)).astype(np.float32)
    d = L.yin_difference(fr, 2048, 1024)
    x = fr.astype(np.float64)
    for t in range(3):
        for tau in (0, 1, 21, 100, 268, 1023):
            ref = np.sum((x[1:1025, t] - x[1 + tau:1025 + tau, t]) ** 2)
            assert abs(d[tau, t] - ref) < 2e-2 * max(1.0, ref) ** 0.5 + 1e-2


def test_viterbi_matches_brute_force():
    rng = np.random.default_rng(5)
    n_states, n_steps = 4, 6
    for _ in range(20):
        prob = rng.random((n_states, n_steps))
        prob[rng.random(prob.shape) < 0.3] = 0.0
        trans = rng.random((n_states, n_states))
        trans[rng.random(trans.shape) < 0.3] = 0.0
        trans += np.eye(n_states) * 0.1
        trans /= trans.sum(axis=1, keepdims=True)
        p0 = np.full(n_states, 1 / n_states)
        got = L.viterbi(prob, trans, p0)
        best, best_path = -np.inf, None
        for path in itertools.product(range(n_states), repeat=n_steps):
            s = np.log(p0[path[0]] + L.TINY64) + np.log(prob[path[0], 0] + L.TINY64)
            for t in range(1, n_steps):
                s += np.log(trans[path[t - 1], path[t]] + L.TINY64) + np.log(prob[path[t], t] + L.TINY64)
            if s > best + 1e-9:
                best, best_path = s, path
        sc = np.log(p0[got[0]] + L.TINY64) + np.log(prob[got[0], 0] + L.TINY64)
        for t in range(1, n_steps):
            sc += np.log(trans[got[t - 1], got[t]] + L.TINY64) + np.log(prob[got[t], t] + L.TINY64)
        assert abs(sc - best) < 1e-9


def test_pyin_tracks_known_pitches():
    sr = 22050
    t = np.arange(int(sr * 1.0)) / sr
    for midi in (40, 52, 64, 76):
        f = float(L.midi_to_hz(midi))
        y = (0.5 * np.sin(2 * np.pi * f * t)).astype(np.float32)
        f0, vf, vp = L.pyin(y, fmin=L.note_to_hz("E2"), fmax=L.note_to_hz("C6"), sr=sr, hop_length=512)
        assert len(f0) == 1 + len(y) // 512
        mid = slice(5, -5)
        assert vf[mid].all()
        cents = 1200 * np.abs(np.log2(f0[mid] / f))
        assert cents.max() < 10.0  # within one 10-cent pitch bin
    f0, vf, vp = L.pyin(np.zeros(8192, np.float32), fmin=L.note_to_hz("E2"), fmax=L.note_to_hz("C6"), sr=sr)
    assert not vf.any() and np.isnan(f0).all() and (vp == 0).all()


def test_rms_and_db_shapes():
    y = corpus.test_track(22050, 0)
    r = L.rms(y)
    assert r.shape == (1, 1 + len(y) // 512) and r.dtype == np.float32
    S = L.load_audio_features(y, 22050)
    assert S.shape == (128, r.shape[1]) and S.dtype == np.float32
    assert S.max() == 0.0 and S.min() >= -80.0


def test_onset_detect_finds_the_plucks():
    y = corpus.test_track(22050, 0)
    env = L.onset_strength(y=y, sr=22050)
    assert env.shape == (1 + len(y) // 512,) and (env[:3] == 0).all()
    on = L.onset_detect(onset_envelope=env, sr=22050)
    starts = np.array([0.2, 1.4, 2.6 + 0.025 + 1000 / 22050]) * 22050 / 512  # E2, A2 .. pluck times
    assert all(np.min(np.abs(on - s)) <= 2 for s in starts[:2])
    assert L.onset_detect(onset_envelope=np.zeros(50, np.float32), sr=22050).size == 0


def test_polyphase_resampler_restatement_matches_scipy():
    """librosa.resample(res_type='polyphase') is scipy.signal.resample_poly (scipy is the pinned third-party
    dependency here; libsoxr, librosa's default, is absent): the float64 restatement agrees with it to float32
    rounding, its filter equals scipy.signal.firwin, and the product's design table carries scipy's centring."""
    import scipy.signal

    from spectrogram_midi_b200 import tables

    rng = np.random.default_rng(0)
    for orig, target, n in [(44100, 22050, 5000), (48000, 22050, 7000), (22050, 44100, 3001), (16000, 22050, 4000),
                            (44100, 22050, 1), (48000, 44100, 2500), (22050, 22050, 100)]:
        y = rng.uniform(-1, 1, n).astype(np.float32)
        g = int(np.gcd(orig, target))
        up, down = target // g, orig // g
        ref = scipy.signal.resample_poly(y, up, down)
        got = L.resample_polyphase(y, orig, target)
        assert got.dtype == np.float32 and got.shape == ref.shape == (-(-n * up // down),)
        np.testing.assert_allclose(got, ref, rtol=0, atol=1e-6)
        if up != down:
            h, half_len = L.kaiser_lowpass(up, down)
            np.testing.assert_allclose(h, scipy.signal.firwin(2 * half_len + 1, 1.0 / max(up, down), window=("kaiser", 5.0)), rtol=1e-12, atol=1e-18)
            taps, n_pre_pad, n_pre_remove = tables.resample_poly_design(up, down)
            assert taps.dtype == np.float32 and len(taps) == 2 * half_len + 1
            assert (n_pre_remove * down - n_pre_pad) == half_len     # output 0 sits on the filter centre
    stereo = rng.uniform(-1, 1, (2, 50)).astype(np.float32)
    np.testing.assert_array_equal(L.to_mono(stereo), (stereo[0] + stereo[1]) / 2)


def test_resampler_kernel_arithmetic_model_equals_scipy_bitwise():
    """The arithmetic K9 performs -- for output m the taps h[phase + j * up] against x[q - j], q = c // up,
    phase = c % up, c = (m + n_pre_remove) * down - n_pre_pad, summed over ascending input index with a separately
    rounded float32 product and sum -- restated in numpy float32, is bit-identical to scipy.signal.resample_poly.
    (The GPU test checks the kernel against scipy directly; this pins the index algebra and the design table on CPU.)"""
    import scipy.signal

    from spectrogram_midi_b200 import tables

    f32 = np.float32
    rng = np.random.default_rng(5)
    for orig, target, n in [(44100, 22050, 400), (48000, 22050, 500), (22050, 44100, 150), (16000, 22050, 260), (88200, 22050, 700)]:
        y = rng.uniform(-1, 1, n).astype(f32)
        g = int(np.gcd(orig, target))
        up, down = target // g, orig // g
        h, n_pre_pad, n_pre_remove = tables.resample_poly_design(up, down)
        taps_per_phase = -(-len(h) // up)
        n_out = -(-n * up // down)
        out = np.zeros(n_out, f32)
        for m in range(n_out):
            c = (m + n_pre_remove) * down - n_pre_pad
            q, phase = divmod(c, up)
            acc = f32(0)
            for j in range(taps_per_phase - 1, -1, -1):
                i, k = phase + j * up, q - j
                if i < len(h) and 0 <= k < n:
                    acc = f32(acc + f32(y[k] * h[i]))
            out[m] = acc
        np.testing.assert_array_equal(out, scipy.signal.resample_poly(y, up, down), err_msg=f"{orig}->{target}")



def test_uncalled_reference_filters_match_golden(golden):
    """atr_filter / ichimoku_baseline / stochastic_oscillator (financial_filters.py:144-249; defined by the reference,
    called by nothing in it) are host numpy in the drop-in class: bit-equal to the real file's outputs."""
    import warnings

    import spectrogram_midi_b200  # noqa: F401
    from spectrogram_midi_b200.financial_filters import FinancialNoiseFilters as F

    for name in _cases(golden, "filt"):
        f0 = golden[f"filt/{name}/f0"]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            at, am = F.atr_filter(f0.copy())
            ich = F.ichimoku_baseline(f0.copy())
            sto = F.stochastic_oscillator(f0.copy())
        np.testing.assert_array_equal(at, golden[f"filt/{name}/atr"], err_msg=name)
        np.testing.assert_array_equal(am, golden[f"filt/{name}/atr_mask"], err_msg=name)
        np.testing.assert_array_equal(ich, golden[f"filt/{name}/ichimoku"], err_msg=name)
        np.testing.assert_array_equal(sto, golden[f"filt/{name}/stochastic"], err_msg=name)
