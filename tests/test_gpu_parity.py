"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path, called through the C ABI via
the Python host layer, against the CPU oracle and the committed golden vectors.

Tolerances (BASELINE.json north_star): spectrogram rtol 1e-4 (+ atol 1e-5 * frame max),
RMS rtol 1e-5, f0 within 1 cent on voiced frames; integer outputs (voiced flags, Viterbi states on
identical observations, rake mask, onset frames, MIDI note events) bit-exact.
"""
import os
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import librosa_ref as L  # noqa: E402
from oracle import reference_files as R  # noqa: E402

import spectrogram_midi_b200 as P  # noqa: E402
from spectrogram_midi_b200 import corpus, tables  # noqa: E402

E2, C6, E6 = L.note_to_hz("E2"), L.note_to_hz("C6"), L.note_to_hz("E6")


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from spectrogram_midi_b200 import _native

    _native.load()  # the native library must be the thing that runs: fail loudly otherwise
    return torch.device("cuda", 0)


def _dev(y, dev):
    y = np.atleast_2d(np.asarray(y, dtype=np.float32))
    return torch.from_numpy(np.ascontiguousarray(y)).to(dev)


SIGNALS = {
    "track22050": lambda: (corpus.test_track(22050, 0), 22050),
    "track44100": lambda: (corpus.test_track(44100, 0), 44100),
    "bench22050": lambda: (corpus.benchmark_signal(22050, 0), 22050),
    "clip3": lambda: (corpus.random_clip(3, 6.0, 22050), 22050),
}


def _assert_mag_close(got, ref):
    tol = 1e-4 * np.abs(ref) + 1e-5 * ref.max(axis=0, keepdims=True) + 1e-30
    bad = np.abs(got - ref) > tol
    assert not bad.any(), f"{bad.sum()} of {bad.size} bins outside rtol 1e-4 (+1e-5*frame max); worst {np.abs(got - ref).max()}"


# ---------------------------------------------------------------------------------- K1 STFT / mel / RMS
@pytest.mark.parametrize("name", list(SIGNALS))
def test_stft_magnitude_matches_oracle(dev, name):
    y, sr = SIGNALS[name]()
    out = P.core.stft_features(_dev(y, dev), sr=sr, want_mag=True, want_mel=True, want_rms=True)
    ref = L.stft_magnitude(y)
    assert out["mag"].shape == (1,) + ref.shape
    _assert_mag_close(out["mag"][0].cpu().numpy(), ref)
    np.testing.assert_allclose(out["rms"][0].cpu().numpy(), L.rms(y)[0], rtol=1e-5, atol=1e-9)
    mel_ref = L.melspectrogram(y, sr)
    mel = out["mel"][0].cpu().numpy()
    np.testing.assert_allclose(mel, mel_ref, rtol=2e-4, atol=2e-6 * mel_ref.max())
    assert abs(float(out["mel_max"][0]) - mel.max()) == 0.0


@pytest.mark.parametrize("n", [0, 1, 511, 512, 2047, 2048, 4097, 12345, 22050])
def test_stft_edge_lengths(dev, n):
    rng = np.random.default_rng(n)
    y = rng.normal(0, 0.3, n).astype(np.float32)
    if n == 0:
        out = P.core.stft_features(torch.zeros((1, 0), device=dev), want_mag=True, want_rms=True)
        assert out["mag"].shape == (1, 1025, 1) and float(out["mag"].abs().max()) == 0.0
        return
    out = P.core.stft_features(_dev(y, dev), want_mag=True, want_rms=True)
    ref = L.stft_magnitude(y)
    assert out["mag"].shape[1:] == ref.shape
    _assert_mag_close(out["mag"][0].cpu().numpy(), ref)
    np.testing.assert_allclose(out["rms"][0].cpu().numpy(), L.rms(y)[0], rtol=1e-5, atol=1e-9)


def test_stft_ragged_batch_unaligned_rows_and_silence(dev):
    # rows of 4099 samples: every second clip starts on a non-16-byte boundary (scalar load path)
    rng = np.random.default_rng(7)
    y = rng.normal(0, 0.2, (5, 4099)).astype(np.float32)
    y[2] = 0.0  # digital silence must give exact zeros
    out = P.core.stft_features(_dev(y, dev), want_mag=True, want_rms=True)
    for c in range(5):
        _assert_mag_close(out["mag"][c].cpu().numpy(), L.stft_magnitude(y[c]))
    assert float(out["mag"][2].abs().max()) == 0.0 and float(out["rms"][2].abs().max()) == 0.0
    # a strided view (clip_stride > n_samples) is honoured
    big = _dev(rng.normal(0, 0.2, (3, 9000)).astype(np.float32), dev)
    view = big[:, :8192]
    o2 = P.core.stft_features(view, want_mag=True)
    for c in range(3):
        _assert_mag_close(o2["mag"][c].cpu().numpy(), L.stft_magnitude(view[c].cpu().numpy()))


def test_stft_output_layouts_agree(dev):
    """|X| through every store path of K1 -- TMA into the default pitched buffer, TMA into a dense caller buffer
    (T % 4 == 0), the 16-byte LSU path (pitch % 4 == 0 but a misaligned base) and the scalar LSU path (odd pitch) --
    is bit-identical, partial tiles at the clip end included; bytes outside the view are never touched."""
    rng = np.random.default_rng(3)
    for n_samples in (512 * 43, 512 * 44 + 100, 512 * 41 + 7):     # T = 44 (dense % 4 == 0), 45, 42
        y = _dev(rng.normal(0, 0.2, (3, n_samples)).astype(np.float32), dev)
        T = 1 + n_samples // 512
        ref = P.core.stft_features(y, want_mag=True)["mag"]            # pitched view, TMA
        assert ref.stride(1) % 8 == 0 and ref.shape == (3, 1025, T)
        dense = torch.full((3, 1025, T), -1.0, device=dev)
        P.core.stft_features(y, want_mag=True, mag_out=dense)
        assert torch.equal(dense, ref)
        big = torch.full((3, 1025, T + 6), -1.0, device=dev)
        for off in (2, 1):                                           # base misaligned by 8 / 4 bytes
            big.fill_(-1.0)
            view = big[:, :, off:off + T]
            P.core.stft_features(y, want_mag=True, mag_out=view)
            assert torch.equal(view, ref)
            assert bool((big[:, :, :off] == -1.0).all()) and bool((big[:, :, off + T:] == -1.0).all())
        odd = torch.full((3, 1025, T + 1), -1.0, device=dev)          # row pitch T + 1
        P.core.stft_features(y, want_mag=True, mag_out=odd[:, :, :T])
        assert torch.equal(odd[:, :, :T], ref) and bool((odd[:, :, T] == -1.0).all())


def test_stft_hop_and_uncentred(dev):
    y, sr = SIGNALS["clip3"]()
    for hop in (128, 256):
        out = P.core.stft_features(_dev(y, dev), hop_length=hop, want_mag=True)
        _assert_mag_close(out["mag"][0].cpu().numpy(), L.stft_magnitude(y, hop_length=hop))
    out = P.core.stft_features(_dev(y, dev), center=False, want_mag=True)
    _assert_mag_close(out["mag"][0].cpu().numpy(), L.stft_magnitude(y, center=False))


def test_stft_size_independent_properties_at_batch_scale(dev):
    """Homogeneity (x2 is exact in fp32), hop-shift equivariance and batch independence on a
    64 x 30 s batch of BASELINE cfg2 clips (too large to compare against the oracle bin by bin)."""
    clips = corpus.clip_batch(4, 30.0, 22050, first_seed=100)
    y = _dev(np.tile(clips, (16, 1)), dev)
    y = y * torch.linspace(0.25, 1.0, 64, device=dev)[:, None]
    a = P.core.stft_features(y, sr=22050, want_mag=True, want_mel=True, want_rms=True)
    b = P.core.stft_features(y * 2.0, sr=22050, want_mag=True, want_rms=True)
    assert torch.equal(a["rms"] * 2.0, b["rms"]), f"max |2a-b| rms = {float((a['rms'] * 2.0 - b['rms']).abs().max())}"
    assert torch.equal(a["mag"] * 2.0, b["mag"]), f"max |2a-b| = {float((a['mag'] * 2.0 - b['mag']).abs().max())}"
    shifted = P.core.stft_features(y[:, 512:].contiguous(), want_mag=True)["mag"]
    T = shifted.shape[2]
    # interior frames move by one hop; each frame now shares its transform with a different partner, so
    # equality holds to rounding (relative to the frame's own norm), not bit for bit
    ref_sh = a["mag"][:, :, 3:T - 1]
    tol = 2e-6 * ref_sh.amax(dim=1, keepdim=True) + 1e-5 * ref_sh
    assert bool(((shifted[:, :, 2:T - 2] - ref_sh).abs() <= tol).all())
    alone = P.core.stft_features(y[37:38].contiguous(), sr=22050, want_mag=True, want_mel=True)
    assert torch.equal(alone["mag"][0], a["mag"][37]) and torch.equal(alone["mel"][0], a["mel"][37])
    _assert_mag_close(a["mag"][5].cpu().numpy(), L.stft_magnitude(y[5].cpu().numpy()))
    # Parseval on the windowed frames: sum |X|^2 (two-sided) == N * sum (w x)^2
    w = torch.from_numpy(tables.hann_window().astype(np.float32)).to(dev)
    fr = torch.nn.functional.pad(y[:2], (1024, 1024)).unfold(1, 2048, 512) * w
    lhs = (a["mag"][:2] ** 2).sum(1) * 2 - a["mag"][:2, 0] ** 2 - a["mag"][:2, 1024] ** 2
    rhs = 2048 * (fr ** 2).sum(2)
    assert torch.allclose(lhs, rhs, rtol=2e-4, atol=1e-3)


# ---------------------------------------------------------------------------------- K4 dB / rake / onsets
def test_rake_mask_kernel_matches_reference_golden(dev, golden):
    names = sorted({k.split("/")[1] for k in golden if k.startswith("rake/")})
    fired = 0
    for n in names:
        hop, sr, ratio = golden[f"rake/{n}/args"]
        got = P.vision.detect_rake_patterns(golden[f"rake/{n}/S_dB"], int(hop), int(sr), float(ratio))
        assert got.dtype == bool
        np.testing.assert_array_equal(got, golden[f"rake/{n}/mask"], err_msg=n)
        fired += int(got.sum())
    assert fired > 0


@pytest.mark.parametrize("name", list(SIGNALS))
def test_db_rake_rms_onsets_match_oracle(dev, name):
    y, sr = SIGNALS[name]()
    spec = P.batch.spectral_features(_dev(y, dev), sr=sr, with_sdb=True, with_onsets=True)
    S_ref = L.load_audio_features(y, sr)
    S = spec["S_dB"][0].cpu().numpy()
    assert S.dtype == np.float32 and S.shape == S_ref.shape
    assert np.abs(S - S_ref).max() < 5e-3 and S.max() == 0.0 and S.min() >= -80.0
    for ratio in (0.6, 0.5):
        m = P.batch.spectral_features(_dev(y, dev), sr=sr, rake_sensitivity=ratio, with_onsets=False)["rake_mask"]
        np.testing.assert_array_equal(m[0].cpu().numpy().astype(bool), R.detect_rake_patterns(S_ref, 512, sr, ratio))
    env_ref = L.onset_strength(y=y, sr=sr)
    env = spec["onset_env"][0].cpu().numpy()
    np.testing.assert_allclose(env, env_ref, atol=2e-3)
    frames = np.flatnonzero(spec["onset_peaks"][0].cpu().numpy())
    np.testing.assert_array_equal(frames, L.onset_detect(onset_envelope=env_ref, sr=sr))
    assert int(spec["n_onsets"][0]) == len(frames)
    # the peak picker on the oracle's own envelope (isolates K4b from fp32 noise in the envelope)
    np.testing.assert_array_equal(P.librosa_compat.onset_detect(onset_envelope=env_ref, sr=sr),
                                  L.onset_detect(onset_envelope=env_ref, sr=sr))


def test_compat_signatures_and_dtypes(dev):
    lib = P.librosa_compat
    y, sr = SIGNALS["track22050"]()
    S = lib.feature.melspectrogram(y=y, sr=sr, n_fft=2048, hop_length=512)
    S_dB = lib.power_to_db(S, ref=np.max)
    assert S.shape == S_dB.shape == (128, 1 + len(y) // 512) and S_dB.dtype == np.float32
    np.testing.assert_allclose(lib.power_to_db(S, ref=1.0), L.power_to_db(S), atol=5e-3)
    r = lib.feature.rms(y=y, hop_length=512)
    assert r.shape == (1, S.shape[1]) and r.dtype == np.float32
    f0, vf, vp = lib.pyin(y, fmin=lib.note_to_hz("E2"), fmax=lib.note_to_hz("C6"), sr=sr, hop_length=512)
    assert f0.dtype == np.float64 and vf.dtype == bool and vp.dtype == np.float64 and len(f0) == S.shape[1]
    assert np.isnan(f0[~vf]).all() and not np.isnan(f0[vf]).any()
    f0b, vfb, _ = lib.pyin(y, fmin=E2, fmax=C6, sr=sr, fill_na=None)
    assert not np.isnan(f0b).any() and np.array_equal(vf, vfb)
    f0r, vfr, _ = lib.pyin(y, fmin=E2, fmax=C6, sr=sr, pad_mode="reflect")
    ref = L.pyin(np.pad(y, 1024, mode="reflect"), fmin=E2, fmax=C6, sr=sr, center=False)
    assert np.array_equal(vfr, ref[1])
    with pytest.raises(lib.ParameterError):
        lib.pyin(y, fmin=None, fmax=C6)
    with pytest.raises(TypeError):
        lib.util.softmask(np.ones(3), np.ones(3), margin=0.5)  # midi_logic.py:43 relies on this TypeError
    assert P.worker._pyin_worker((y[:8192], sr, 512))[0].shape == (17,)
    eng = P.AegisEngine(sample_rate=sr)
    assert eng.audio_to_midi(np.zeros(0, np.float32), None) is None
    res = eng.audio_to_midi(y, "ignored.mid")
    assert sorted(res) == ["f0", "rake_mask", "rms", "voiced_flag", "voiced_probs", "y"]
    assert res["y"] is not None and not np.isnan(res["f0"]).any() and res["rake_mask"].dtype == bool
    y2, S2 = eng.load_audio(y)
    np.testing.assert_array_equal(eng.detect_rake_patterns(S2), res["rake_mask"])


# ---------------------------------------------------------------------------------- K2 / K3 pYIN
def _oracle_pyin(y, sr, fmax=C6):
    return L.pyin(y, fmin=E2, fmax=fmax, sr=sr, hop_length=512, return_intermediates=True)


def _sparse_from_dense(obs, n_bins, max_cand, dev):
    T = obs.shape[1]
    cb = np.zeros((T, max_cand), np.int16)
    cp = np.zeros((T, max_cand), np.float64)
    cc = np.zeros(T, np.int32)
    for t in range(T):
        nz = np.flatnonzero(obs[:n_bins, t])
        cc[t] = len(nz)
        cb[t, : len(nz)] = nz
        cp[t, : len(nz)] = obs[nz, t]
    vp = np.clip(obs[:n_bins].sum(axis=0), 0, 1)
    return dict(cand_bin=torch.from_numpy(cb).to(dev), cand_prob=torch.from_numpy(cp).to(dev),
                cand_count=torch.from_numpy(cc).to(dev), voiced_prob=torch.from_numpy(vp).to(dev)[None],
                n_frames=T, max_cand=max_cand)


@pytest.mark.parametrize("name,fmax", [("track22050", C6), ("track44100", C6), ("clip3", C6), ("clip3", E6), ("bench22050", C6)])
def test_viterbi_is_bit_exact_on_oracle_observations(dev, name, fmax):
    y, sr = SIGNALS[name]()
    f0, vf, vp, it = _oracle_pyin(y, sr, fmax)
    cfg = tables.pyin_config(float(sr), 512, E2, fmax)
    assert cfg.n_pitch_bins == it["n_pitch_bins"]
    obs = _sparse_from_dense(it["observation_probs"], cfg.n_pitch_bins, cfg.max_troughs, dev)
    dec = P.core.viterbi_decode(obs, cfg, 1)
    states = dec["states"][0].cpu().numpy().astype(np.uint16)
    np.testing.assert_array_equal(states, it["states"])
    np.testing.assert_array_equal(dec["voiced_flag"][0].cpu().numpy().astype(bool), vf)
    np.testing.assert_array_equal(dec["f0"][0].cpu().numpy(), f0)  # NaN == NaN under assert_array_equal


def test_viterbi_adversarial_observations(dev):
    """Random sparse observations incl. voiced_prob == 1 (unvoiced rows log(tiny)), far jumps that force
    out-of-band transitions, empty frames and edge bins: states must still equal the dense float64 decoder."""
    sr = 22050
    cfg = tables.pyin_config(float(sr), 512, E2, C6)
    n = cfg.n_pitch_bins
    trans, _ = L.pyin_transition(n, 10, sr, 512)
    p_init = np.zeros(2 * n)
    p_init[n:] = 1 / n
    rng = np.random.default_rng(11)
    for trial in range(3):
        T = 120
        obs = np.zeros((2 * n, T))
        for t in range(T):
            k = int(rng.integers(0, 6))
            bins = rng.choice(n, size=k, replace=False) if trial else rng.choice([0, 1, 2, n - 1, n - 2, 200, 260], size=min(k, 7), replace=False)
            pr = rng.random(len(bins))
            if len(bins) and rng.random() < 0.4:
                pr = pr / pr.sum()  # voiced_prob == 1 up to rounding, clipped to exactly 1 below
            else:
                pr = pr * rng.random() / max(1, len(bins))
            obs[bins, t] = pr
        vp = np.clip(obs[:n].sum(axis=0, keepdims=True), 0, 1)
        obs[n:, :] = (1 - vp) / n
        ref = L.viterbi(obs, trans, p_init)
        dec = P.core.viterbi_decode(_sparse_from_dense(obs, n, cfg.max_troughs, dev), cfg, 1)
        np.testing.assert_array_equal(dec["states"][0].cpu().numpy().astype(np.uint16), ref, err_msg=f"trial {trial}")


def test_viterbi_collapse_runs_equal_dense_decoder(dev):
    """Long stretches of voiced_prob == 1 frames (K3 decodes them in its one-warp collapse run) with everything that can end,
    interrupt or stress a run: slowly drifting note candidates with octave partners, far jumps (out-of-band transitions),
    equal candidate masses (ties), a candidate whose mass is exactly zero, frames with 33..40 candidates in the middle of a
    run and as its first frame, single voiced_prob < 1 frames, silence, a run that reaches the last frame, runs at the edge
    bins.  The decoded states must equal the dense float64 decoder's (oracle) exactly."""
    sr = 22050
    cfg = tables.pyin_config(float(sr), 512, E2, C6)
    n = cfg.n_pitch_bins
    assert cfg.max_troughs >= 40
    trans, _ = L.pyin_transition(n, 10, sr, 512)
    p_init = np.zeros(2 * n)
    p_init[n:] = 1 / n
    rng = np.random.default_rng(23)
    for trial in range(4):
        T = 260
        obs = np.zeros((2 * n, T))
        centre = [int(rng.integers(60, n - 60)), 3, n - 4, int(rng.integers(0, n))][trial]
        t = 0
        while t < T:
            kind = rng.random()
            if kind < 0.70:      # a run of exact voiced_prob == 1 frames
                ln = int(rng.integers(2, 60))
                for u in range(t, min(T, t + ln)):
                    centre = int(np.clip(centre + rng.integers(-2, 3), 0, n - 1))
                    if rng.random() < 0.04:
                        centre = int(rng.integers(0, n))          # far jump inside the run
                    k = int(rng.integers(1, 5))
                    if rng.random() < 0.03:
                        k = int(rng.integers(33, 41))              # more candidates than a run can hold
                    bins = {centre}
                    while len(bins) < k:
                        bins.add(int(np.clip(centre + rng.choice([-120, -70, -12, -1, 1, 12, 70, 120]) + rng.integers(-3, 4), 0, n - 1)) if k < 8 else int(rng.integers(0, n)))
                    bins = sorted(bins)
                    if rng.random() < 0.3:
                        pr = np.full(len(bins), 1.0 / len(bins))   # ties between candidates
                        if len(bins) in (2, 4, 8):                  # masses that sum to exactly 1
                            pass
                        else:
                            pr = rng.dirichlet(np.ones(len(bins)))
                    else:
                        pr = rng.dirichlet(np.ones(len(bins)))
                    obs[bins, u] = pr
                    if len(bins) > 1 and rng.random() < 0.05:
                        obs[bins[0], u] = 0.0                      # drops out of the sparse list: the frame is no longer exactly 1
                t += ln
            elif kind < 0.85:    # one or two frames with voiced_prob < 1
                for u in range(t, min(T, t + int(rng.integers(1, 3)))):
                    k = int(rng.integers(0, 4))
                    bins = rng.choice(n, size=k, replace=False)
                    obs[bins, u] = rng.random(k) * 0.3 / max(1, k)
                t += 2
            else:                # silence
                t += int(rng.integers(1, 6))
        if trial == 1:
            obs[:n, T - 5:] = 0.0
            obs[3, T - 5:] = 1.0                                    # the run reaches the last frame
        vp = np.clip(obs[:n].sum(axis=0, keepdims=True), 0, 1)
        snap = np.abs(vp - 1.0) < 1e-12
        vp[snap] = 1.0                                             # what K2 hands over when the masses sum to one
        obs[n:, :] = (1 - vp) / n
        assert (vp == 1.0).mean() > 0.4
        ref = L.viterbi(obs, trans, p_init)
        sp = _sparse_from_dense(obs, n, cfg.max_troughs, dev)
        sp["voiced_prob"] = torch.from_numpy(vp).to(dev)
        dec = P.core.viterbi_decode(sp, cfg, 1)
        np.testing.assert_array_equal(dec["states"][0].cpu().numpy().astype(np.uint16), ref, err_msg=f"trial {trial}")


def _dense_obs(obs, n):
    T = obs["n_frames"]
    cb = obs["cand_bin"].cpu().numpy().astype(np.int64)
    cp = obs["cand_prob"].cpu().numpy()
    cc = obs["cand_count"].cpu().numpy()
    dense = np.zeros((n, T))
    for t in range(T):
        assert (np.diff(cb[t, : cc[t]]) > 0).all()  # ascending, unique
        dense[cb[t, : cc[t]], t] = cp[t, : cc[t]]
    return dense


@pytest.mark.parametrize("name,fmax", [("track22050", C6), ("track44100", C6), ("clip3", C6), ("clip3", E6), ("bench22050", C6)])
def test_yin_candidates_match_oracle(dev, name, fmax):
    """Sparse observations vs the oracle.  librosa evaluates the difference function in float32, so its
    OWN output moves when a trough sits on a threshold edge or a .5 bin boundary; the float64 evaluation
    of the same algorithm (`frames_dtype=float64`) measures that noise.  Required: on the
    generate_test_signal tracks every frame agrees; elsewhere the kernel is at least as close to the
    float64 result as the float32 oracle is (plus 1% slack), and frames that agree, agree tightly."""
    y, sr = SIGNALS[name]()
    f0, vf, vp, it = _oracle_pyin(y, sr, fmax)
    it64 = L.pyin(y, fmin=E2, fmax=fmax, sr=sr, hop_length=512, return_intermediates=True, frames_dtype=np.float64)[3]
    cfg = tables.pyin_config(float(sr), 512, E2, fmax)
    n = cfg.n_pitch_bins
    obs = P.core.yin_candidates(_dev(y, dev), cfg)
    assert int(obs["overflow"][0]) == 0
    dense = _dense_obs(obs, n)
    ref32, ref64 = it["observation_probs"][:n], it64["observation_probs"][:n]
    T = dense.shape[1]

    def differing(a, b):
        return (np.abs(a - b).max(axis=0) > 2e-3)

    gpu_vs_64, ref_vs_64, gpu_vs_32 = differing(dense, ref64), differing(ref32, ref64), differing(dense, ref32)
    print(f"{name}: frames={T} gpu!=f64 {gpu_vs_64.sum()}  oracle32!=f64 {ref_vs_64.sum()}  gpu!=oracle32 {gpu_vs_32.sum()}")
    if name.startswith("track"):
        assert gpu_vs_32.sum() == 0 and gpu_vs_64.sum() == 0
    assert gpu_vs_64.sum() <= ref_vs_64.sum() + max(1, int(0.01 * T))
    ok = ~gpu_vs_64
    assert (np.abs(dense[:, ok] - ref64[:, ok]).max(axis=0) < 1e-5).mean() >= 0.95
    np.testing.assert_allclose(obs["voiced_prob"][0].cpu().numpy()[ok], np.clip(ref64[:, ok].sum(axis=0), 0, 1), atol=2e-3)


@pytest.mark.parametrize("name", ["track22050", "track44100", "bench22050"])
def test_pyin_end_to_end_voicing_exact_f0_within_one_cent(dev, name):
    """BASELINE bar on the generate_test_signal / benchmark_aegis signals: voiced flags and MIDI note
    numbers bit-exact, f0 within one cent on voiced frames."""
    y, sr = SIGNALS[name]()
    f0_ref, vf_ref, vp_ref = L.pyin(y, fmin=E2, fmax=C6, sr=sr, hop_length=512)
    f0, vf, vp = P.librosa_compat.pyin(y, fmin=E2, fmax=C6, sr=sr, hop_length=512)
    np.testing.assert_array_equal(vf, vf_ref)          # voiced flags: bit-exact
    both = vf & vf_ref
    cents = 1200 * np.abs(np.log2(f0[both] / f0_ref[both]))
    assert cents.max() <= 1.0, f"{(cents > 1).sum()} voiced frames off by more than one cent (max {cents.max():.2f})"
    np.testing.assert_allclose(vp, vp_ref, atol=2e-3)
    midi = np.round(L.hz_to_midi(f0[both])).astype(int)
    np.testing.assert_array_equal(midi, np.round(L.hz_to_midi(f0_ref[both])).astype(int))  # MIDI note numbers


@pytest.mark.parametrize("name", ["track22050", "track44100", "bench22050"])
def test_cmnd_curves_match_oracle(dev, name):
    """The CMND curve itself (SURVEY.md build plan step 5) on the generate_test_signal / benchmark signals."""
    y, sr = SIGNALS[name]()
    it = _oracle_pyin(y, sr)[3]
    cfg = tables.pyin_config(float(sr), 512, E2, C6)
    got = P.core.yin_candidates(_dev(y, dev), cfg, want_cmnd=True)["cmnd"][0].cpu().numpy().T
    ref = it["yin_frames"]
    assert got.shape == ref.shape
    loud = (L.frame_signal(y) ** 2).sum(axis=0) > 1e-3      # near-silent frames are 0/0-conditioned
    err = np.abs(got - ref)[:, loud]
    assert np.percentile(err, 99.9) < 2e-4 and np.median(err) < 2e-6


@pytest.mark.parametrize("seed,dur", [(3, 6.0), (7, 12.0), (20, 5.0), (21, 8.0)])
def test_pyin_random_clips_no_worse_than_reference_rounding(dev, seed, dur):
    """Random KS clips have long, very smooth decay tails where d[tau] ~ 1e-7 * energy at small lags:
    float32 evaluation of e0 + e[tau] - 2 acf[tau] cancels completely there (librosa's own d[1] comes out
    negative), its CMND is off by ~1e-2 and the decode of those frames is rounding noise.  Parity is
    therefore stated against the float64 evaluation of the same algorithm:
      * the kernel's CMND error vs float64 is no larger than the float32 oracle's (in aggregate);
      * on frames where the float32 oracle itself is well conditioned (its CMND within 1e-3 of float64,
        +-8 frames of HMM context) the decoded states equal the float64 decode (>= 99%), voiced flags too."""
    y = corpus.random_clip(seed, dur, 22050)
    cfg = tables.pyin_config(22050.0, 512, E2, C6)
    r32 = L.pyin(y, fmin=E2, fmax=C6, sr=22050, hop_length=512, return_intermediates=True)
    r64 = L.pyin(y, fmin=E2, fmax=C6, sr=22050, hop_length=512, return_intermediates=True, frames_dtype=np.float64)
    yd = _dev(y, dev)
    cm = P.core.yin_candidates(yd, cfg, want_cmnd=True)["cmnd"][0].cpu().numpy().T
    c32, c64 = r32[3]["yin_frames"], r64[3]["yin_frames"]
    en = (L.frame_signal(y, dtype=np.float64) ** 2).sum(axis=0)
    loud = en > 1e-4 * en.max()
    e_gpu = np.abs(cm - c64).max(axis=0)[loud]
    e_ref = np.abs(c32 - c64).max(axis=0)[loud]
    q = lambda e: "median %.2e p90 %.2e p99 %.2e" % tuple(np.percentile(e, [50, 90, 99]))
    print(f"seed {seed}: CMND err vs f64 over {loud.sum()} frames:  gpu {q(e_gpu)} | oracle32 {q(e_ref)}")
    assert np.median(e_gpu) <= 2.0 * np.median(e_ref) + 1e-6
    assert np.percentile(e_gpu, 90) <= 3.0 * np.percentile(e_ref, 90) + 1e-5

    out = P.core.pyin_batch(yd, sr=22050, fmin=E2, fmax=C6)
    st = out["states"][0].cpu().numpy().astype(np.uint16)
    vf = out["voiced_flag"][0].cpu().numpy().astype(bool)
    T = len(st)
    ill = (np.abs(c32 - c64).max(axis=0) > 1e-3) | ~loud | (r32[3]["states"] != r64[3]["states"])
    ctx = np.convolve(ill.astype(float), np.ones(17), mode="same") > 0
    ok = ~ctx
    d_all, d_ok = (st != r64[3]["states"]).sum(), (st[ok] != r64[3]["states"][ok]).sum()
    print(f"seed {seed}: frames={T} well-conditioned={ok.sum()} state diffs gpu!=f64: all {d_all}, well-conditioned {d_ok}; "
          f"oracle32!=f64 all {(r32[3]['states'] != r64[3]['states']).sum()}")
    assert ok.sum() >= 0.3 * T
    assert d_ok <= max(1, int(0.01 * ok.sum()))
    assert (vf[ok] != r64[1][ok]).sum() <= max(1, int(0.01 * ok.sum()))
    same = vf & r64[1] & ok
    cents = 1200 * np.abs(np.log2(out["f0"][0].cpu().numpy()[same] / r64[0][same]))
    assert (cents <= 1.0).mean() >= 0.99


def test_pyin_tables_are_not_aliased_across_hmm_parameters(dev):
    """the device-table cache was keyed on (sr, hop, fmin, fmax) only.  Two calls in one process that differ in
    switch_prob / resolution / beta_parameters must each see their own tables (a stale row_variant with another
    n_pitch_bins was an out-of-bounds read), and each must still equal the oracle."""
    y, sr = SIGNALS["track22050"]()
    y = y[: sr * 4]
    yd = _dev(y, dev)
    for kw in ({}, {"switch_prob": 0.05}, {"resolution": 0.2}, {"beta_parameters": (3, 12)}, {}):
        got = P.core.pyin_batch(yd, sr=float(sr), fmin=E2, fmax=C6, **kw)
        f0, vf, vp = L.pyin(y, fmin=E2, fmax=C6, sr=sr, hop_length=512, **kw)
        np.testing.assert_array_equal(got["voiced_flag"][0].cpu().numpy().astype(bool), vf, err_msg=str(kw))
        g = got["f0"][0].cpu().numpy()
        assert (np.abs(1200 * np.log2(g[vf] / f0[vf])) <= 1.0).all(), kw


def test_turbo_worker_and_parallel_pitch_tracking(dev):
    """Row a-6: `_pyin_worker((chunk, sr, hop))` (aegis_engine_core/worker.py:3-15) returns librosa.pyin of the chunk --
    checked against the oracle on a chunk of the test track -- and `AegisEngine._parallel_pitch_tracking` returns the
    serial full-clip decode (what Turbo Mode approximates with independent chunks, aegis_engine.py:183-216)."""
    y, sr = SIGNALS["track22050"]()
    chunk = y[4000 : 4000 + 3 * sr]
    f0, vf, vp = P.worker._pyin_worker((chunk, sr, 512))
    r0, rv, rp = L.pyin(chunk, fmin=E2, fmax=C6, sr=sr, hop_length=512)
    assert f0.dtype == np.float64 and vf.dtype == bool and vp.dtype == np.float64 and f0.shape == r0.shape
    np.testing.assert_array_equal(vf, rv)
    np.testing.assert_allclose(vp, rp, atol=2e-3)
    assert (np.abs(1200 * np.log2(f0[vf] / r0[vf])) <= 1.0).all() and np.isnan(f0[~vf]).all()
    eng = P.AegisEngine(sample_rate=sr)
    t0, tv, tp = eng._parallel_pitch_tracking(y)
    s0, sv, sp = P.librosa_compat.pyin(y, fmin=E2, fmax=C6, sr=sr, hop_length=512)
    np.testing.assert_array_equal(tv, sv)
    np.testing.assert_array_equal(t0, s0)
    np.testing.assert_array_equal(tp, sp)


def test_yin_fused_and_split_kernels_agree_bit_for_bit(dev):
    """K2 at hop 512 has two forms: block sums + per-frame stage as two kernels (with the workspace; the product path) or
    one fused kernel (without).  A block sum depends only on its own samples and both forms sum it with the same lane map
    in the same order, so both forms -- and any batch / window a frame is analysed in -- give identical bits.  Other hops
    take the FFT kernel (not bit-equal)."""
    for sr in (22050, 44100):
        y = np.stack([corpus.random_clip(40 + i, 3.0, sr) for i in range(3)])
        y[2, 20000:] = 0.0
        cfg = tables.pyin_config(float(sr), 512, E2, C6)
        yd = _dev(y, dev)
        a = P.core.yin_candidates(yd, cfg, want_cmnd=True)
        b = P.core.yin_candidates(yd, cfg, want_cmnd=True, split=False)
        one = P.core.yin_candidates(yd[1:2], cfg, want_cmnd=True)
        for k in ("cand_count", "voiced_prob", "cmnd"):
            assert torch.equal(a[k], b[k]), (sr, k)
        T = a["n_frames"]
        cnt = a["cand_count"].view(3, T)
        m = torch.arange(a["max_cand"], device=dev)[None, None, :] < cnt[:, :, None]
        for k in ("cand_bin", "cand_prob"):
            assert torch.equal(a[k].view(3, T, -1)[m], b[k].view(3, T, -1)[m]), (sr, k)
        assert torch.equal(a["cmnd"][1], one["cmnd"][0]) and torch.equal(a["voiced_prob"][1], one["voiced_prob"][0])
        sil = slice(20000 // 512 + 5, None)   # digital silence is exact
        assert float(a["cmnd"][2, sil].abs().max()) == 0.0


def test_pyin_batch_equals_single_and_handles_silence(dev):
    clips = corpus.clip_batch(3, 4.0, 22050, first_seed=40)
    clips[1] = 0.0
    res = P.core.pyin_batch(_dev(clips, dev), sr=22050, fmin=E2, fmax=C6)
    assert not bool(res["voiced_flag"][1].any()) and float(res["voiced_prob"][1].abs().max()) == 0.0
    for c in (0, 2):
        one = P.core.pyin_batch(_dev(clips[c], dev), sr=22050, fmin=E2, fmax=C6)
        assert torch.equal(one["states"][0], res["states"][c])
        assert torch.equal(one["voiced_prob"][0], res["voiced_prob"][c])
    split = P.core.pyin_batch(_dev(clips, dev), sr=22050, fmin=E2, fmax=C6, clips_per_launch=2)
    assert torch.equal(split["states"], res["states"])


# ---------------------------------------------------------------------------------- K5 trend filters
EXACT = {"kalman": "kalman", "holt": "holt", "ema5": "ema", "macd": "macd_line", "macd_signal": "macd_sig",
         "macd_hist": "macd_hist"}
CLOSE = {"savgol": "savgol", "consensus": "consensus", "consensus_conf": "consensus_conf", "sma5": "sma",
         "boll_ma": "boll_ma", "boll_up": "boll_upper", "boll_lo": "boll_lower"}


def test_trend_filters_match_reference_golden(dev, golden):
    names = sorted({k.split("/")[1] for k in golden if k.startswith("filt/")})
    assert len(names) >= 6
    for n in names:
        f0 = golden[f"filt/{n}/f0"]
        out = P.core.trend_filters(torch.from_numpy(f0).to(dev), sma_window=5, ema_span=5, boll_window=10)
        got = {k: v[0].cpu().numpy() for k, v in out.items()}
        for gk, ok in EXACT.items():   # scalar recurrences: same IEEE operations in the same order
            np.testing.assert_array_equal(got[ok], golden[f"filt/{n}/{gk}"], err_msg=f"{n}/{gk}")
        for gk, ok in CLOSE.items():   # windowed sums: numpy/BLAS summation order is unspecified
            np.testing.assert_allclose(got[ok], golden[f"filt/{n}/{gk}"], rtol=1e-12, atol=1e-9, err_msg=f"{n}/{gk}")
        sma10 = P.core.trend_filters(torch.from_numpy(f0).to(dev), want=("sma",), sma_window=10)["sma"][0].cpu().numpy()
        np.testing.assert_allclose(sma10, golden[f"filt/{n}/sma10"], rtol=1e-12, atol=1e-9)


def test_filter_classes_mirror_reference_api(dev, golden):
    from spectrogram_midi_b200.financial_analysis import FinancialPitchAnalyzer
    from spectrogram_midi_b200.financial_filters import FinancialNoiseFilters, multi_filter_consensus

    for n in ("gappy", "allnan", "onevalid", "short"):
        f0 = golden[f"filt/{n}/f0"]
        np.testing.assert_array_equal(FinancialNoiseFilters.kalman_filter(f0), golden[f"filt/{n}/kalman"])
        np.testing.assert_array_equal(FinancialNoiseFilters.holt_winters(f0), golden[f"filt/{n}/holt"])
        np.testing.assert_allclose(FinancialNoiseFilters.savitzky_golay(f0), golden[f"filt/{n}/savgol"], rtol=1e-12, atol=1e-9)
        c, conf = multi_filter_consensus(f0)
        np.testing.assert_allclose(c, golden[f"filt/{n}/consensus"], rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(conf, golden[f"filt/{n}/consensus_conf"], rtol=1e-9, atol=1e-9)
        an = FinancialPitchAnalyzer(sr=22050, hop_length=512)
        np.testing.assert_allclose(an.bollinger_confidence(f0, 10), golden[f"filt/{n}/an_conf"], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(an.analyze_pitch_numeric(f0)["trend"], golden[f"filt/{n}/an_trend"], rtol=1e-12, atol=1e-9)
    with pytest.raises(IndexError):
        FinancialPitchAnalyzer().simple_moving_average(np.ones(4), window=5)  # the reference raises here too
    c, conf = multi_filter_consensus(golden["filt/gappy/f0"], filters=[])
    assert (conf == 1).all()
    # subsets of the filters vote through the kernel's consensus_mask (no eager-torch path): numpy on the golden rows
    import warnings
    for subset in (["kalman", "holt"], ["savgol"], ["holt", "savgol"]):
        for n in ("gappy", "onevalid", "long"):
            f0 = golden[f"filt/{n}/f0"]
            rows = np.array([golden[f"filt/{n}/{f}"] for f in ("savgol", "kalman", "holt") if f in subset])
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                want_c, want_conf = np.nanmedian(rows, axis=0), 1.0 / (1.0 + np.nanstd(rows, axis=0))
            c, conf = multi_filter_consensus(f0, filters=subset)
            np.testing.assert_allclose(c, want_c, rtol=1e-12, atol=1e-9, err_msg=f"{subset} {n}")
            np.testing.assert_allclose(conf, want_conf, rtol=1e-9, atol=1e-9, err_msg=f"{subset} {n}")


# ---------------------------------------------------------------------------------- K7 note events
TECH = {None: 0, "vibrato": 1, "bend": 2, "slide": 3, "hammer_on": 4, "pull_off": 5}


def _event_rows(ev):
    ints = np.array([[e["note"], e["start"], e["end"], e["velocity"], int(e["track"] == "main"), TECH[e.get("technique")]] for e in ev],
                    dtype=np.int64).reshape(-1, 6)
    flt = np.array([[e["confidence"], e["rms_energy"], e.get("slope", 0.0)] for e in ev], dtype=np.float64).reshape(-1, 3)
    return ints, flt


def test_note_events_match_reference_golden(dev, golden):
    """aegis_note_events on the golden INPUT arrays == the events the REAL aegis_engine_core/midi_logic.py produced
    from them (integers exact; confidence exact, rms dB and slope to rounding)."""
    from spectrogram_midi_b200 import midi_logic as M

    total = 0
    for name in ("track22050", "track44100", "clip7"):
        k = f"midi/{name}"
        sr = int(golden[f"{k}/sr"][0])
        ev = M.get_midi_events(rake_mask=golden[f"{k}/rake_mask"], f0=golden[f"{k}/f0"], voiced_flag=golden[f"{k}/voiced_flag"],
                               active_probs=golden[f"{k}/voiced_prob"], rms=golden[f"{k}/rms"], sr=sr, hop_length=512,
                               confidence_threshold=0.70)
        ints, flt = _event_rows(ev)
        np.testing.assert_array_equal(ints, golden[f"{k}/events"], err_msg=name)
        want = golden[f"{k}/event_float"]
        np.testing.assert_array_equal(flt[:, 0], want[:, 0])
        np.testing.assert_allclose(flt[:, 1], want[:, 1], rtol=0, atol=4e-6)
        np.testing.assert_allclose(flt[:, 2], want[:, 2], rtol=1e-9, atol=1e-12)
        assert all(isinstance(e["note"], int) and e["track"] in ("main", "safe") for e in ev)
        total += len(ev)
    assert total >= 10


def test_note_events_random_frames_equal_oracle(dev):
    """Adversarial frame arrays (short runs, rake holes, quiet frames, merges, hammer-ons, an exact quarter-tone tie)
    for a batch of clips at several rates: integer fields identical to the pinned restatement of midi_logic.py.

    Pitch drifts are continuous (random cents), not on pYIN's 10-cent grid: on the grid a 3-frame event that rises by
    one bin has a least-squares slope of EXACTLY the reference's 0.05 "bend" threshold, and which side np.polyfit
    (LAPACK gelsd) lands on is its own rounding noise (2.6 % of random grid events differ between np.polyfit and the
    closed form evaluated in the same float64) -- no implementation other than that LAPACK build can match it."""
    rng = np.random.default_rng(5)
    from spectrogram_midi_b200 import midi_logic as M

    E2hz = 82.4068892282175
    seen = set()
    for sr, hop, T in ((22050, 512, 700), (44100, 512, 900), (44100, 256, 600)):
        n = 6
        f0 = np.zeros((n, T))
        vf = np.zeros((n, T), bool)
        rms = np.full((n, T), 1e-4, np.float32)
        for c in range(n):
            t = 0
            while t < T:
                ln = int(rng.integers(1, 30))
                m = max(0, min(ln, T - t))
                if rng.random() < 0.8 and m:
                    base = 100.0 * rng.integers(0, 44) + rng.uniform(-12, 12)   # cents above E2, near a semitone centre
                    kind = rng.integers(0, 4)
                    x = np.arange(m)
                    if kind == 0:
                        cents = base + np.cumsum(rng.normal(0, 1.0, m))          # steady
                    elif kind == 1:
                        cents = base - 30 + 61.7 * x / max(m - 1, 1)            # bend up inside the note (never exactly 0.05 / frame)
                    elif kind == 2:
                        cents = base + 25 - 3.0 * x                               # slow slide down
                    else:
                        cents = base + 22.0 * np.sin(x * 1.3)                    # vibrato
                    f0[c, t:t + m] = E2hz * 2.0 ** (cents / 1200.0)
                    vf[c, t:t + m] = True
                    rms[c, t:t + m] = 10 ** rng.uniform(-2.3, -0.5) * rng.uniform(0.7, 1.0, m)
                t += ln + int(rng.integers(0, 3))
        f0[:, 5:9] = E2hz * 2.0 ** (5 / 120.0)   # exact quarter-tone tie: hz_to_midi == 40.5 up to rounding (constant pitch)
        vf[:, 5:9] = True
        f0[:, 9:12] = 0.0
        rms[:, 5:9] = 0.2
        vp = rng.uniform(0.4, 1.0, (n, T))
        rake = rng.random((n, T)) < 0.03
        rake[:, 4:10] = False
        for c in range(n):
            got = M.get_midi_events(rake[c], f0[c], vf[c], vp[c], rms[c], sr, hop, 0.7)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                ref = R.get_midi_events(rake[c], f0[c], vf[c], vp[c], rms[c], sr, hop, 0.7)
            gi, gf = _event_rows(got)
            ri, rf = _event_rows(ref)
            np.testing.assert_array_equal(gi, ri, err_msg=f"{sr}/{hop}/{c}")
            np.testing.assert_allclose(gf[:, 2], rf[:, 2], rtol=1e-8, atol=1e-11)
            assert len(got) > 10
            seen |= {int(v) for v in ri[:, 5]}
    assert seen >= {0, 1, 2, 3, 4, 5}   # none, vibrato, bend, slide, hammer_on, pull_off all occur


def test_note_events_batch_from_states(dev):
    """batch.note_events_batch (notes from the Viterbi states) == the numpy drop-in on each clip's host arrays."""
    from spectrogram_midi_b200 import midi_logic as M

    clips = corpus.clip_batch(3, 8.0, 22050, first_seed=40)
    res = P.batch.analyze_batch(_dev(clips, dev), sr=22050)
    ev = P.batch.note_events_batch(res, sr=22050, confidence_threshold=0.7)
    for c in range(3):
        host = P.batch.to_host(res, c)
        want = M.get_midi_events(host["rake_mask"], host["f0"], host["voiced_flag"], host["voiced_probs"], host["rms"], 22050, 512, 0.7)
        got = P.core.note_events_to_list(ev["events"], ev["n_events"], c)
        assert _event_rows(got)[0].tolist() == _event_rows(want)[0].tolist()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = R.get_midi_events(host["rake_mask"], host["f0"], host["voiced_flag"], host["voiced_probs"], host["rms"], 22050, 512, 0.7)
        assert _event_rows(got)[0].tolist() == _event_rows(ref)[0].tolist()


def test_rake_mask_fast_path_equals_general_kernel(dev):
    """The rake-mask-only kernel (four columns per thread, 16-byte loads; what the transcription step runs) and the general
    mel_post kernel (one column per thread, used whenever the dB image or the onset envelope is wanted too) give identical
    masks: synthetic mel images with broadband bursts of every length around the 20 ms .. 150 ms gate, placed across the
    192-column CTA borders and at both ends of the clip, for frame counts that end anywhere inside a group of four."""
    rng = np.random.default_rng(3)
    total = 0
    for T in (97, 386, 771, 1292, 1293, 1294, 1295):
        pitch = (T + 3) // 4 * 4
        buf = torch.zeros((5, 128, pitch), dtype=torch.float32, device=dev)
        mel = rng.random((5, 128, T)).astype(np.float32) * 1e-4
        mel[:, :20, :] += 0.5 * rng.random((5, 20, T)).astype(np.float32)          # a narrow-band "note": never a rake column
        starts = list(rng.integers(0, T, size=40)) + [0, 1, T - 8, T - 3, 185, 190, 191, 192, 380]
        for c in range(5):
            for t0 in starts:
                ln = int(rng.integers(1, 12))
                mel[c, :, max(0, t0): max(0, t0) + ln] = (0.3 + 0.7 * rng.random((128, 1))).astype(np.float32)[:, : 1]   # broadband burst
        buf[:, :, :T] = torch.from_numpy(mel).to(dev)
        view = buf[:, :, :T]                       # rows padded to a multiple of four floats: the fast path's precondition
        mx = view.amax(dim=(1, 2)).contiguous()
        fast = P.core.mel_post(view, mx, sr=22050, want_sdb=False, want_rake=True)
        slow = P.core.mel_post(view, mx, sr=22050, want_sdb=True, want_rake=True)
        assert torch.equal(fast["rake_mask"], slow["rake_mask"]), T
        tight = P.core.mel_post(torch.from_numpy(mel).to(dev), mx, sr=22050, want_sdb=False, want_rake=True)   # unpadded rows: general kernel
        if T % 4:
            assert torch.equal(tight["rake_mask"], slow["rake_mask"]), T
        total += int(fast["rake_mask"].sum())
    assert total > 50     # bursts that survive the run-length gate (overlapping bursts merge into runs that are too long)


def test_onset_fast_path_equals_general_kernel(dev):
    """The onset-only kernel (four columns per thread, 16-byte loads) and the general mel_post kernel (one column per
    thread, used whenever the dB image or the rake mask is wanted) give bit-identical envelopes, extrema and peaks,
    for frame counts that end anywhere inside a group of four columns."""
    for seconds in (1.0, 1.03, 1.05, 1.07, 7.3):
        clips = corpus.clip_batch(3, seconds, 22050, first_seed=int(seconds * 100))
        feat = P.core.stft_features(_dev(clips, dev), sr=22050, want_mag=False, want_mel=True, want_rms=False)
        fast = P.core.mel_post(feat["mel"], feat["mel_max"], sr=22050, want_sdb=False, want_rake=False, want_onset=True)
        full = P.core.mel_post(feat["mel"], feat["mel_max"], sr=22050, want_sdb=True, want_rake=True, want_onset=True)
        assert torch.equal(fast["onset_env"], full["onset_env"]), seconds
        assert torch.equal(fast["env_minmax"], full["env_minmax"])
        a = P.core.onset_peaks(fast["onset_env"], fast["env_minmax"], sr=22050)
        b = P.core.onset_peaks(full["onset_env"], full["env_minmax"], sr=22050)
        assert torch.equal(a["peaks"], b["peaks"]) and torch.equal(a["n_peaks"], b["n_peaks"])
        assert int(a["n_peaks"].sum()) == int(a["peaks"].sum())


# ---------------------------------------------------------------------------------- K8 v2 logic filter
def _fin_frames(seed, n, steady_grid=False):
    """Frame series of a made-up performance (see tests/golden/make_golden_financial_events.py): notes with jitter,
    vibrato, bends, slides and outliers, unvoiced gaps, rake frames, an energy envelope crossing the noise gate."""
    r = np.random.default_rng(seed)
    f0 = np.full(n, np.nan)
    voiced = np.zeros(n, bool)
    rms = np.full(n, 1e-4, dtype=np.float32)
    t = int(r.integers(0, 6))
    scale = np.array([40, 43, 45, 46, 47, 50, 52, 55, 57, 58, 59, 62, 64, 66, 69])
    while t < n:
        dur = int(r.integers(3, 60))
        m = float(r.choice(scale)) + 0.37 * float(r.random() < 0.15)
        k = np.arange(dur)
        kind = r.integers(0, 6)
        cents = np.zeros(dur) if (steady_grid and kind in (0, 5)) else r.normal(0, 4.0, dur)   # pYIN-like flat stretches
        if kind == 1:
            cents += 45.0 * np.sin(2 * np.pi * k / r.uniform(5, 9))
        elif kind == 2:
            cents += np.minimum(k * r.uniform(4, 14), 190.0)
        elif kind == 3:
            cents += k * r.uniform(-25, 25)
        elif kind == 4:
            cents[r.integers(0, dur, size=max(1, dur // 8))] += r.choice([-1, 1]) * r.uniform(60, 300)
        e = min(n, t + dur)
        f0[t:e] = 440.0 * 2.0 ** ((m - 69.0 + cents[: e - t] / 100.0) / 12.0)
        voiced[t:e] = True
        rms[t:e] = (r.uniform(0.004, 0.4) * np.exp(-k[: e - t] / r.uniform(15, 80))).astype(np.float32)
        t = e + int(r.choice([0, 0, 1, 2, 3, 6, 12]))
    voiced &= ~(r.random(n) < 0.03)
    probs = np.where(voiced, r.uniform(0.25, 1.0, n), r.uniform(0.0, 0.3, n))
    return r.random(n) < 0.02, np.where(voiced, f0, np.nan), voiced, probs, rms


def test_financial_note_events_match_reference_golden(dev, fin_golden):
    """aegis_fin_prepare / aegis_trend_filters / aegis_fin_events on the golden INPUT arrays == the events the REAL
    aegis_engine_core_v2/midi_logic_financial.py produced from them: every integer field (note, frames, velocity,
    track, articulation and slide labels, harmonic flag, key), confidences to 1e-12."""
    from oracle import financial_events as FE
    from spectrogram_midi_b200 import midi_logic_financial as MF

    g = fin_golden
    total = 0
    for n in [str(v) for v in g["fin/names"]]:
        (rake, f0, vf, vp, rms), sr, kw = FE.golden_case(g, n)
        ev = MF.get_midi_events_financial(rake_mask=rake, f0=f0, voiced_flag=vf, active_probs=vp, rms=rms, sr=sr, hop_length=512,
                                          use_financial=True, **kw)
        ints, conf, key = FE.events_rows(ev)
        np.testing.assert_array_equal(ints, g[f"fin/{n}/events"], err_msg=n)
        np.testing.assert_allclose(conf, g[f"fin/{n}/confidence"], rtol=1e-12, atol=0, err_msg=n)
        want_key = g[f"fin/{n}/key"]
        if want_key[0] < 0:
            assert key is None, n
        else:
            assert key is not None and key[:2] == (int(want_key[0]), int(want_key[1])), n
            assert abs(key[2] - want_key[2]) < 1e-12
        assert all(isinstance(e["note"], int) and e["track"] in ("main", "safe") for e in ev)
        total += len(ev)
    assert total > 200
    # use_financial=False is the reference's fallback branch: built too (see the fallback golden test)
    assert isinstance(MF.get_midi_events_financial(rake, f0, vf, vp, rms, sr, 512, use_financial=False), list)
    assert MF.get_midi_events_financial([], [], [], [], [], 22050, 512) == []
    assert MF.get_midi_events_financial([], [], [], [], [], 22050, 512, use_financial=False) == []


@pytest.mark.parametrize("kw", [
    {}, {"harmonic_tolerance": 0}, {"confidence_threshold": 0.55, "harmonic_tolerance": 0},
    {"min_note_duration_ms": 0, "sustain_ms": 150, "use_harmonic_filter": False}, {"noise_gate_db": -22},
])
def test_financial_note_events_batch_equals_oracle(dev, kw):
    """A batch of clips through core.note_events_financial (one launch sequence for all clips) == the pinned
    restatement clip by clip, including the adaptive threshold, flat pYIN-like pitch stretches (bands a few ulps
    wide) and a silent clip."""
    from oracle import financial_events as FE

    for sr, T in ((22050, 700), (44100, 1300)):
        clips = [_fin_frames(100 + 7 * i + T, T, steady_grid=(i % 2 == 1)) for i in range(7)]
        clips.append((np.zeros(T, bool), np.full(T, np.nan), np.zeros(T, bool), np.zeros(T), np.full(T, 1e-3, np.float32)))
        stack = [torch.from_numpy(np.stack([c[j] for c in clips])).to(dev) for j in range(5)]
        res = P.core.note_events_financial(*stack, sr=sr, hop_length=512, **kw)
        thr = res["threshold"].cpu().numpy()
        events = 0
        for c, (rake, f0, vf, vp, rms) in enumerate(clips):
            want = FE.get_midi_events_financial(rake, f0, vf, vp, rms, sr, 512, **kw)
            got = P.core.fin_events_to_list(res, c)
            wi, wc, wk = FE.events_rows(want)
            gi, gc, gk = FE.events_rows(got)
            np.testing.assert_array_equal(gi, wi, err_msg=f"clip {c} sr {sr}")
            np.testing.assert_allclose(gc, wc, rtol=1e-12, atol=0)
            assert (gk is None) == (wk is None) and (gk is None or (gk[:2] == wk[:2] and abs(gk[2] - wk[2]) < 1e-12))
            if "confidence_threshold" not in kw:
                S = FE.frame_series(f0, vf, vp)
                np.testing.assert_allclose(thr[c], FE.adaptive_confidence_threshold(S["combined"]), rtol=1e-13)
            events += len(want)
        assert events > 20


def test_financial_rsi_closed_form_equals_the_stepwise_walk(dev, monkeypatch):
    """K8 decides the RSI ghost-note verdicts from a closed form with an error bound and walks the 10 T-step density
    series only when a verdict is closer to the threshold than the bound.  AEGIS_FIN_EXACT_RSI=1 forces the walk (the
    round-1 path, itself pinned to the reference golden above): the event records must be byte-identical, over thresholds
    that keep everything, drop some and drop most, and on clips with 25 s of silence between notes (the decaying averages
    reach the subnormal range there)."""
    T = 1292
    clips = [_fin_frames(900 + i, T, steady_grid=(i % 3 == 1)) for i in range(24)]
    for i in (3, 9, 15):   # a long silence in the middle / at the start
        rake, f0, vf, vp, rms = (a.copy() for a in clips[i])
        lo, hi = (60, 1150) if i != 9 else (0, 1080)
        f0[lo:hi], vf[lo:hi], vp[lo:hi] = np.nan, False, 0.05
        clips[i] = (rake, f0, vf, vp, rms)
    stack = [torch.from_numpy(np.stack([c[j] for c in clips])).to(dev) for j in range(5)]
    dropped = []
    for thr in (70, 55.0, 50.0, 35.5, 20, 100.0, 0.0, 1e-5):
        monkeypatch.delenv("AEGIS_FIN_EXACT_RSI", raising=False)
        fast = P.core.note_events_financial(*stack, sr=22050, hop_length=512, rsi_threshold=thr, use_harmonic_filter=False)
        monkeypatch.setenv("AEGIS_FIN_EXACT_RSI", "1")
        walk = P.core.note_events_financial(*stack, sr=22050, hop_length=512, rsi_threshold=thr, use_harmonic_filter=False)
        monkeypatch.delenv("AEGIS_FIN_EXACT_RSI")
        assert torch.equal(fast["n_events"], walk["n_events"]), thr
        for c in range(len(clips)):
            n = int(fast["n_events"][c])
            assert torch.equal(fast["events"][c, :n], walk["events"][c, :n]), (thr, c)
        dropped.append(int(fast["n_events"].sum()))
    assert len(set(dropped)) >= 3 and min(dropped) < max(dropped), dropped   # the thresholds really filter
    print("events kept per threshold:", dropped)


@pytest.mark.parametrize("T", [8500, 21000])
def test_logic_filters_on_long_clips(dev, T):
    """K7 / K8 stage a clip's frame rows in shared memory when they fit (T <= 20 000 / 9 000 frames: up to 200 KB, opted in
    per launch); longer clips -- the one-hour recording of cfg4 -- walk the global rows.  Both logic filters on 8 500-frame
    (staged, large shared-memory carve-out) and 21 000-frame (unstaged) clips == the restatements."""
    from oracle import financial_events as FE
    from spectrogram_midi_b200 import midi_logic as M

    clips = [_fin_frames(4000 + i, T, steady_grid=bool(i)) for i in range(2)]
    stack = [torch.from_numpy(np.stack([c[j] for c in clips])).to(dev) for j in range(5)]
    res = P.core.note_events_financial(*stack, sr=22050, hop_length=512)
    for c, (rake, f0, vf, vp, rms) in enumerate(clips):
        want = FE.get_midi_events_financial(rake, f0, vf, vp, rms, 22050, 512)
        wi, wc, wk = FE.events_rows(want)
        gi, gc, gk = FE.events_rows(P.core.fin_events_to_list(res, c))
        np.testing.assert_array_equal(gi, wi, err_msg=f"v2 clip {c}")
        np.testing.assert_allclose(gc, wc, rtol=1e-12, atol=0)
        assert (gk is None) == (wk is None) and (gk is None or gk[:2] == wk[:2])
        assert len(want) > 50
        # v1 filter (K7) on continuous pitches: see test_note_events_random_frames_equal_oracle for why not pYIN's grid
        f0c = np.where(vf, np.nan_to_num(f0) * 2.0 ** (np.random.default_rng(c).uniform(-3, 3, T) / 1200.0), 0.0)
        got = M.get_midi_events(rake, f0c, vf, vp, rms, 22050, 512, 0.7)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = R.get_midi_events(rake, f0c, vf, vp, rms, 22050, 512, 0.7)
        assert len(ref) > 50
        assert _event_rows(got)[0].tolist() == _event_rows(ref)[0].tolist(), f"v1 clip {c}"


def test_financial_logic_filter_fallback_branch_matches_reference_golden(dev):
    """get_midi_events_financial(use_financial=False) -- the reference function's fallback branch
    (midi_logic_financial.py:178-196 + detect_articulations_financial per note; never taken by the v2 engine): host frame loop,
    Bollinger bands / MACD of every note's pitch slice from kernel K5.  Every field equals what the real file returns."""
    from spectrogram_midi_b200.midi_logic_financial import get_midi_events_financial

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fin_fallback_golden.npz"))
    tech = [None, "normal", "bend", "vibrato", "noise", "slide"]
    for name in g["fb/names"]:
        k, a = f"fb/{name}", g[f"fb/{name}/args"]
        ev = get_midi_events_financial(g[f"{k}/rake_mask"], g[f"{k}/f0"], g[f"{k}/voiced_flag"], g[f"{k}/voiced_prob"], g[f"{k}/rms"],
                                       int(a[0]), 512, confidence_threshold=None if np.isnan(a[1]) else float(a[1]),
                                       use_financial=False, noise_gate_db=float(a[2]), sustain_ms=float(a[3]), min_note_duration_ms=float(a[4]))
        ints = np.array([[e["note"], e["start"], e["end"], e["velocity"], int(e["track"] == "main"), tech.index(e.get("technique")),
                          int("technique" in e)] for e in ev], dtype=np.int64).reshape(-1, 7)
        np.testing.assert_array_equal(ints, g[f"{k}/events"], err_msg=str(name))
        np.testing.assert_array_equal(np.array([e["confidence"] for e in ev]), g[f"{k}/confidence"], err_msg=str(name))
        assert all(e["financial_artic"] is None and e["financial_slide"] is None for e in ev)


def test_financial_engine_note_events(dev):
    """AegisFinancialEngine.perception -> note_events (K1..K6, K5, K8) == the restatement on the same host arrays."""
    from oracle import financial_events as FE

    y = corpus.random_clip(21, 14.0, 22050)
    eng = P.engine.AegisFinancialEngine(sample_rate=22050)
    raw = eng.perception(y)
    got = eng.note_events(raw)
    voiced = raw["voiced_flag"] & ~raw["mute_mask"]
    want = FE.get_midi_events_financial(raw["rake_mask"], raw["f0"], voiced, raw["voiced_probs"], raw["rms"], 22050, 512)
    assert len(want) >= 3
    np.testing.assert_array_equal(FE.events_rows(got)[0], FE.events_rows(want)[0])
    res = P.batch.analyze_batch(_dev(y[None], dev), sr=22050, nan_to_num=False, with_guitar=True)
    ev = P.batch.note_events_financial_batch(res, sr=22050)
    np.testing.assert_array_equal(FE.events_rows(P.core.fin_events_to_list(ev, 0))[0], FE.events_rows(want)[0])


# ---------------------------------------------------------------------------------- whole engines, audio -> MIDI file
def test_engines_write_midi_files(dev, tmp_path):
    """AegisEngine.audio_to_midi -> extract_events (K7 + native writer) and AegisFinancialEngine.audio_to_midi_financial
    (K1..K6, K5 + K8, native writer) on one clip: the files decode to the messages the reference would hand to mido."""
    import io
    import warnings

    from oracle import financial_events as FE
    from oracle import midi_messages as MM

    sr = 22050
    y = corpus.random_clip(33, 12.0, sr)
    eng = P.engine.AegisEngine(sr)
    raw = eng.audio_to_midi(y, None)
    buf = io.BytesIO()
    events = eng.extract_events(raw, buf, confidence_threshold=0.7, vibrato_rate=6.0, midi_program=29)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = R.get_midi_events(raw["rake_mask"], raw["f0"], raw["voiced_flag"], raw["voiced_probs"], raw["rms"], sr, 512, 0.7)
    assert R.events_key(events) == R.events_key(want) and len(events) >= 3
    fmt, tpb, tracks = MM.read_smf(buf.getvalue())
    assert (fmt, tpb) == (1, 480) and tracks == MM.v1_tracks(events, sr, 512, midi_program=29, vibrato_rate=6.0)
    assert eng.extract_events(raw, None) is not None          # no file requested
    tab = eng.generate_tabs(events)
    assert len(tab) == sum(40 <= e["note"] <= 88 for e in events)
    assert os.path.getsize(eng.export_musicxml(tab, str(tmp_path / "clip.xml"))) > 400

    fin = P.engine.AegisFinancialEngine(sr)
    out = str(tmp_path / "fin.mid")
    assert fin.audio_to_midi_financial(y, out) == out
    rawf = fin.perception(y)
    voiced = rawf["voiced_flag"] & ~rawf["mute_mask"]
    wantf = FE.get_midi_events_financial(rawf["rake_mask"], rawf["f0"], voiced, rawf["voiced_probs"], rawf["rms"], sr, 512)
    _, _, tracks = MM.read_smf(open(out, "rb").read())
    assert tracks == MM.v2_tracks(wantf, sr, 512)
    assert fin.audio_to_midi_financial(np.zeros(0, np.float32), out) is None
    assert fin.audio_to_midi_financial(np.zeros(30000, np.float32), out) is None      # silence: no notes


# ---------------------------------------------------------------------------------- K9 ingest / resampler
@pytest.mark.parametrize("orig,target", [(44100, 22050), (88200, 22050), (48000, 22050), (22050, 44100), (16000, 22050), (48000, 44100), (22050, 22050)])
def test_resample_poly_is_bit_identical_to_scipy(dev, orig, target):
    """aegis_resample_poly == scipy.signal.resample_poly (librosa's res_type='polyphase') bit for bit on float32 audio:
    clip lengths that end mid-tile, a one-sample clip, a batch with a padded row stride."""
    import scipy.signal

    g = int(np.gcd(orig, target))
    up, down = target // g, orig // g
    rng = np.random.default_rng(orig + target)
    for n in (1, 777, 40000):
        y = rng.uniform(-1, 1, (3, n)).astype(np.float32)
        y[1] = corpus.random_clip(3, 2.0, 22050)[:n] if n <= 44100 else y[1]
        got = P.core.resample_poly(_dev(y, dev), orig, target).cpu().numpy()
        for c in range(3):
            ref = scipy.signal.resample_poly(y[c], up, down) if up != down else y[c]
            assert got[c].shape == ref.shape
            np.testing.assert_array_equal(got[c], ref, err_msg=f"{orig}->{target} n={n} clip {c}")
    wide = torch.zeros((2, 5000 + 13), dtype=torch.float32, device=dev)
    y = rng.uniform(-1, 1, (2, 5000)).astype(np.float32)
    wide[:, :5000] = torch.from_numpy(y).to(dev)
    got = P.core.resample_poly(wide[:, :5000], orig, target).cpu().numpy()
    np.testing.assert_array_equal(got[1], scipy.signal.resample_poly(y[1], up, down) if up != down else y[1])
    np.testing.assert_allclose(got[0], L.resample_polyphase(y[0], orig, target), rtol=0, atol=2e-6)   # and the oracle


def test_pcm_ingest_and_load(dev, tmp_path):
    """int16 stereo PCM -> float / 32768 -> channel mean -> 44.1 kHz to 22.05 kHz in one kernel == the same steps in
    numpy + scipy; librosa_compat.load does it for a WAV file; other res_types are refused, equal rates pass through."""
    import wave

    import scipy.signal

    lib = P.librosa_compat
    rng = np.random.default_rng(9)
    n = 30000
    pcm = rng.integers(-32768, 32768, size=(n, 2), dtype=np.int16)
    mono = (pcm.astype(np.float32) / 32768.0)
    mono = L.to_mono(mono.T).astype(np.float32)
    ref = scipy.signal.resample_poly(mono, 1, 2)
    got = P.core.resample_poly(torch.from_numpy(pcm.reshape(1, -1)).to(dev), 44100, 22050, n_channels=2)[0].cpu().numpy()
    np.testing.assert_array_equal(got, ref)
    path = str(tmp_path / "stereo44k.wav")
    with wave.open(path, "wb") as w:
        w.setnchannels(2)
        w.setsampwidth(2)
        w.setframerate(44100)
        w.writeframes(pcm.tobytes())
    y, sr = lib.load(path, sr=22050, res_type="polyphase")
    assert sr == 22050 and y.dtype == np.float32
    np.testing.assert_array_equal(y, ref)
    # no res_type: the reference's librosa.load would use soxr_hq; the substitution must be announced, never silent
    with pytest.warns(lib.ResampleDivergenceWarning):
        y_warn, _ = lib.load(path, sr=22050)
    np.testing.assert_array_equal(y_warn, ref)
    y_off, _ = lib.load(path, sr=22050, offset=0.1, duration=0.25, res_type="polyphase")
    a = int(0.1 * 44100)   # librosa truncates offset / duration to whole source frames
    np.testing.assert_array_equal(y_off, scipy.signal.resample_poly(mono[a:a + int(0.25 * 44100)], 1, 2))
    y_off2, _ = lib.load(path, sr=None, offset=0.30001, duration=0.10001)
    a2 = int(0.30001 * 44100)
    np.testing.assert_array_equal(y_off2, mono[a2:a2 + int(0.10001 * 44100)])
    y_native, sr_native = lib.load(path, sr=None)
    assert sr_native == 44100
    np.testing.assert_array_equal(y_native, mono)
    with pytest.raises(NotImplementedError):
        lib.load(path, sr=22050, res_type="soxr_hq")
    with pytest.raises(NotImplementedError):
        lib.resample(mono, orig_sr=44100, target_sr=22050, res_type="kaiser_best")
    with pytest.raises(NotImplementedError):
        lib.resample(mono, orig_sr=44100, target_sr=22050)     # librosa's default res_type is soxr_hq: refused, not substituted
    np.testing.assert_array_equal(lib.resample(mono, orig_sr=44100, target_sr=22050, res_type="polyphase"), ref)
    assert lib.resample(mono, orig_sr=22050, target_sr=22050, res_type="polyphase") is mono
    with pytest.raises(ValueError):
        P.core.resample_poly(torch.zeros((1, 8), device=dev), 44100.5, 22050)


def test_host_pipeline_chunking_does_not_change_results(dev):
    """HostPipeline cuts its last chunk into a half and two quarters (so that little work is left once the last byte has
    landed): 70 one-second clips in chunks of 16 (4 full chunks, then 3 + 2 + 1 clips... and 6 = 3 + 1 + 2) == one chunk
    for everything == the device-level calls on the whole batch."""
    sr, n_clips, n = 22050, 80, 22050
    clips = corpus.clip_batch(n_clips, 1.0, sr, first_seed=900)
    y = torch.from_numpy(clips).pin_memory()
    whole = {k: v.clone() for k, v in P.batch.HostPipeline(n_clips, n, sr=sr, device=dev, chunk_clips=n_clips).run(y).items()}
    for chunk in (16, 24):   # 5 pieces of 16 (the last tapered 8 + 4 + 4); 24, 24, 24, then 8 -> no taper below 16
        got = P.batch.HostPipeline(n_clips, n, sr=sr, device=dev, chunk_clips=chunk).run(y)
        for k in whole:
            assert torch.equal(got[k], whole[k]), (chunk, k)
    feat = P.core.stft_features(_dev(clips, dev), sr=sr, want_mag=False, want_rms=True)
    assert torch.equal(whole["rms"], feat["rms"].cpu())


def test_host_pipeline_pcm_ingest(dev):
    """HostPipeline fed with int16 PCM (same rate, and stereo 44.1 kHz) == HostPipeline fed with the float32 audio
    that numpy / scipy make of that PCM: the ingest kernel sits in front of an unchanged path."""
    import scipy.signal

    sr, n_clips, n = 22050, 5, 22050 * 2
    rng = np.random.default_rng(12)
    clips = corpus.clip_batch(n_clips, 2.0, sr, first_seed=70)
    pcm = np.round(clips * 32767.0).astype(np.int16)
    ref_pipe = P.batch.HostPipeline(n_clips, n, sr=sr, device=dev, chunk_clips=2)
    want = {k: v.clone() for k, v in ref_pipe.run(torch.from_numpy(pcm.astype(np.float32) / 32768.0).pin_memory()).items()}
    pipe = P.batch.HostPipeline(n_clips, n, sr=sr, device=dev, chunk_clips=2, pcm_rate=sr)
    got = pipe.run(torch.from_numpy(pcm).pin_memory())
    assert pipe.h2d_bytes * 2 == ref_pipe.h2d_bytes
    for k in want:
        assert torch.equal(got[k], want[k]), k
    # stereo 44.1 kHz source
    n_src = n * 2
    stereo = rng.integers(-20000, 20000, size=(n_clips, n_src, 2), dtype=np.int16)
    mono = ((stereo[..., 0].astype(np.float32) / 32768.0) + (stereo[..., 1].astype(np.float32) / 32768.0)) / np.float32(2)
    y = np.stack([scipy.signal.resample_poly(m, 1, 2) for m in mono])
    want = {k: v.clone() for k, v in ref_pipe.run(torch.from_numpy(y).pin_memory()).items()}
    pipe2 = P.batch.HostPipeline(n_clips, n, sr=sr, device=dev, chunk_clips=2, pcm_rate=44100, pcm_channels=2)
    got = pipe2.run(torch.from_numpy(stereo.reshape(n_clips, -1)).pin_memory())
    for k in want:
        assert torch.equal(got[k], want[k]), k
    with pytest.raises(ValueError):
        pipe2.run(torch.from_numpy(pcm).pin_memory())


def test_transcribe_pipeline_equals_batch_calls(dev):
    """TranscribePipeline (PCM host buffers in, pieces copied on side streams, frame-parallel kernels per piece, decoder +
    logic filter per group) returns exactly what analyze_batch + note_events_batch return for the whole batch."""
    sr, n_clips, n = 22050, 7, 22050 * 3
    clips = corpus.clip_batch(n_clips, 3.0, sr, first_seed=500)
    pcm = np.round(clips * 32767.0).astype(np.int16)
    y = torch.from_numpy(pcm.astype(np.float32) / 32768.0).to(dev)
    want = P.batch.analyze_batch(y, sr=sr)
    want_ev = P.batch.note_events_batch(want, sr=sr)
    for chunk, group in ((2, 4), (3, 3), (7, 7), (1, 2)):
        pipe = P.batch.TranscribePipeline(n_clips, n, sr=sr, device=dev, chunk_clips=chunk, group_clips=group, pcm=True)
        for _ in range(2):   # buffers are reused across runs
            got = pipe.run(torch.from_numpy(pcm).pin_memory())
        for k in ("rake_mask", "f0", "voiced_flag", "voiced_probs", "rms"):
            assert torch.equal(got[k], want[k].cpu()), (chunk, group, k)
        assert torch.equal(got["n_events"], want_ev["n_events"].cpu())
        for c in range(n_clips):
            ne = int(got["n_events"][c])
            assert ne > 0 and torch.equal(got["events"][c, :ne], want_ev["events"][c, :ne].cpu()), (chunk, group, c)
    pipe32 = P.batch.TranscribePipeline(n_clips, n, sr=sr, device=dev, chunk_clips=4, pcm=False)
    got = pipe32.run(torch.from_numpy(pcm.astype(np.float32) / 32768.0).pin_memory())
    assert torch.equal(got["f0"], want["f0"].cpu()) and pipe32.h2d_bytes == 2 * pipe.h2d_bytes
    with pytest.raises(ValueError):
        pipe32.run(torch.from_numpy(pcm).pin_memory())


# ---------------------------------------------------------------------------------- K6 guitar filters
def test_guitar_filters_match_reference_golden(dev, guitar_golden):
    """aegis_guitar_filters against the outputs of the real aegis_engine_core_v2/guitar_specific.py (bit-exact)."""
    from spectrogram_midi_b200 import guitar_specific as G

    g = guitar_golden
    names = sorted({k.split("/")[1] for k in g})
    assert len(names) >= 8
    code = {"clean": 0, "light": 1, "heavy": 2}
    for n in names:
        hop, sr = (int(v) for v in g[f"guitar/{n}/args"])
        res = G.apply_guitar_filters(g[f"guitar/{n}/f0"], g[f"guitar/{n}/voiced"], g[f"guitar/{n}/S_dB"], hop, sr, g[f"guitar/{n}/rake_in"])
        np.testing.assert_array_equal(res["f0"], g[f"guitar/{n}/out_f0"], err_msg=n)
        assert res["f0"].dtype == np.float64 and res["voiced"].dtype == bool and res["mute_mask"].dtype == bool
        np.testing.assert_array_equal(res["voiced"], g[f"guitar/{n}/out_voiced"], err_msg=n)
        np.testing.assert_array_equal(res["rake_mask"], g[f"guitar/{n}/out_rake"], err_msg=n)
        np.testing.assert_array_equal(res["mute_mask"], g[f"guitar/{n}/out_mute"], err_msg=n)
        assert code[res["distortion"]] == int(g[f"guitar/{n}/out_distortion"][0]), n
        # the individual static methods of the reference's class
        F = G.GuitarSpecificFilters
        f0o, vo = F.filter_subharmonic_noise(g[f"guitar/{n}/f0"], g[f"guitar/{n}/voiced"])
        np.testing.assert_array_equal(f0o, g[f"guitar/{n}/out_f0"])
        np.testing.assert_array_equal(vo, g[f"guitar/{n}/out_voiced"])
        np.testing.assert_array_equal(F.detect_palm_mute(g[f"guitar/{n}/S_dB"], hop, sr), g[f"guitar/{n}/out_mute"])
        np.testing.assert_array_equal(F.detect_rake_enhanced(g[f"guitar/{n}/S_dB"], hop, sr, g[f"guitar/{n}/rake_in"]), g[f"guitar/{n}/out_rake"])
        assert code[F.classify_distortion_level(g[f"guitar/{n}/S_dB"])] == int(g[f"guitar/{n}/out_distortion"][0])


def test_guitar_filters_batch_long_and_block_edges(dev):
    """Batch of ragged-content images longer than one 192-column block: every clip equals the oracle; runs and
    triggers that straddle block boundaries are handled by the halo."""
    rng = np.random.default_rng(11)
    T, n = 1000, 5
    imgs = rng.uniform(-62.0, -48.0, size=(n, 128, T)).astype(np.float32)
    imgs[:, 64:] -= 6.0
    for c in range(n):
        for s in rng.integers(1, T - 12, 60):
            ln = int(rng.integers(1, 7))
            if rng.random() < 0.5:   # mute-like run (also across columns 191/192, 383/384, ...)
                imgs[c, :64, s:s + ln] = rng.uniform(-75.0, -65.0, size=(64, ln))
                imgs[c, 64:, s:s + ln] = rng.uniform(-32.0, -28.0, size=(64, ln))
            else:                    # broadband burst
                imgs[c, :, s:s + ln] = rng.uniform(-14.0, -2.0, size=(128, ln))
        for edge in (192, 384, 576):
            imgs[c, :64, edge - 2:edge + 1] = -70.0
            imgs[c, 64:, edge - 2:edge + 1] = -30.0
            imgs[c, :, edge + 20 - 1:edge + 20 + 1] = -5.0
    f0 = rng.uniform(30.0, 500.0, (n, T))
    f0[rng.random((n, T)) < 0.3] = np.nan
    vf = ~np.isnan(f0)
    rk = rng.random((n, T)) < 0.05
    for sr, hop in ((22050, 512), (44100, 512), (44100, 256)):
        res = P.core.guitar_filters(torch.from_numpy(imgs).to(dev), sr=sr, hop_length=hop, f0=torch.from_numpy(f0).to(dev),
                                    voiced_flag=torch.from_numpy(vf).to(dev), rake_mask=torch.from_numpy(rk).to(dev))
        fired = 0
        for c in range(n):
            ref = R.apply_guitar_filters(f0[c], vf[c], imgs[c], hop, sr, rk[c])
            np.testing.assert_array_equal(res["f0"][c].cpu().numpy(), ref["f0"])
            np.testing.assert_array_equal(res["voiced"][c].cpu().numpy().astype(bool), ref["voiced"])
            np.testing.assert_array_equal(res["rake_mask"][c].cpu().numpy().astype(bool), ref["rake_mask"], err_msg=f"{sr}/{hop}/{c}")
            np.testing.assert_array_equal(res["mute_mask"][c].cpu().numpy().astype(bool), ref["mute_mask"], err_msg=f"{sr}/{hop}/{c}")
            assert P.core.DISTORTION_LABELS[int(res["distortion"][c])] == ref["distortion"]
            fired += int(ref["mute_mask"].sum()) + int((ref["rake_mask"] ^ rk[c]).sum())
        assert fired > 0


def test_financial_perception_applies_guitar_filters(dev):
    """AegisFinancialEngine.perception = the arrays audio_to_midi_financial builds (:105-154) incl. guitar filters."""
    y = corpus.test_track(22050, 0)
    eng = P.AegisFinancialEngine()
    got = eng.perception(y)
    plain = eng.perception(y, use_guitar_filters=False)
    ref = R.apply_guitar_filters(plain["f0"], plain["voiced_flag"], plain["S_dB"], 512, 22050, plain["rake_mask"])
    np.testing.assert_array_equal(got["f0"], ref["f0"])
    np.testing.assert_array_equal(got["rake_mask"], ref["rake_mask"])
    np.testing.assert_array_equal(got["mute_mask"], ref["mute_mask"])
    np.testing.assert_array_equal(got["voiced_flag"], ref["voiced"] & ~ref["mute_mask"])
    assert got["distortion"] == ref["distortion"]


# ---------------------------------------------------------------------------------- end to end: note events
def test_midi_note_events_identical_to_reference(dev, golden):
    """GPU perception -> the (pinned) restatement of midi_logic.get_midi_events == the events the REAL
    reference midi_logic produced from the oracle's perception arrays (tests/golden/make_golden.py)."""
    tech = {None: 0, "vibrato": 1, "bend": 2, "slide": 3, "hammer_on": 4, "pull_off": 5}
    inputs = {"track22050": (corpus.test_track(22050, 0, 10.0), 22050), "track44100": (corpus.test_track(44100, 0), 44100),
              "clip7": (corpus.random_clip(7, 12.0, 22050), 22050)}
    total = 0
    for name, (y, sr) in inputs.items():
        res = P.AegisEngine(sample_rate=sr).audio_to_midi(y, None)
        np.testing.assert_array_equal(res["rake_mask"], golden[f"midi/{name}/rake_mask"])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ev = R.get_midi_events(res["rake_mask"], res["f0"], res["voiced_flag"], res["voiced_probs"], res["rms"], sr, 512, 0.70)
        got = np.array([[e["note"], e["start"], e["end"], e["velocity"], int(e["track"] == "main"), tech[e.get("technique")]]
                        for e in ev], dtype=np.int64).reshape(-1, 6)
        want = golden[f"midi/{name}/events"]
        if name.startswith("track"):   # the generate_test_signal corpus: identical note events
            np.testing.assert_array_equal(res["voiced_flag"], golden[f"midi/{name}/voiced_flag"])
            np.testing.assert_array_equal(got, want, err_msg=name)
        else:                          # random clip: the reference's own fp32 rounding moves a few frames
            def roll(evs, n):
                r = np.zeros(n, np.int64)
                for e in evs:
                    r[e[1]: e[2] + 1] = e[0]
                return r
            n = len(res["f0"])
            agree = (roll(got, n) == roll(want, n)).mean()
            print(f"{name}: events {len(got)} vs {len(want)}, frame-level note agreement {agree:.4f}")
            assert agree >= 0.99 and len(got) == len(want)   # round 2: identical on this clip (was 0.97 with the FFT-based K2)
        total += len(ev)
    assert total >= 10


def test_full_batch_pipeline_with_trend(dev):
    clips = corpus.clip_batch(6, 5.0, 22050, first_seed=20)
    res = P.batch.analyze_batch(_dev(clips, dev), sr=22050, with_onsets=True, with_trend=True, with_sdb=True, nan_to_num=False)
    T = 1 + clips.shape[1] // 512
    assert res["f0"].shape == (6, T) and res["trend"].shape == (6, T) and res["S_dB"].shape == (6, 128, T)
    for c in (0, 5):
        f0_ref, vf_ref, _ = L.pyin(clips[c], fmin=E2, fmax=C6, sr=22050, hop_length=512)
        assert (res["voiced_flag"][c].cpu().numpy().astype(bool) == vf_ref).mean() > 0.97
        f0 = res["f0"][c].cpu().numpy()
        trend_ref, _ = R.multi_filter_consensus(f0)
        np.testing.assert_allclose(res["trend"][c].cpu().numpy(), trend_ref, rtol=1e-12, atol=1e-9)
    host = P.batch.to_host(res, 3, y=clips[3])
    assert set(host) >= {"rake_mask", "f0", "voiced_flag", "voiced_probs", "rms", "y", "onset_frames", "trend"}


def test_synth_corpus_kernel(dev):
    plan = corpus.plan_events(8, 3.0, 22050, first_seed=0)
    y = P.core.synth_events(8, int(3.0 * 22050), plan, dev)
    assert y.shape == (8, 66150) and torch.isfinite(y).all()
    assert torch.allclose(y.abs().amax(dim=1), torch.full((8,), 0.9, device=dev), atol=1e-6)
    f0, vf, _ = P.librosa_compat.pyin(y[0].cpu().numpy(), fmin=E2, fmax=C6, sr=22050)
    assert vf.mean() > 0.3  # plucked strings are pitched


# ---------------------------------------------------------------------------------- long clip in windows
@pytest.mark.parametrize("sr,mode", [(22050, "exact"), (22050, "windowed"), (44100, "exact")])
def test_long_clip_windows_equal_the_full_clip_result(dev, sr, mode):
    """distributed.analyze_long_clip on one GPU with 5 windows (the per-rank windows of a multi-GPU run,
    processed in turn): frame-local outputs and -- in exact mode -- the decode are bit-identical to analysing
    the whole clip at once (the window starts are aligned to the kernels' 8-frame tiles)."""
    from spectrogram_midi_b200 import distributed as D

    y = np.concatenate([corpus.random_clip(300 + i, 6.0, sr) for i in range(4)])
    ref = P.AegisEngine(sample_rate=sr).audio_to_midi(y, None)
    res = D.analyze_long_clip(y, sr=sr, mode=mode, windows_per_rank=5, burn_seconds=2.0)
    np.testing.assert_array_equal(res["rake_mask"], ref["rake_mask"])
    np.testing.assert_array_equal(res["voiced_probs"], ref["voiced_probs"])
    np.testing.assert_array_equal(res["rms"], ref["rms"])
    if mode == "exact":
        np.testing.assert_array_equal(res["voiced_flag"], ref["voiced_flag"])
        np.testing.assert_array_equal(np.nan_to_num(res["f0"]), ref["f0"])
        # the note events assembled by the event gather (here: one rank, five windows) == the engine's own list
        res2 = D.analyze_long_clip(y, sr=sr, mode="exact", windows_per_rank=5, return_events=True)
        eng = P.AegisEngine(sample_rate=sr)
        want = eng.note_events(ref)
        got = res2["events"]
        assert res2["events_local"] == len(got) == len(want) > 0
        assert [(int(r["note"]), int(r["start"]), int(r["end"]), int(r["velocity"]), bool(r["track"])) for r in got] == \
            [(e["note"], e["start"], e["end"], e["velocity"], e["track"] == "main") for e in want]
    else:
        assert (res["voiced_flag"] == ref["voiced_flag"]).mean() >= 0.99


def test_one_hour_clip_at_baseline_cfg4_size(dev):
    """BASELINE cfg4 at full size on one GPU: a 3600 s, 44.1 kHz recording (158 760 000 samples, T = 310 079 frames)
    tiled from seeded cfg2-style segments, analysed as 8 overlapping windows (the per-rank windows of an 8-GPU run,
    processed in turn).  Too long to compare with the oracle, so size-independent properties are checked:
    frame count, frame-local outputs identical between the two stitching modes, windowed decode (2 s burn-in) against the
    exact single-chain decode, and spot windows against a direct analysis of the same audio."""
    import time

    from spectrogram_midi_b200 import distributed as D

    sr, seg_s, n_seg = 44100, 30.0, 120
    plan = corpus.plan_events(n_seg, seg_s, sr, first_seed=7000)
    y = P.core.synth_events(n_seg, int(seg_s * sr), plan, dev).reshape(-1).cpu().numpy()
    assert y.shape[0] == 158_760_000
    t0 = time.time()
    ex = D.analyze_long_clip(y, sr=sr, mode="exact", windows_per_rank=8)
    t1 = time.time()
    wi = D.analyze_long_clip(y, sr=sr, mode="windowed", windows_per_rank=8, burn_seconds=2.0)
    t2 = time.time()
    T = 1 + y.shape[0] // 512
    assert T == 310_079 and all(len(ex[k]) == T for k in ("rake_mask", "f0", "voiced_flag", "voiced_probs", "rms"))
    for k in ("rake_mask", "voiced_probs", "rms"):
        np.testing.assert_array_equal(ex[k], wi[k])
    agree = float((ex["voiced_flag"] == wi["voiced_flag"]).mean())
    both = ex["voiced_flag"] & wi["voiced_flag"]
    f0_same = float((ex["f0"][both] == wi["f0"][both]).mean())
    print(f"cfg4 full size: exact {t1 - t0:.1f} s, windowed {t2 - t1:.1f} s ({3600 / (t2 - t1):.0f} x realtime on one GPU), "
          f"voiced agreement {agree:.6f}, f0 identical on {f0_same:.6f} of common voiced frames, voiced {ex['voiced_flag'].mean():.3f}")
    assert agree >= 0.9995 and f0_same >= 0.9995
    assert 0.2 < ex["voiced_flag"].mean() < 0.99
    # a window from the middle analysed directly: frame-local outputs are identical (8-frame aligned start)
    f_lo = 8 * 20000
    s_lo = f_lo * 512
    seg = y[s_lo - 1024 : s_lo - 1024 + 512 * 400 + 2048]
    direct = P.core.stft_features(_dev(seg, dev), sr=sr, want_mag=False, want_rms=True, center=False)
    np.testing.assert_array_equal(direct["rms"][0].cpu().numpy()[:400], ex["rms"][f_lo : f_lo + 400])
