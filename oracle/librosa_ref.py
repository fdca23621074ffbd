"""CPU restatement of the librosa subset on the Aegis hot path.  TEST INFRASTRUCTURE ONLY.

**Parity unpinned** (see ``oracle/__init__.py``): librosa itself is a third-party, unpinned
dependency of the reference (``/root/reference/requirements.txt:1``) that is absent from
``/root/reference`` and not installable here.  This file restates its published algorithm
(librosa >= 0.10 semantics) in plain numpy/scipy with the same dtypes
librosa uses at every step (float32 frames, float64 window, float32 CMND, float64 HMM).

Reference call sites that reach each function (``/root/reference`` paths):

* ``stft`` / ``melspectrogram`` / ``power_to_db`` : ``aegis_engine.py:25-26``,
  ``aegis_engine_financial.py:45-51``
* ``rms``                                         : ``aegis_engine.py:70``, ``aegis_engine_financial.py:154``
* ``pyin``                                        : ``aegis_engine.py:63,67,190,216``,
  ``aegis_engine_core/worker.py:9-15``, ``aegis_engine_financial.py:63-69``
* ``hz_to_midi`` / ``amplitude_to_db``            : ``aegis_engine_core/midi_logic.py:17,51,69``
* ``onset_strength`` / ``onset_detect``           : no call site in the reference;
  named by BASELINE.json's north_star; defined here from librosa's documented algorithm.
"""
from __future__ import annotations

import numpy as np
import scipy.fft
import scipy.signal
import scipy.stats

try:  # librosa's own Viterbi is numba-jitted; use numba when present so the CPU baseline is fair
    import numba
except Exception:  # pragma: no cover
    numba = None

TINY64 = float(np.finfo(np.float64).tiny)
TINY32 = float(np.finfo(np.float32).tiny)

# --------------------------------------------------------------------------------------------
# A.1 unit helpers
# --------------------------------------------------------------------------------------------
_NOTE_BASE = {"C": 0, "D": 2, "E": 4, "F": 5, "G": 7, "A": 9, "B": 11}


def note_to_midi(note: str) -> int:
    """'E2' -> 40, 'C6' -> 84 (librosa.note_to_midi for plain/sharp/flat names)."""
    pitch = _NOTE_BASE[note[0].upper()]
    i = 1
    while i < len(note) and note[i] in "#b!":
        pitch += {"#": 1, "b": -1, "!": -1}[note[i]]
        i += 1
    octave = int(note[i:]) if i < len(note) else 0
    return 12 * (octave + 1) + pitch


def midi_to_hz(m):
    return 440.0 * (2.0 ** ((np.asanyarray(m, dtype=np.float64) - 69.0) / 12.0))


def note_to_hz(note: str) -> float:
    return float(midi_to_hz(note_to_midi(note)))


def hz_to_midi(f):
    return 12.0 * (np.log2(np.asanyarray(f, dtype=np.float64)) - np.log2(440.0)) + 69.0


def power_to_db(S, ref=1.0, amin=1e-10, top_db=80.0):
    S = np.asarray(S)
    magnitude = np.abs(S) if np.iscomplexobj(S) else S
    ref_value = ref(magnitude) if callable(ref) else np.abs(ref)
    log_spec = 10.0 * np.log10(np.maximum(amin, magnitude))
    log_spec -= 10.0 * np.log10(np.maximum(amin, ref_value))
    if top_db is not None:
        log_spec = np.maximum(log_spec, log_spec.max() - top_db)
    return log_spec


def amplitude_to_db(S, ref=1.0, amin=1e-5, top_db=80.0):
    magnitude = np.abs(np.asarray(S))
    ref_value = ref(magnitude) if callable(ref) else np.abs(ref)
    power = np.square(magnitude, out=magnitude.copy())
    return power_to_db(power, ref=ref_value**2, amin=amin**2, top_db=top_db)


# --------------------------------------------------------------------------------------------
# A.2 framing + STFT
# --------------------------------------------------------------------------------------------
def frame_count(n_samples: int, hop_length: int = 512, n_fft: int = 2048, center: bool = True) -> int:
    if center:
        return 1 + n_samples // hop_length
    return 1 + (n_samples - n_fft) // hop_length


def frame_signal(y, frame_length=2048, hop_length=512, center=True, dtype=np.float32):
    """[frame_length, T] float32 view-like array of (zero centre-padded) frames."""
    y = np.asarray(y, dtype=dtype)
    if center:
        y = np.pad(y, frame_length // 2, mode="constant")
    T = 1 + (len(y) - frame_length) // hop_length
    idx = np.arange(frame_length)[:, None] + hop_length * np.arange(T)[None, :]
    return y[idx]


def hann_window(n_fft=2048):
    return scipy.signal.get_window("hann", n_fft, fftbins=True)  # float64, periodic


def stft(y, n_fft=2048, hop_length=512, center=True):
    """complex64 [1 + n_fft//2, T]; float64 window x float32 frames, FFT in float64, cast."""
    frames = frame_signal(y, n_fft, hop_length, center)
    win = hann_window(n_fft)[:, None]
    X = scipy.fft.rfft(win * frames, axis=0)  # float64 -> complex128
    return X.astype(np.complex64)


def stft_magnitude(y, n_fft=2048, hop_length=512, center=True):
    return np.abs(stft(y, n_fft, hop_length, center))  # float32


# --------------------------------------------------------------------------------------------
# A.3 mel filterbank + melspectrogram
# --------------------------------------------------------------------------------------------
def _hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    if f.ndim:
        m = f >= min_log_hz
        mels[m] = min_log_mel + np.log(f[m] / min_log_hz) / logstep
    elif f >= min_log_hz:
        mels = min_log_mel + np.log(f / min_log_hz) / logstep
    return mels


def _mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    if m.ndim:
        lg = m >= min_log_mel
        freqs[lg] = min_log_hz * np.exp(logstep * (m[lg] - min_log_mel))
    elif m >= min_log_mel:
        freqs = min_log_hz * np.exp(logstep * (m - min_log_mel))
    return freqs


def mel_filterbank(sr, n_fft=2048, n_mels=128, fmin=0.0, fmax=None):
    """float32 [n_mels, 1 + n_fft//2], Slaney scale + Slaney area norm (librosa.filters.mel)."""
    if fmax is None:
        fmax = float(sr) / 2
    n_bins = 1 + n_fft // 2
    weights = np.zeros((n_mels, n_bins), dtype=np.float32)
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2 : n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


def melspectrogram(y, sr, n_fft=2048, hop_length=512, n_mels=128, power=2.0, center=True):
    S = stft_magnitude(y, n_fft, hop_length, center) ** power  # float32
    M = mel_filterbank(sr, n_fft, n_mels)
    return np.einsum("ft,mf->mt", S, M, optimize=True).astype(np.float32, copy=False)


def load_audio_features(y, sr, n_fft=2048, hop_length=512):
    """The two librosa calls of AegisEngine.load_audio after decoding (aegis_engine.py:25-26)."""
    S = melspectrogram(y, sr, n_fft, hop_length)
    return power_to_db(S, ref=np.max)


# --------------------------------------------------------------------------------------------
# A.4 RMS
# --------------------------------------------------------------------------------------------
def rms(y, frame_length=2048, hop_length=512, center=True):
    x = frame_signal(y, frame_length, hop_length, center)
    power = np.mean(np.abs(x) ** 2, axis=-2, keepdims=True)
    return np.sqrt(power)  # float32 [1, T]


# --------------------------------------------------------------------------------------------
# A.5 pYIN
# --------------------------------------------------------------------------------------------
def pyin_periods(sr, fmin, fmax, frame_length=2048, win_length=None):
    if win_length is None:
        win_length = frame_length // 2
    min_period = int(max(np.floor(sr / fmax), 1))
    max_period = int(min(np.ceil(sr / fmin), frame_length - win_length - 1))
    return min_period, max_period


def n_pitch_bins_for(fmin, fmax, resolution=0.1):
    n_bins_per_semitone = int(np.ceil(1.0 / resolution))
    return int(np.floor(12 * n_bins_per_semitone * np.log2(fmax / fmin))) + 1, n_bins_per_semitone


def yin_difference(y_frames, frame_length=2048, win_length=1024):
    """d[tau, t] for tau=0..frame_length-win_length-1... (librosa: 1024 lags), float32, via FFT."""
    a = scipy.fft.rfft(y_frames, frame_length, axis=-2)
    b = scipy.fft.rfft(y_frames[..., win_length:0:-1, :], frame_length, axis=-2)
    acf_frames = scipy.fft.irfft(a * b, frame_length, axis=-2)[..., win_length:, :]
    acf_frames[np.abs(acf_frames) < 1e-6] = 0
    energy_frames = np.cumsum(y_frames**2, axis=-2)
    energy_frames = energy_frames[..., win_length:, :] - energy_frames[..., :-win_length, :]
    energy_frames[np.abs(energy_frames) < 1e-6] = 0
    return energy_frames[..., :1, :] + energy_frames - 2 * acf_frames


def cmnd(y_frames, frame_length, win_length, min_period, max_period):
    """Cumulative-mean-normalised difference, float32 [max_period-min_period+1, T]."""
    yin_frames = yin_difference(y_frames, frame_length, win_length)
    tiny = np.finfo(yin_frames.dtype).tiny
    yin_numerator = yin_frames[..., min_period : max_period + 1, :]
    tau_range = np.arange(1, max_period + 1)[:, None]
    cumulative_mean = np.cumsum(yin_frames[..., 1 : max_period + 1, :], axis=-2) / tau_range
    yin_denominator = cumulative_mean[..., min_period - 1 : max_period, :]
    return yin_numerator / (yin_denominator + tiny)


def parabolic_interpolation(x):
    """Shifts in [-1, 1] along axis 0; first and last rows 0 (librosa.pitch._parabolic_interpolation)."""
    shifts = np.zeros_like(x)
    a = x[2:] + x[:-2] - 2 * x[1:-1]
    b = (x[2:] - x[:-2]) / 2
    with np.errstate(divide="ignore", invalid="ignore"):
        s = -b / a
    s[np.abs(b) >= np.abs(a)] = 0
    shifts[1:-1] = s
    return shifts


def beta_threshold_probs(n_thresholds=100, beta_parameters=(2, 18)):
    thresholds = np.linspace(0, 1, n_thresholds + 1)
    beta_cdf = scipy.stats.beta.cdf(thresholds, beta_parameters[0], beta_parameters[1])
    return thresholds, np.diff(beta_cdf)


def localmin_rows(x):
    """librosa.util.localmin along axis 0 with pyin's first-row override."""
    is_trough = np.zeros(x.shape, dtype=bool)
    is_trough[1:-1] = (x[1:-1] < x[:-2]) & (x[1:-1] <= x[2:])
    is_trough[-1] = x[-1] < x[-2]
    is_trough[0] = x[0] < x[1]
    return is_trough


def pyin_observations(
    yin_frames,
    parabolic_shifts,
    sr,
    thresholds,
    boltzmann_parameter,
    beta_probs,
    no_trough_prob,
    min_period,
    fmin,
    n_pitch_bins,
    n_bins_per_semitone,
    return_sparse=False,
):
    """Observation matrix [2*n_pitch_bins, T] float64 and voiced_prob [1, T] (librosa __pyin_helper)."""
    yin_probs = np.zeros_like(yin_frames)
    is_trough_all = localmin_rows(yin_frames)
    for i in range(yin_frames.shape[1]):
        (trough_index,) = np.nonzero(is_trough_all[:, i])
        if len(trough_index) == 0:
            continue
        trough_heights = yin_frames[trough_index, i]
        trough_thresholds = np.less.outer(trough_heights, thresholds[1:])
        trough_positions = np.cumsum(trough_thresholds, axis=0) - 1
        n_troughs = np.count_nonzero(trough_thresholds, axis=0)
        with np.errstate(all="ignore"):
            trough_prior = scipy.stats.boltzmann.pmf(trough_positions, boltzmann_parameter, n_troughs)
        trough_prior[~trough_thresholds] = 0
        probs = trough_prior.dot(beta_probs)
        global_min = np.argmin(trough_heights)
        n_thresholds_below_min = np.count_nonzero(~trough_thresholds[global_min, :])
        probs[global_min] += no_trough_prob * np.sum(beta_probs[:n_thresholds_below_min])
        yin_probs[trough_index, i] = probs

    yin_period, frame_index = np.nonzero(yin_probs)
    period_candidates = min_period + yin_period
    period_candidates = period_candidates + parabolic_shifts[yin_period, frame_index]
    f0_candidates = sr / period_candidates
    bin_index = 12 * n_bins_per_semitone * np.log2(f0_candidates / fmin)
    bin_index = np.clip(np.round(bin_index), 0, n_pitch_bins).astype(int)
    observation_probs = np.zeros((2 * n_pitch_bins, yin_frames.shape[1]))
    observation_probs[bin_index, frame_index] = yin_probs[yin_period, frame_index]
    voiced_prob = np.clip(np.sum(observation_probs[:n_pitch_bins, :], axis=0, keepdims=True), 0, 1)
    observation_probs[n_pitch_bins:, :] = (1 - voiced_prob) / n_pitch_bins
    if return_sparse:
        return observation_probs, voiced_prob, yin_probs
    return observation_probs, voiced_prob


def transition_local_triangle(n_states, width):
    """librosa.sequence.transition_local(n_states, width, window='triangle', wrap=False)."""
    transition = np.zeros((n_states, n_states), dtype=np.float64)
    win = scipy.signal.get_window("triangle", width, fftbins=False)
    lpad = (n_states - width) // 2
    for i in range(n_states):
        if n_states >= width:
            trans_row = np.zeros(n_states)
            trans_row[lpad : lpad + width] = win  # util.pad_center
        else:  # pragma: no cover - never reached for pyin sizes
            raise ValueError("n_states < width")
        trans_row = np.roll(trans_row, n_states // 2 + i + 1)
        trans_row[min(n_states, i + width // 2 + 1) :] = 0
        trans_row[: max(0, i - width // 2)] = 0
        transition[i] = trans_row
    transition /= transition.sum(axis=1, keepdims=True)
    return transition


def transition_loop(n_states, prob):
    transition = np.empty((n_states, n_states), dtype=np.float64)
    p = np.full(n_states, prob, dtype=np.float64)
    for i, p_i in enumerate(p):
        transition[i] = (1.0 - p_i) / (n_states - 1)
        transition[i, i] = p_i
    return transition


def pyin_transition(n_pitch_bins, n_bins_per_semitone, sr, hop_length, max_transition_rate=35.92, switch_prob=0.01):
    max_semitones_per_frame = round(max_transition_rate * 12 * hop_length / sr)
    transition_width = max_semitones_per_frame * n_bins_per_semitone + 1
    transition = transition_local_triangle(n_pitch_bins, transition_width)
    t_switch = transition_loop(2, 1 - switch_prob)
    return np.kron(t_switch, transition), transition_width


def _viterbi_core_py(log_prob, log_trans, log_p_init):
    n_steps, n_states = log_prob.shape
    value = np.zeros((n_steps, n_states), dtype=np.float64)
    ptr = np.zeros((n_steps, n_states), dtype=np.uint16)
    state = np.zeros(n_steps, dtype=np.uint16)
    value[0] = log_prob[0] + log_p_init
    log_trans_t = np.ascontiguousarray(log_trans.T)
    for t in range(1, n_steps):
        trans_out = value[t - 1] + log_trans_t  # [j, k]
        for j in range(n_states):
            ptr[t, j] = np.argmax(trans_out[j])
            value[t, j] = log_prob[t, j] + trans_out[j, ptr[t, j]]
    state[-1] = np.argmax(value[-1])
    for t in range(n_steps - 2, -1, -1):
        state[t] = ptr[t + 1, state[t + 1]]
    return state


if numba is not None:
    _viterbi_core = numba.njit(cache=False, nogil=True)(_viterbi_core_py)
else:  # pragma: no cover
    def _viterbi_core(log_prob, log_trans, log_p_init):
        n_steps, n_states = log_prob.shape
        value = log_prob[0] + log_p_init
        ptr = np.zeros((n_steps, n_states), dtype=np.uint16)
        lt = np.ascontiguousarray(log_trans.T)
        for t in range(1, n_steps):
            trans_out = value[None, :] + lt
            p = np.argmax(trans_out, axis=1)
            ptr[t] = p
            value = log_prob[t] + trans_out[np.arange(n_states), p]
        state = np.zeros(n_steps, dtype=np.uint16)
        state[-1] = np.argmax(value)
        for t in range(n_steps - 2, -1, -1):
            state[t] = ptr[t + 1, state[t + 1]]
        return state


def viterbi(prob, transition, p_init):
    """librosa.sequence.viterbi: prob [n_states, T] -> states [T] (log-space, +tiny inside logs)."""
    eps = TINY64
    log_trans = np.log(transition + eps)
    log_prob = np.ascontiguousarray(np.log(prob.T + eps))
    log_p_init = np.log(p_init + eps)
    return _viterbi_core(log_prob, log_trans, log_p_init)


def pyin(
    y,
    *,
    fmin,
    fmax,
    sr=22050,
    frame_length=2048,
    win_length=None,
    hop_length=None,
    n_thresholds=100,
    beta_parameters=(2, 18),
    boltzmann_parameter=2,
    resolution=0.1,
    max_transition_rate=35.92,
    switch_prob=0.01,
    no_trough_prob=0.01,
    fill_na=np.nan,
    center=True,
    return_intermediates=False,
    frames_dtype=np.float32,
):
    """librosa.pyin (>= 0.10, pad_mode='constant').  Returns (f0, voiced_flag, voiced_prob).

    ``frames_dtype=np.float64`` evaluates the same algorithm in exact-ish arithmetic (librosa itself
    runs the difference function in float32); the tests use it to measure how much of any
    disagreement is the reference's own rounding noise.
    """
    if win_length is None:
        win_length = frame_length // 2
    if hop_length is None:
        hop_length = frame_length // 4
    y = np.asarray(y, dtype=np.float32)
    y_frames = frame_signal(y, frame_length, hop_length, center, dtype=frames_dtype)
    min_period, max_period = pyin_periods(sr, fmin, fmax, frame_length, win_length)
    yin_frames = cmnd(y_frames, frame_length, win_length, min_period, max_period)
    parabolic_shifts = parabolic_interpolation(yin_frames)
    thresholds, beta_probs = beta_threshold_probs(n_thresholds, beta_parameters)
    n_pitch_bins, n_bins_per_semitone = n_pitch_bins_for(fmin, fmax, resolution)
    observation_probs, voiced_prob, yin_probs = pyin_observations(
        yin_frames, parabolic_shifts, sr, thresholds, boltzmann_parameter, beta_probs,
        no_trough_prob, min_period, fmin, n_pitch_bins, n_bins_per_semitone, return_sparse=True,
    )
    transition, _ = pyin_transition(n_pitch_bins, n_bins_per_semitone, sr, hop_length, max_transition_rate, switch_prob)
    p_init = np.zeros(2 * n_pitch_bins)
    p_init[n_pitch_bins:] = 1 / n_pitch_bins
    states = viterbi(observation_probs, transition, p_init)
    freqs = fmin * 2 ** (np.arange(n_pitch_bins) / (12 * n_bins_per_semitone))
    f0 = freqs[states % n_pitch_bins]
    voiced_flag = states < n_pitch_bins
    if fill_na is not None:
        f0[~voiced_flag] = fill_na
    if return_intermediates:
        return f0, voiced_flag, voiced_prob[0], dict(
            yin_frames=yin_frames, parabolic_shifts=parabolic_shifts, yin_probs=yin_probs,
            observation_probs=observation_probs, states=states, min_period=min_period,
            max_period=max_period, n_pitch_bins=n_pitch_bins,
        )
    return f0, voiced_flag, voiced_prob[0]


# --------------------------------------------------------------------------------------------
# A.6 onset strength / detection (no reference call site; north_star-defined)
# --------------------------------------------------------------------------------------------
def onset_strength(y=None, sr=22050, S=None, n_fft=2048, hop_length=512, lag=1, center=True):
    if S is None:
        S = power_to_db(melspectrogram(y, sr, n_fft, hop_length))
    S = np.atleast_2d(S)
    onset_env = S[..., lag:] - S[..., :-lag]
    onset_env = np.maximum(0.0, onset_env)
    onset_env = np.mean(onset_env, axis=-2)
    pad_width = lag
    if center:
        pad_width += n_fft // (2 * hop_length)
    onset_env = np.pad(onset_env, (int(pad_width), 0), mode="constant")
    if center:
        onset_env = onset_env[: S.shape[-1]]
    return onset_env


def peak_pick(x, pre_max, post_max, pre_avg, post_avg, delta, wait):
    """Truncated-window greedy peak picker (librosa >= 0.10.2 numba version), float64 means."""
    pre_max, post_max, pre_avg, post_avg, wait = (int(v) for v in (pre_max, post_max, pre_avg, post_avg, wait))
    n_x = x.shape[0]
    peaks = np.zeros(n_x, dtype=bool)
    if n_x == 0:
        return np.zeros(0, dtype=np.int64)
    x64 = x.astype(np.float64)
    peaks[0] = x64[0] >= np.max(x64[: min(post_max, n_x)])
    peaks[0] &= x64[0] >= np.mean(x64[: min(post_avg, n_x)]) + delta
    n = wait + 1 if peaks[0] else 1
    while n < n_x:
        maxn = np.max(x64[max(0, n - pre_max) : min(n + post_max, n_x)])
        if x64[n] != maxn:
            n += 1
            continue
        avgn = np.mean(x64[max(0, n - pre_avg) : min(n + post_avg, n_x)])
        if not (x64[n] >= avgn + delta):
            n += 1
            continue
        peaks[n] = True
        n += wait + 1
    return np.flatnonzero(peaks)


def onset_detect_params(sr, hop_length):
    return dict(
        pre_max=0.03 * sr // hop_length,
        post_max=0.00 * sr // hop_length + 1,
        pre_avg=0.10 * sr // hop_length,
        post_avg=0.10 * sr // hop_length + 1,
        wait=0.03 * sr // hop_length,
        delta=0.07,
    )


def onset_detect(y=None, sr=22050, onset_envelope=None, hop_length=512, normalize=True):
    if onset_envelope is None:
        onset_envelope = onset_strength(y=y, sr=sr, hop_length=hop_length)
    onset_envelope = np.asarray(onset_envelope)
    if not onset_envelope.any() or not np.all(np.isfinite(onset_envelope)):
        return np.array([], dtype=np.int64)
    if normalize:
        onset_envelope = onset_envelope - np.min(onset_envelope)
        onset_envelope = onset_envelope / (np.max(onset_envelope) + np.finfo(onset_envelope.dtype).tiny)
    return peak_pick(onset_envelope, **onset_detect_params(sr, hop_length))


# --------------------------------------------------------------------------------------------
# librosa.load's tail: to_mono + resample(res_type='polyphase') (aegis_engine.py:24, aegis_engine_financial.py:45)
# Third-party: librosa.resample(res_type='polyphase') calls scipy.signal.resample_poly (scipy 1.18 here); its published
# algorithm is restated below from first principles (zero stuffing, FIR, decimation) in float64 and pinned against
# scipy.signal.resample_poly itself in tests/test_oracle.py.  librosa's DEFAULT res_type (soxr_hq, libsoxr) is absent
# from this image: parity for it is unpinned and the product refuses it.
# --------------------------------------------------------------------------------------------
def kaiser_lowpass(up, down):
    """scipy.signal.firwin(2 * 10 * max(up, down) + 1, 1 / max(up, down), window=('kaiser', 5.0)), written out."""
    import scipy.special

    max_rate = max(up, down)
    half_len = 10 * max_rate
    n = np.arange(-half_len, half_len + 1, dtype=np.float64)
    fc = 1.0 / max_rate
    h = fc * np.sinc(fc * n)
    h *= scipy.special.i0(5.0 * np.sqrt(np.clip(1.0 - (n / half_len) ** 2, 0.0, None))) / scipy.special.i0(5.0)
    return h / h.sum(), half_len


def resample_polyphase(y, orig_sr, target_sr):
    """librosa.resample(y, orig_sr=, target_sr=, res_type='polyphase') in float64 arithmetic, float32 result."""
    y = np.asarray(y, dtype=np.float64)
    g = int(np.gcd(int(orig_sr), int(target_sr)))
    up, down = int(target_sr) // g, int(orig_sr) // g
    if up == down == 1:
        return y.astype(np.float32)
    h, half_len = kaiser_lowpass(up, down)
    h = h.astype(np.float32).astype(np.float64) * up        # scipy matches the filter's dtype to x before the gain
    n_out = -(-len(y) * up // down)
    stuffed = np.zeros(len(y) * up)
    stuffed[::up] = y
    full = np.convolve(stuffed, h)                          # full[n + half_len] is centred on stuffed sample n
    idx = half_len + np.arange(n_out) * down
    out = np.zeros(n_out)
    ok = idx < len(full)
    out[ok] = full[idx[ok]]
    return out.astype(np.float32)


def to_mono(y):
    """librosa.to_mono: mean over the channel axis of [channels, n]."""
    y = np.asarray(y)
    return np.mean(y, axis=0) if y.ndim > 1 else y

