"""Host-side constant tables for the CUDA kernels (window, twiddles, mel triangles, pYIN priors,
HMM log-transition bands).  Computed once per configuration with numpy/scipy in float64 using the
same expressions librosa uses, then uploaded; the kernels never recompute them.

librosa equivalents (third-party): ``filters.get_window`` (A.2),
``filters.mel`` (A.3), ``pyin`` steps 2/6/8/10/11 (A.5), ``onset_detect`` defaults (A.6).
"""
from __future__ import annotations

import functools
from dataclasses import dataclass

import numpy as np
import scipy.signal
import scipy.stats

N_FFT = 2048
N_BINS = N_FFT // 2 + 1
TINY64 = float(np.finfo(np.float64).tiny)
LOG_TINY64 = float(np.log(TINY64))  # -708.3964185322641


# ------------------------------------------------------------------------------------------
# unit helpers (librosa.note_to_hz / hz_to_midi / midi_to_hz)
# ------------------------------------------------------------------------------------------
_PITCH_CLASS = {"C": 0, "D": 2, "E": 4, "F": 5, "G": 7, "A": 9, "B": 11}


def note_to_midi(note: str) -> int:
    acc = 0
    pos = 1
    while pos < len(note) and note[pos] in "#b":
        acc += 1 if note[pos] == "#" else -1
        pos += 1
    octave = int(note[pos:]) if pos < len(note) else 0
    return 12 * (octave + 1) + _PITCH_CLASS[note[0].upper()] + acc


def midi_to_hz(midi):
    return 440.0 * (2.0 ** ((np.asanyarray(midi, dtype=np.float64) - 69.0) / 12.0))


def note_to_hz(note: str) -> float:
    return float(midi_to_hz(note_to_midi(note)))


def hz_to_midi(freq):
    return 12.0 * (np.log2(np.asanyarray(freq, dtype=np.float64)) - np.log2(440.0)) + 69.0


# ------------------------------------------------------------------------------------------
# STFT tables
# ------------------------------------------------------------------------------------------
@functools.lru_cache(maxsize=None)
def hann_window(n_fft: int = N_FFT) -> np.ndarray:
    """Periodic Hann, float64 (``get_window('hann', n_fft, fftbins=True)``)."""
    return scipy.signal.get_window("hann", n_fft, fftbins=True)


@functools.lru_cache(maxsize=None)
def fft_twiddles(n: int = N_FFT) -> np.ndarray:
    """float32 [n, 2]: (cos, -sin)(2*pi*k/n), rounded once from float64."""
    k = np.arange(n, dtype=np.float64)
    ang = 2.0 * np.pi * k / n
    tw = np.stack([np.cos(ang), -np.sin(ang)], axis=1)
    # exact values at the quadrant points
    for q, (c, s) in {0: (1.0, 0.0), n // 4: (0.0, -1.0), n // 2: (-1.0, 0.0), 3 * n // 4: (0.0, 1.0)}.items():
        tw[q] = (c, s)
    return np.ascontiguousarray(tw.astype(np.float32))


def _hz_to_mel(f):
    f = np.atleast_1d(np.asarray(f, dtype=np.float64)).copy()
    lin = 200.0 / 3
    brk_hz = 1000.0
    brk_mel = brk_hz / lin
    step = np.log(6.4) / 27.0
    m = f / lin
    hi = f >= brk_hz
    m[hi] = brk_mel + np.log(f[hi] / brk_hz) / step
    return m


def _mel_to_hz(m):
    m = np.atleast_1d(np.asarray(m, dtype=np.float64)).copy()
    lin = 200.0 / 3
    brk_hz = 1000.0
    brk_mel = brk_hz / lin
    step = np.log(6.4) / 27.0
    f = lin * m
    hi = m >= brk_mel
    f[hi] = brk_hz * np.exp(step * (m[hi] - brk_mel))
    return f


@functools.lru_cache(maxsize=None)
def mel_filterbank(sr: float, n_fft: int = N_FFT, n_mels: int = 128) -> np.ndarray:
    """Dense float32 [n_mels, 1+n_fft//2] Slaney mel basis (``librosa.filters.mel`` defaults)."""
    edges_hz = _mel_to_hz(np.linspace(_hz_to_mel(0.0)[0], _hz_to_mel(sr / 2.0)[0], n_mels + 2))
    bins_hz = np.fft.rfftfreq(n_fft, 1.0 / sr)
    width = np.diff(edges_hz)
    dist = edges_hz[:, None] - bins_hz[None, :]
    fb = np.zeros((n_mels, bins_hz.size), dtype=np.float32)
    for b in range(n_mels):
        rising = -dist[b] / width[b]
        falling = dist[b + 2] / width[b + 1]
        fb[b] = np.maximum(0, np.minimum(rising, falling))
    fb *= (2.0 / (edges_hz[2:] - edges_hz[:-2]))[:, None]
    return fb


@dataclass(frozen=True)
class SparseMel:
    n_mels: int
    start: np.ndarray   # int32 [n_mels] first non-zero FFT bin of each band
    length: np.ndarray  # int32 [n_mels]
    offset: np.ndarray  # int32 [n_mels] into weights
    weights: np.ndarray  # float32 [nnz]
    max_len: int
    seg_start: np.ndarray  # int32 [n_mels + 2]: bins [seg_start[j], seg_start[j+1]) lie between band edges j and j+1
    rise_fall: np.ndarray  # float32 [n_bins, 2]: weight of bin k in band seg(k) (rising side) and band seg(k)-1 (falling)


@functools.lru_cache(maxsize=None)
def sparse_mel(sr: float, n_fft: int = N_FFT, n_mels: int = 128) -> SparseMel:
    fb = mel_filterbank(sr, n_fft, n_mels)
    start = np.zeros(n_mels, np.int32)
    length = np.zeros(n_mels, np.int32)
    offset = np.zeros(n_mels, np.int32)
    chunks = []
    pos = 0
    for b in range(n_mels):
        nz = np.flatnonzero(fb[b])
        if nz.size:
            start[b] = nz[0]
            length[b] = nz[-1] - nz[0] + 1
            chunks.append(fb[b, nz[0] : nz[-1] + 1])
        offset[b] = pos
        pos += int(length[b])
    weights = np.concatenate(chunks).astype(np.float32) if chunks else np.zeros(1, np.float32)
    max_len = int(length.max())
    # Every FFT bin lies in at most two triangles: the falling side of band j-1 and the rising side of
    # band j, where j is the "segment" between band edges j and j+1.  mel[b] = sum_{k in seg b} rise[k] P[k]
    # + sum_{k in seg b+1} fall[k] P[k]: each bin is read once.  Built from the dense basis and verified
    # to reproduce it exactly.
    n_bins = fb.shape[1]
    edges_hz = _mel_to_hz(np.linspace(_hz_to_mel(0.0)[0], _hz_to_mel(sr / 2.0)[0], n_mels + 2))
    bins_hz = np.fft.rfftfreq(n_fft, 1.0 / sr)
    seg = np.clip(np.searchsorted(edges_hz, bins_hz, side="right") - 1, 0, n_mels)
    rf = np.zeros((n_bins, 2), np.float32)
    for k in range(n_bins):
        nz = np.flatnonzero(fb[:, k])
        if nz.size == 2:
            seg[k] = nz[1]
        elif nz.size == 1:
            b = int(nz[0])
            seg[k] = b if bins_hz[k] < edges_hz[b + 1] else b + 1
        elif nz.size > 2:
            raise RuntimeError("mel basis is not a partition of triangles (bin in more than two bands)")
        j = int(seg[k])
        if j < n_mels:
            rf[k, 0] = fb[j, k]
        if j >= 1:
            rf[k, 1] = fb[j - 1, k]
    if np.any(np.diff(seg) < 0):
        raise RuntimeError("mel segments are not monotone in frequency")
    check = np.zeros_like(fb)
    for k in range(n_bins):
        j = int(seg[k])
        if j < n_mels:
            check[j, k] = rf[k, 0]
        if j >= 1:
            check[j - 1, k] = rf[k, 1]
    if not np.array_equal(check, fb):
        raise RuntimeError("segment form of the mel basis does not reproduce the dense basis")
    seg_start = np.searchsorted(seg, np.arange(n_mels + 2), side="left").astype(np.int32)
    return SparseMel(n_mels, start, length, offset, weights, max_len, seg_start, np.ascontiguousarray(rf))


# ------------------------------------------------------------------------------------------
# pYIN tables
# ------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class PyinConfig:
    sr: float
    hop_length: int
    frame_length: int
    win_length: int
    fmin: float
    fmax: float
    min_period: int
    max_period: int
    n_lags: int            # max_period - min_period + 1
    n_pitch_bins: int
    bins_per_semitone: int
    n_thresholds: int
    half_width: int        # transition band half width (bins)
    max_troughs: int       # worst-case number of local minima of the CMND curve
    no_trough_prob: float
    thresholds: np.ndarray      # float64 [n_thresholds]  (linspace(0,1,n+1)[1:])
    beta_probs: np.ndarray      # float64 [n_thresholds]
    beta_cumsum: np.ndarray     # float64 [n_thresholds+1]  np.sum(beta_probs[:m])
    boltz_fact: np.ndarray      # float64 [max_troughs+1]   (1-e^-l)/(1-e^-lN)
    boltz_exp: np.ndarray       # float64 [max_troughs+1]   e^-lk
    freqs: np.ndarray           # float64 [n_pitch_bins]
    lt_variants: np.ndarray     # float64 [n_var, 2, 2*hw+1]: distinct source-row bands (same, switch voicing)
    row_variant: np.ndarray     # int32 [n_pitch_bins]: band variant of each source bin
    n_interior_variants: int    # the first n variants cover all untruncated (interior) rows
    log_init_unvoiced: float
    log_tiny: float

    @functools.cached_property
    def cache_key(self) -> tuple:
        """Identifies every device table derived from this configuration: all scalar parameters plus a digest of the
        table bytes (beta / Boltzmann / HMM parameters change the tables without changing sr, hop, fmin or fmax)."""
        import hashlib

        h = hashlib.blake2b(digest_size=16)
        for a in (self.thresholds, self.beta_probs, self.beta_cumsum, self.boltz_fact, self.boltz_exp, self.freqs,
                  self.lt_variants, self.row_variant):
            h.update(np.ascontiguousarray(a).tobytes())
        return (self.sr, self.hop_length, self.frame_length, self.fmin, self.fmax, self.min_period, self.max_period,
                self.n_pitch_bins, self.bins_per_semitone, self.n_thresholds, self.half_width, self.max_troughs,
                self.no_trough_prob, self.n_interior_variants, h.hexdigest())


def _transition_local_triangle(n_states: int, width: int) -> np.ndarray:
    """``librosa.sequence.transition_local(n, width, window='triangle', wrap=False)``."""
    tri = scipy.signal.get_window("triangle", width, fftbins=False)
    base = np.zeros(n_states)
    left = (n_states - width) // 2
    base[left : left + width] = tri
    out = np.zeros((n_states, n_states))
    half = width // 2
    for i in range(n_states):
        row = np.roll(base, n_states // 2 + i + 1)
        row[min(n_states, i + half + 1) :] = 0
        row[: max(0, i - half)] = 0
        out[i] = row
    out /= out.sum(axis=1, keepdims=True)
    return out


def _transition_loop(n_states: int, p_stay: float) -> np.ndarray:
    out = np.empty((n_states, n_states))
    for i in range(n_states):
        out[i] = (1.0 - p_stay) / (n_states - 1)
        out[i, i] = p_stay
    return out


@functools.lru_cache(maxsize=None)
def pyin_config(
    sr: float,
    hop_length: int,
    fmin: float,
    fmax: float,
    frame_length: int = 2048,
    n_thresholds: int = 100,
    beta_parameters: tuple = (2, 18),
    boltzmann_parameter: float = 2,
    resolution: float = 0.1,
    max_transition_rate: float = 35.92,
    switch_prob: float = 0.01,
    no_trough_prob: float = 0.01,
) -> PyinConfig:
    win_length = frame_length // 2
    min_period = int(max(np.floor(sr / fmax), 1))
    max_period = int(min(np.ceil(sr / fmin), frame_length - win_length - 1))
    n_lags = max_period - min_period + 1
    bps = int(np.ceil(1.0 / resolution))
    n_bins = int(np.floor(12 * bps * np.log2(fmax / fmin))) + 1

    th = np.linspace(0, 1, n_thresholds + 1)
    beta_probs = np.diff(scipy.stats.beta.cdf(th, beta_parameters[0], beta_parameters[1]))
    beta_cumsum = np.array([np.sum(beta_probs[:m]) for m in range(n_thresholds + 1)])

    max_troughs = n_lags // 2 + 1
    lam = float(boltzmann_parameter)
    ns = np.arange(max_troughs + 1, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        fact = (1 - np.exp(-lam)) / (1 - np.exp(-lam * ns))
    fact[0] = 0.0
    expk = np.exp(-lam * ns)
    # the factorisation must reproduce scipy's pmf bit for bit (scipy computes fact*exp(-l*k), clipped)
    for n_t in (1, 2, 3, 7, max_troughs):
        k = np.arange(n_t)
        ref = scipy.stats.boltzmann.pmf(k, lam, n_t)
        if not np.array_equal(ref, np.clip(fact[n_t] * expk[:n_t], 0, 1)):
            raise RuntimeError("boltzmann prior factorisation differs from scipy.stats.boltzmann.pmf")

    semis = round(max_transition_rate * 12 * hop_length / sr)
    width = semis * bps + 1
    hw = width // 2
    if n_bins < 2 * width:
        raise ValueError("pitch range too narrow for the banded transition tables")
    local = _transition_local_triangle(n_bins, width)
    full = np.kron(_transition_loop(2, 1 - switch_prob), local)
    with np.errstate(divide="ignore"):
        log_full = np.log(full + TINY64)
    same = log_full[:n_bins, :n_bins]
    switch = log_full[:n_bins, n_bins:]
    if not np.array_equal(switch, log_full[n_bins:, :n_bins]) or not np.array_equal(same, log_full[n_bins:, n_bins:]):
        raise RuntimeError("voicing blocks of the transition matrix are not symmetric")

    def band(mat, k):  # values log_trans[k, k-hw .. k+hw]; out-of-range destinations = -inf
        row = np.full(width, -np.inf)
        lo, hi = max(0, k - hw), min(n_bins - 1, k + hw)
        row[lo - (k - hw) : hi - (k - hw) + 1] = mat[k, lo : hi + 1]
        return row

    # librosa row-normalises with a pairwise np.sum over the whole (mostly zero) row, so rows that
    # are mathematically identical differ in the last ulp depending on where the band sits.  Keep
    # every distinct band ("variant") so the decoder sees exactly librosa's numbers; interior rows
    # collapse to a handful of variants, the 2*hw truncated edge rows each get their own.
    variants: dict = {}
    bands = []
    row_variant = np.zeros(n_bins, np.int32)
    order = list(range(hw, n_bins - hw)) + list(range(hw)) + list(range(n_bins - hw, n_bins))
    n_interior = 0
    for k in order:
        pair = np.stack([band(same, k), band(switch, k)])
        key = pair.tobytes()
        if key not in variants:
            variants[key] = len(bands)
            bands.append(pair)
        row_variant[k] = variants[key]
        if hw <= k < n_bins - hw:
            n_interior = max(n_interior, row_variant[k] + 1)
    lt_variants = np.ascontiguousarray(np.stack(bands))
    # anything outside the band must be exactly log(tiny)
    probe = same[n_bins // 2].copy()
    probe[n_bins // 2 - hw : n_bins // 2 + hw + 1] = LOG_TINY64
    if not np.all(probe == LOG_TINY64):
        raise RuntimeError("out-of-band transitions are not log(tiny)")

    freqs = fmin * 2 ** (np.arange(n_bins) / (12 * bps))
    return PyinConfig(
        sr=float(sr), hop_length=int(hop_length), frame_length=frame_length, win_length=win_length,
        fmin=float(fmin), fmax=float(fmax), min_period=min_period, max_period=max_period, n_lags=n_lags,
        n_pitch_bins=n_bins, bins_per_semitone=bps, n_thresholds=n_thresholds, half_width=hw,
        max_troughs=max_troughs, no_trough_prob=float(no_trough_prob),
        thresholds=np.ascontiguousarray(th[1:]), beta_probs=beta_probs, beta_cumsum=beta_cumsum,
        boltz_fact=fact, boltz_exp=expk, freqs=freqs,
        lt_variants=lt_variants, row_variant=row_variant, n_interior_variants=int(n_interior),
        log_init_unvoiced=float(np.log(1 / n_bins + TINY64)), log_tiny=LOG_TINY64,
    )


# ------------------------------------------------------------------------------------------
# onset / rake / filter constants
# ------------------------------------------------------------------------------------------
def onset_peak_params(sr: float, hop_length: int) -> dict:
    """``librosa.onset.onset_detect`` defaults -> integer frame counts for ``util.peak_pick``."""
    return dict(
        pre_max=int(0.03 * sr // hop_length),
        post_max=int(0.00 * sr // hop_length + 1),
        pre_avg=int(0.10 * sr // hop_length),
        post_avg=int(0.10 * sr // hop_length + 1),
        wait=int(0.03 * sr // hop_length),
        delta=0.07,
    )


def rake_frame_limits(hop_length: int, sr: float) -> tuple:
    """(min_frames, max_frames) exactly as ``aegis_engine_core/vision.py:23-25`` computes them."""
    ms_per_frame = (hop_length / sr) * 1000
    return int(10 / ms_per_frame), int(30 / ms_per_frame)


@functools.lru_cache(maxsize=None)
def savgol_coeffs(window: int = 11, polyorder: int = 3) -> np.ndarray:
    """Correlation weights used by ``scipy.signal.savgol_filter`` (financial_filters.py:46-51)."""
    return np.ascontiguousarray(scipy.signal.savgol_coeffs(window, polyorder)[::-1])


@functools.lru_cache(maxsize=16)
def resample_poly_design(up: int, down: int):
    """(taps float32 incl. the `* up` gain, n_pre_pad, n_pre_remove) of scipy.signal.resample_poly(x_float32, up, down)
    for a reduced ratio: Kaiser(5.0) windowed sinc of 2 * 10 * max(up, down) + 1 taps with cutoff 1 / max(up, down),
    zero-padded in front so that output sample 0 sits at the filter centre (scipy/signal/_signaltools.py)."""
    max_rate = max(up, down)
    half_len = 10 * max_rate
    h = scipy.signal.firwin(2 * half_len + 1, 1.0 / max_rate, window=("kaiser", 5.0)).astype(np.float32)
    h *= up
    n_pre_pad = down - half_len % down
    n_pre_remove = (half_len + n_pre_pad) // down
    return h, int(n_pre_pad), int(n_pre_remove)

