"""Device-level batched operators: torch CUDA tensors in, torch CUDA tensors out.

torch is plumbing only (device memory, streams); every numeric step is a hand-written sm_100a
kernel in libaegis_b200.so reached through the C ABI.  Inputs are batches of equally long clips
``y[n_clips, n_samples]`` (float32, contiguous rows).  No function here has a CPU path.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch

from . import _native as nat
from . import tables

N_FFT = tables.N_FFT
N_BINS = tables.N_BINS

_table_cache: dict = {}


def _dev_tensor(key, device, make):
    k = (key, str(device))
    t = _table_cache.get(k)
    if t is None:
        t = torch.from_numpy(np.ascontiguousarray(make())).to(device)
        _table_cache[k] = t
    return t


def _audio_ptr(y: torch.Tensor) -> int:
    """Device pointer of the audio; an empty tensor has none, so hand the kernels a dummy word."""
    if y.numel():
        return y.data_ptr()
    return _dev_tensor("dummy", y.device, lambda: np.zeros(4, np.float32)).data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _check_audio(y: torch.Tensor) -> torch.Tensor:
    if not isinstance(y, torch.Tensor) or not y.is_cuda:
        raise nat.AegisNativeError("expected a CUDA tensor: the Aegis B200 path has no CPU implementation")
    if y.dim() == 1:
        y = y[None]
    if y.dtype != torch.float32:
        y = y.float()
    if y.stride(-1) != 1:
        y = y.contiguous()
    return y


def frame_count(n_samples: int, hop_length: int, center: bool = True) -> int:
    if center:
        return 1 + n_samples // hop_length
    return 1 + (n_samples - N_FFT) // hop_length if n_samples >= N_FFT else 0


def frame_pitch(n_frames: int) -> int:
    """Row pitch (in frames) of per-frame images in HBM: rows start on 32-byte sector boundaries, so the 8-frame
    row chunks the STFT kernel writes are whole sectors (a dense [.., T] row pitch costs ~20 % more DRAM traffic
    when T % 8 != 0: measured with ncu, profiles/r1_stft_notes.md)."""
    return (int(n_frames) + 7) // 8 * 8


def alloc_frames(shape_prefix, n_frames: int, device, dtype=torch.float32) -> torch.Tensor:
    """[*shape_prefix, n_frames] view of a buffer whose rows are ``frame_pitch(n_frames)`` apart."""
    full = torch.empty((*shape_prefix, frame_pitch(n_frames)), dtype=dtype, device=device)
    return full[..., :n_frames]


def _check_fft(n_fft: int, hop_length: int):
    if n_fft != N_FFT:
        raise NotImplementedError(f"n_fft/frame_length={n_fft}: the B200 kernels are built for 2048 (aegis_engine.py:17)")
    if hop_length < 4 or hop_length > 512 or hop_length % 4:
        raise NotImplementedError(f"hop_length={hop_length}: supported 4..512, multiple of 4")


# ----------------------------------------------------------------------------------------------
# K1
# ----------------------------------------------------------------------------------------------
def stft_features(y: torch.Tensor, *, sr: Optional[float] = None, hop_length: int = 512, n_fft: int = N_FFT,
                  center: bool = True, n_mels: int = 128, want_mag: bool = True, want_mel: bool = False,
                  want_rms: bool = False, mag_out: Optional[torch.Tensor] = None,
                  pad: Optional[int] = None, n_frames: Optional[int] = None) -> dict:
    """Fused STFT magnitude / mel power / RMS of a batch of clips.

    Returns a dict with the requested keys: ``mag`` [n_clips, 1025, T], ``mel`` [n_clips, n_mels, T],
    ``mel_max`` [n_clips], ``rms`` [n_clips, T] (all float32, librosa layouts; ``mag`` / ``mel`` are views whose rows
    start on 32-byte boundaries, see ``frame_pitch``).  ``pad`` / ``n_frames``
    override the framing (frame t covers samples [t*hop - pad, t*hop - pad + 2048)): used when ``y`` is a
    window cut out of a longer recording (``distributed.analyze_long_clip``).
    """
    _check_fft(n_fft, hop_length)
    y = _check_audio(y)
    dev = y.device
    n_clips, n_samples = y.shape
    T = frame_count(n_samples, hop_length, center) if n_frames is None else int(n_frames)
    out: dict = {"n_frames": T}
    P = nat.StftParams()
    P.y, P.clip_stride, P.n_samples, P.n_clips = _audio_ptr(y), y.stride(0), n_samples, n_clips
    P.hop, P.pad, P.n_frames = hop_length, ((N_FFT // 2 if center else 0) if pad is None else int(pad)), T
    P.window = _dev_tensor("hann32", dev, lambda: tables.hann_window().astype(np.float32)).data_ptr()
    P.twiddle = _dev_tensor("twiddle", dev, tables.fft_twiddles).data_ptr()
    if want_mag:
        mag = mag_out if mag_out is not None else alloc_frames((n_clips, N_BINS), T, dev)
        if mag.shape != (n_clips, N_BINS, T) or mag.dtype != torch.float32 or mag.stride(2) != 1 and T > 1:
            raise ValueError("mag_out must be float32 [n_clips, 1025, T] with unit stride along T")
        P.mag, P.mag_clip_stride, P.mag_row_stride = mag.data_ptr(), mag.stride(0), mag.stride(1)
        out["mag"] = mag
    if want_mel:
        if sr is None:
            raise ValueError("sr is required for the mel projection")
        sm = tables.sparse_mel(float(sr), N_FFT, n_mels)
        key = ("mel", float(sr), n_mels)
        P.mel_seg_start = _dev_tensor(key + ("seg",), dev, lambda: sm.seg_start).data_ptr()
        P.mel_rise_fall = _dev_tensor(key + ("rf",), dev, lambda: sm.rise_fall).data_ptr()
        mel = alloc_frames((n_clips, n_mels), T, dev)
        mel_max = torch.zeros((n_clips,), dtype=torch.float32, device=dev)
        P.n_mels = n_mels
        P.mel, P.mel_clip_stride, P.mel_row_stride = mel.data_ptr(), mel.stride(0), mel.stride(1)
        P.mel_max = mel_max.data_ptr()
        out["mel"], out["mel_max"] = mel, mel_max
    if want_rms:
        rms = torch.empty((n_clips, T), dtype=torch.float32, device=dev)
        P.rms, P.rms_clip_stride = rms.data_ptr(), rms.stride(0)
        out["rms"] = rms
    nat.call("aegis_stft_fused", P, _stream())
    return out


# ----------------------------------------------------------------------------------------------
# K4
# ----------------------------------------------------------------------------------------------
def mel_post(mel: torch.Tensor, mel_max: Optional[torch.Tensor], *, sr: float, hop_length: int = 512,
             rake_ratio: float = 0.6, want_sdb: bool = True, want_rake: bool = True,
             want_onset: bool = False, n_fft: int = N_FFT, center: bool = True, lag: int = 1,
             ref_power: Optional[torch.Tensor] = None, input_is_db: bool = False) -> dict:
    """dB conversion (ref = clip max unless ``ref_power``, top_db 80), rake mask, onset envelope.

    With ``input_is_db`` the first argument already is a dB image [n_clips, n_mels, T] and only the
    rake mask is produced (the call shape of ``detect_rake_patterns``).
    """
    if lag != 1:
        raise NotImplementedError("onset lag != 1")
    if mel.dtype != torch.float32 or mel.stride(2) != 1:
        mel = mel.float().contiguous()
    n_clips, n_mels, T = mel.shape
    dev = mel.device
    P = nat.MelPostParams()
    P.mel, P.mel_clip_stride, P.mel_row_stride = mel.data_ptr(), mel.stride(0), mel.stride(1)
    P.n_mels, P.n_clips, P.n_frames = n_mels, n_clips, T
    P.mel_max = None if mel_max is None else mel_max.data_ptr()
    P.ref_power = None if ref_power is None else ref_power.data_ptr()
    P.input_is_db = int(bool(input_is_db))
    if input_is_db:
        want_sdb = want_onset = False
    lo, hi = tables.rake_frame_limits(hop_length, sr)
    P.rake_min_frames, P.rake_max_frames, P.rake_ratio = lo, hi, float(rake_ratio)
    P.onset_pad = lag + (n_fft // (2 * hop_length) if center else 0)
    out: dict = {}
    if want_sdb:
        s_db = torch.empty_like(mel)
        P.s_db, P.sdb_clip_stride, P.sdb_row_stride = s_db.data_ptr(), s_db.stride(0), s_db.stride(1)
        out["S_dB"] = s_db
    if want_rake:
        mask = torch.empty((n_clips, T), dtype=torch.uint8, device=dev)
        P.rake_mask = mask.data_ptr()
        out["rake_mask"] = mask
    if want_onset:
        # every index is written except possibly none when T < pad; start from zeros to be safe
        env = torch.zeros((n_clips, T), dtype=torch.float32, device=dev)
        mm = torch.empty((n_clips, 2), dtype=torch.float32, device=dev)
        mm[:, 0] = float("inf")
        mm[:, 1] = 0.0
        P.onset_env, P.env_minmax = env.data_ptr(), mm.data_ptr()
        out["onset_env"], out["env_minmax"] = env, mm
    nat.call("aegis_mel_post", P, _stream())
    return out


def onset_peaks(onset_env: torch.Tensor, env_minmax: torch.Tensor, *, sr: float, hop_length: int = 512,
                normalize: bool = True, **overrides) -> dict:
    """``librosa.onset.onset_detect`` peak picking; returns ``peaks`` uint8 [n_clips, T], ``n_peaks``."""
    n_clips, T = onset_env.shape
    dev = onset_env.device
    prm = tables.onset_peak_params(sr, hop_length)
    prm.update(overrides)
    P = nat.PeaksParams()
    P.onset_env, P.env_minmax, P.n_clips, P.n_frames = onset_env.data_ptr(), env_minmax.data_ptr(), n_clips, T
    P.pre_max, P.post_max, P.pre_avg, P.post_avg = int(prm["pre_max"]), int(prm["post_max"]), int(prm["pre_avg"]), int(prm["post_avg"])
    P.wait, P.normalize, P.delta = int(prm["wait"]), int(bool(normalize)), float(prm["delta"])
    cand = torch.empty((n_clips, T), dtype=torch.uint8, device=dev)
    peaks = torch.empty((n_clips, T), dtype=torch.uint8, device=dev)
    n_peaks = torch.empty((n_clips,), dtype=torch.int32, device=dev)
    P.cand, P.peaks, P.n_peaks = cand.data_ptr(), peaks.data_ptr(), n_peaks.data_ptr()
    nat.call("aegis_onset_peaks", P, _stream())
    return {"peaks": peaks, "n_peaks": n_peaks}


# ----------------------------------------------------------------------------------------------
# K2 + K3
# ----------------------------------------------------------------------------------------------
def yin_candidates(y: torch.Tensor, cfg: tables.PyinConfig, *, center: bool = True,
                   max_cand: Optional[int] = None, pad: Optional[int] = None, n_frames: Optional[int] = None,
                   want_cmnd: bool = False, split: bool = True, out: Optional[dict] = None) -> dict:
    """Sparse pYIN observations per frame (bins ascending, unique) + voiced probability.  ``split=False`` withholds the
    block-sum workspace, so hop 512 runs as one fused kernel instead of two (identical results; for tests).  ``out`` may
    hold preallocated ``cand_bin`` / ``cand_prob`` / ``cand_count`` / ``voiced_prob`` tensors for these clips (slices of
    batch-wide buffers: ``batch.TranscribePipeline`` fills them piece by piece and decodes whole groups)."""
    _check_fft(cfg.frame_length, cfg.hop_length)
    y = _check_audio(y)
    dev = y.device
    n_clips, n_samples = y.shape
    T = frame_count(n_samples, cfg.hop_length, center) if n_frames is None else int(n_frames)
    if max_cand is None:
        max_cand = cfg.max_troughs
    n_fr = n_clips * T
    key = ("pyin",) + cfg.cache_key   # every table-determining parameter (sr/hop/fmin/fmax alone alias configs)
    P = nat.YinParams()
    P.y, P.clip_stride, P.n_samples, P.n_clips = _audio_ptr(y), y.stride(0), n_samples, n_clips
    P.hop, P.pad, P.n_frames = cfg.hop_length, ((cfg.frame_length // 2 if center else 0) if pad is None else int(pad)), T
    P.twiddle = _dev_tensor("twiddle", dev, tables.fft_twiddles).data_ptr()
    P.sr, P.fmin = cfg.sr, cfg.fmin
    P.min_period, P.max_period = cfg.min_period, cfg.max_period
    P.n_pitch_bins, P.bins_per_semitone = cfg.n_pitch_bins, cfg.bins_per_semitone
    P.n_thresholds, P.max_cand = cfg.n_thresholds, max_cand
    P.thresholds = _dev_tensor(key + ("th",), dev, lambda: cfg.thresholds).data_ptr()
    P.beta_probs = _dev_tensor(key + ("bp",), dev, lambda: cfg.beta_probs).data_ptr()
    P.beta_cumsum = _dev_tensor(key + ("bc",), dev, lambda: cfg.beta_cumsum).data_ptr()
    P.boltz_fact = _dev_tensor(key + ("bf",), dev, lambda: cfg.boltz_fact).data_ptr()
    P.boltz_exp = _dev_tensor(key + ("be",), dev, lambda: cfg.boltz_exp).data_ptr()
    P.no_trough_prob = cfg.no_trough_prob
    if out is not None:
        cand_bin, cand_prob, cand_count, voiced_prob = out["cand_bin"], out["cand_prob"], out["cand_count"], out["voiced_prob"].view(-1)
        ok = (cand_bin.shape == (n_fr, max_cand) and cand_prob.shape == (n_fr, max_cand) and cand_count.shape == (n_fr,)
              and voiced_prob.shape == (n_fr,) and cand_bin.dtype == torch.int16 and cand_prob.dtype == torch.float64
              and cand_count.dtype == torch.int32 and voiced_prob.dtype == torch.float64
              and all(t.is_contiguous() and t.device == dev for t in (cand_bin, cand_prob, cand_count, voiced_prob)))
        if not ok:
            raise ValueError("yin_candidates(out=...): buffers must be contiguous CUDA tensors of this call's shapes and dtypes")
    else:
        cand_bin = torch.empty((n_fr, max_cand), dtype=torch.int16, device=dev)
        cand_prob = torch.empty((n_fr, max_cand), dtype=torch.float64, device=dev)
        cand_count = torch.empty((n_fr,), dtype=torch.int32, device=dev)
        voiced_prob = torch.empty((n_fr,), dtype=torch.float64, device=dev)
    overflow = torch.zeros((1,), dtype=torch.int32, device=dev)
    P.cand_bin, P.cand_prob, P.cand_count = cand_bin.data_ptr(), cand_prob.data_ptr(), cand_count.data_ptr()
    P.voiced_prob, P.overflow = voiced_prob.data_ptr(), overflow.data_ptr()
    work = None
    if split and cfg.hop_length == 512 and n_fr > 0:   # block-sum workspace: the two-kernel form of K2
        nbytes = int(nat.load().aegis_yin_workspace_bytes(n_clips, T, cfg.max_period))
        work = torch.empty(((nbytes + 3) // 4,), dtype=torch.float32, device=dev)
        P.block_sums = work.data_ptr()
    cmnd = None
    if want_cmnd:
        cmnd = torch.empty((n_fr, cfg.n_lags), dtype=torch.float64, device=dev)
        P.cmnd_out = cmnd.data_ptr()
    nat.call("aegis_yin_candidates", P, _stream())
    out = dict(cand_bin=cand_bin, cand_prob=cand_prob, cand_count=cand_count,
               voiced_prob=voiced_prob.view(n_clips, T), overflow=overflow, n_frames=T, max_cand=max_cand)
    if cmnd is not None:
        out["cmnd"] = cmnd.view(n_clips, T, cfg.n_lags)
    return out


def viterbi_decode(obs: dict, cfg: tables.PyinConfig, n_clips: int, *, fill_na: Optional[float] = float("nan")) -> dict:
    """Decode the pitch/voicing HMM for each clip from sparse observations (see ``yin_candidates``)."""
    T = obs["n_frames"]
    dev = obs["cand_bin"].device
    nb = cfg.n_pitch_bins
    key = ("hmm",) + cfg.cache_key
    P = nat.ViterbiParams()
    P.n_clips, P.n_frames, P.n_pitch_bins, P.half_width = n_clips, T, nb, cfg.half_width
    P.n_variants, P.n_interior_variants, P.max_cand = cfg.lt_variants.shape[0], cfg.n_interior_variants, obs["max_cand"]
    P.cand_bin, P.cand_prob = obs["cand_bin"].data_ptr(), obs["cand_prob"].data_ptr()
    P.cand_count, P.voiced_prob = obs["cand_count"].data_ptr(), obs["voiced_prob"].data_ptr()
    P.lt_variants = _dev_tensor(key + ("lt",), dev, lambda: cfg.lt_variants).data_ptr()
    P.row_variant = _dev_tensor(key + ("rv",), dev, lambda: cfg.row_variant).data_ptr()
    P.freqs = _dev_tensor(key + ("fr",), dev, lambda: cfg.freqs).data_ptr()
    P.log_tiny, P.log_init_unvoiced = cfg.log_tiny, cfg.log_init_unvoiced
    P.fill_value = float("nan") if fill_na is None else float(fill_na)
    backptr = torch.empty((n_clips, max(T, 1), 2 * nb), dtype=torch.int16, device=dev)
    final_value = torch.empty((n_clips, 2 * nb), dtype=torch.float64, device=dev)
    states = torch.empty((n_clips, T), dtype=torch.int16, device=dev)
    f0 = torch.empty((n_clips, T), dtype=torch.float64, device=dev)
    voiced = torch.empty((n_clips, T), dtype=torch.uint8, device=dev)
    P.backptr, P.final_value = backptr.data_ptr(), final_value.data_ptr()
    P.states, P.f0, P.voiced_flag = states.data_ptr(), f0.data_ptr(), voiced.data_ptr()
    nat.call("aegis_viterbi", P, _stream())
    if fill_na is None:  # librosa: keep the best-guess pitch on unvoiced frames
        freqs = _dev_tensor(key + ("fr",), dev, lambda: cfg.freqs)
        f0 = freqs[(states.to(torch.int64) & 0xFFFF) % nb]
    return dict(states=states, f0=f0, voiced_flag=voiced)


def pyin_batch(y: torch.Tensor, *, sr: float, fmin: float, fmax: float, hop_length: int = 512,
               frame_length: int = N_FFT, center: bool = True, fill_na: Optional[float] = float("nan"),
               clips_per_launch: Optional[int] = None, **hmm_kwargs) -> dict:
    """Batched ``librosa.pyin``: f0 [n_clips, T] float64 (NaN unvoiced), voiced_flag, voiced_prob."""
    cfg = tables.pyin_config(float(sr), int(hop_length), float(fmin), float(fmax), int(frame_length), **hmm_kwargs)
    y = _check_audio(y)
    n_clips = y.shape[0]
    if clips_per_launch is None:
        clips_per_launch = n_clips
    parts = []
    for c0 in range(0, n_clips, clips_per_launch):
        yc = y[c0 : c0 + clips_per_launch]
        obs = yin_candidates(yc, cfg, center=center)
        dec = viterbi_decode(obs, cfg, yc.shape[0], fill_na=fill_na)
        dec["voiced_prob"] = obs["voiced_prob"]
        parts.append(dec)
    if len(parts) == 1:
        return parts[0]
    return {k: torch.cat([p[k] for p in parts]) for k in parts[0]}


# ----------------------------------------------------------------------------------------------
# K5
# ----------------------------------------------------------------------------------------------
TREND_OUTPUTS = ("savgol", "kalman", "holt", "consensus", "consensus_conf", "sma", "ema",
                 "boll_ma", "boll_upper", "boll_lower", "macd_line", "macd_sig", "macd_hist")


def trend_filters(x: torch.Tensor, *, want=TREND_OUTPUTS, savgol_window: int = 11, savgol_polyorder: int = 3,
                  kalman_q: float = 1e-5, kalman_r: float = 1e-1, holt_alpha: float = 0.3, holt_beta: float = 0.1,
                  sma_window: int = 5, ema_span: int = 5, boll_window: int = 20, boll_num_std: float = 2.0,
                  macd_fast: int = 12, macd_slow: int = 26, macd_signal: int = 9,
                  consensus_filters=("savgol", "kalman", "holt")) -> dict:
    """Batched float64 trend filters over f0 series [n_series, n] (NaN = unvoiced).  ``consensus_filters``: the filters
    that vote in ``consensus`` / ``consensus_conf`` (``multi_filter_consensus(data, filters=[...])``)."""
    if not x.is_cuda:
        raise nat.AegisNativeError("expected a CUDA tensor: the Aegis B200 path has no CPU implementation")
    if x.dim() == 1:
        x = x[None]
    x = x.to(torch.float64).contiguous()
    n_series, n = x.shape
    dev = x.device
    P = nat.TrendParams()
    P.x, P.n_series, P.n = x.data_ptr(), n_series, n
    coeffs = _dev_tensor(("savgol", savgol_window, savgol_polyorder), dev,
                         lambda: tables.savgol_coeffs(savgol_window, savgol_polyorder))
    P.savgol_coeffs, P.savgol_window = coeffs.data_ptr(), savgol_window
    P.sma_window, P.ema_span, P.boll_window = sma_window, ema_span, boll_window
    P.macd_fast, P.macd_slow, P.macd_signal = macd_fast, macd_slow, macd_signal
    P.kalman_q, P.kalman_r, P.holt_alpha, P.holt_beta, P.boll_num_std = kalman_q, kalman_r, holt_alpha, holt_beta, boll_num_std
    compact = torch.empty_like(x)
    scratch = torch.empty_like(x)
    P.compact, P.scratch = compact.data_ptr(), scratch.data_ptr()
    out = {}
    need = set(want)
    if need & {"consensus", "consensus_conf"}:
        votes = [f for f in ("savgol", "kalman", "holt") if f in consensus_filters]
        if not votes:
            raise ValueError("consensus needs at least one of savgol / kalman / holt")
        need |= set(votes)
        P.consensus_mask = sum(1 << ("savgol", "kalman", "holt").index(f) for f in votes)
    if need & {"boll_upper", "boll_lower"}:
        need |= {"boll_ma"}
    if need & {"macd_sig", "macd_hist"}:
        need |= {"macd_line"}
    for name in TREND_OUTPUTS:
        if name in need:
            t = torch.empty_like(x)
            setattr(P, name, t.data_ptr())
            out[name] = t
    nat.call("aegis_trend_filters", P, _stream())
    return {k: v for k, v in out.items() if k in want}


# ----------------------------------------------------------------------------------------------
# K6
# ----------------------------------------------------------------------------------------------
DISTORTION_LABELS = ("clean", "light", "heavy")


def guitar_frame_limits(hop_length: int, sr: float, duration_ms: float = 50):
    """(mute_max_frames, rake_frames) of guitar_specific.py:97-98,137-138."""
    ms_per_frame = (hop_length / sr) * 1000
    return int(duration_ms / ms_per_frame), int(30 / ms_per_frame)


def guitar_filters(s_db: Optional[torch.Tensor], *, sr: float, hop_length: int = 512, f0: Optional[torch.Tensor] = None,
                   voiced_flag: Optional[torch.Tensor] = None, rake_mask: Optional[torch.Tensor] = None,
                   fmin_hz: float = 82.4, duration_ms: float = 50, want_rake: bool = True, want_mute: bool = True,
                   want_distortion: bool = True) -> dict:
    """``apply_guitar_filters`` for a batch (aegis_engine_core_v2/guitar_specific.py:240-277).

    ``s_db`` f32 [n_clips, n_mels, T] dB images; ``f0`` f64 [n_clips, T] (NaN unvoiced) with ``voiced_flag`` u8/bool;
    ``rake_mask`` u8/bool [n_clips, T] basic mask.  Returns the requested keys among ``f0``, ``voiced``, ``rake_mask``,
    ``mute_mask`` (u8), ``distortion`` (int32 codes into ``DISTORTION_LABELS``).
    """
    P = nat.GuitarParams()
    out: dict = {}
    dev = None
    n_clips = T = 0
    if s_db is not None:
        if s_db.dtype != torch.float32 or s_db.stride(2) != 1:
            s_db = s_db.float().contiguous()
        n_clips, n_mels, T = s_db.shape
        dev = s_db.device
        P.s_db, P.sdb_clip_stride, P.sdb_row_stride, P.n_mels = s_db.data_ptr() if s_db.numel() else None, s_db.stride(0), s_db.stride(1), n_mels
        P.mute_max_frames, P.rake_frames = guitar_frame_limits(hop_length, sr, duration_ms)
        if want_rake:
            if rake_mask is not None:
                rake_mask = rake_mask.to(torch.uint8).contiguous()
                P.rake_in = rake_mask.data_ptr()
            out["rake_mask"] = torch.empty((n_clips, T), dtype=torch.uint8, device=dev)
            P.rake_out = out["rake_mask"].data_ptr()
        if want_mute:
            out["mute_mask"] = torch.empty((n_clips, T), dtype=torch.uint8, device=dev)
            P.mute_out = out["mute_mask"].data_ptr()
        if want_distortion:
            nb = int(nat.load().aegis_guitar_blocks(T))
            out["distortion"] = torch.zeros((n_clips,), dtype=torch.int32, device=dev)
            work = torch.empty((max(1, n_clips * nb * 2),), dtype=torch.float64, device=dev)
            P.distortion, P.dist_work = out["distortion"].data_ptr(), work.data_ptr()
    if f0 is not None:
        f0 = f0.to(torch.float64).contiguous()
        if s_db is not None and tuple(f0.shape) != (n_clips, T):
            raise ValueError("f0 must be [n_clips, T] like the dB images")
        n_clips, T = f0.shape
        dev = f0.device
        P.f0 = f0.data_ptr()
        if voiced_flag is not None:
            voiced_flag = voiced_flag.to(torch.uint8).contiguous()
            P.voiced = voiced_flag.data_ptr()
        out["f0"] = torch.empty_like(f0)
        out["voiced"] = torch.empty((n_clips, T), dtype=torch.uint8, device=dev)
        P.f0_out, P.voiced_out = out["f0"].data_ptr(), out["voiced"].data_ptr()
    if dev is None:
        raise ValueError("guitar_filters needs s_db and / or f0")
    P.n_clips, P.n_frames, P.fmin_hz = n_clips, T, float(fmin_hz)
    nat.call("aegis_guitar_filters", P, _stream())
    return out


# ----------------------------------------------------------------------------------------------
# K7
# ----------------------------------------------------------------------------------------------
NOTE_EVENT_DTYPE = np.dtype([("note", "<i4"), ("start", "<i4"), ("end", "<i4"), ("velocity", "<i4"), ("rms_energy", "<f4"),
                             ("track", "u1"), ("technique", "u1"), ("_pad", "u1", (2,)), ("confidence", "<f8"), ("slope", "<f8")])
TECHNIQUES = (None, "vibrato", "bend", "slide", "hammer_on", "pull_off")
assert NOTE_EVENT_DTYPE.itemsize == 40


def note_frame_limits(sr: float, hop_length: int, sustain_ms: float = 50, min_note_duration_ms: float = 50):
    """(min_note_duration_frames, sustain_frames) of midi_logic.py:56-57."""
    return int((min_note_duration_ms / 1000.0) * sr / hop_length), int((sustain_ms / 1000.0) * sr / hop_length)


def note_events(rake_mask: torch.Tensor, f0: torch.Tensor, voiced_flag: torch.Tensor, voiced_prob: torch.Tensor,
                rms: torch.Tensor, *, sr: float, hop_length: int = 512, confidence_threshold: float = 0.7,
                noise_gate_db: float = -40, sustain_ms: float = 50, min_note_duration_ms: float = 50,
                pitch_index: Optional[torch.Tensor] = None, note_lut: Optional[torch.Tensor] = None,
                max_events: Optional[int] = None) -> dict:
    """``get_midi_events`` for a batch (aegis_engine_core/midi_logic.py:32-148): frame arrays [n_clips, T] in,
    ``events`` (uint8 [n_clips, max_events, 40], records of ``NOTE_EVENT_DTYPE``) and ``n_events`` int32 [n_clips] out.

    ``f0`` is the v1 pitch track (0 where unvoiced, aegis_engine.py:69).  With ``pitch_index`` (uint16 index of every
    frame's f0 value, stored in an int16 tensor, e.g. the Viterbi states) and ``note_lut`` (int16 MIDI note per index, computed on the host with the
    reference's numpy expression) note numbers are exact even at quarter-tone ties.
    """
    f0 = f0.to(torch.float64).contiguous()
    n_clips, T = f0.shape
    dev = f0.device
    rake_mask = rake_mask.to(torch.uint8).contiguous()
    voiced_flag = voiced_flag.to(torch.uint8).contiguous()
    voiced_prob = voiced_prob.to(torch.float64).contiguous()
    if rms.dtype != torch.float32 or rms.stride(-1) != 1:
        rms = rms.float().contiguous()
    for name, t in (("rake_mask", rake_mask), ("voiced_flag", voiced_flag), ("voiced_prob", voiced_prob), ("rms", rms)):
        if tuple(t.shape) != (n_clips, T):
            raise ValueError(f"{name} must be [n_clips, T] like f0")
    min_frames, sustain_frames = note_frame_limits(sr, hop_length, sustain_ms, min_note_duration_ms)
    if max_events is None:   # an event that survives the duration filter spans min_frames + 1 frames
        max_events = T // (min_frames + 1) + 1
    P = nat.NotesParams()
    P.rake_mask, P.f0, P.voiced_flag, P.voiced_prob = rake_mask.data_ptr(), f0.data_ptr(), voiced_flag.data_ptr(), voiced_prob.data_ptr()
    P.rms, P.rms_clip_stride = rms.data_ptr(), rms.stride(0)
    if (pitch_index is None) != (note_lut is None):
        raise ValueError("pitch_index and note_lut go together")
    if pitch_index is not None:
        if pitch_index.dtype != torch.int16:   # the uint16 bit pattern travels in an int16 tensor (as the Viterbi states do)
            raise ValueError("pitch_index must be an int16 tensor holding uint16 bit patterns")
        pitch_index = pitch_index.contiguous()
        note_lut = note_lut.to(torch.int16).contiguous()
        P.pitch_index, P.note_lut, P.n_lut = pitch_index.data_ptr(), note_lut.data_ptr(), note_lut.numel()
    P.n_clips, P.n_frames, P.hop, P.sr = n_clips, T, hop_length, float(sr)
    P.confidence_threshold, P.noise_gate_db = float(confidence_threshold), float(noise_gate_db)
    P.min_note_frames, P.sustain_frames, P.max_events = min_frames, sustain_frames, max_events
    n_bytes = int(nat.load().aegis_note_events_bytes(n_clips, T, max_events))
    buf = torch.empty((n_bytes,), dtype=torch.uint8, device=dev)
    n_events = torch.zeros((n_clips,), dtype=torch.int32, device=dev)
    P.events, P.n_events = buf.data_ptr(), n_events.data_ptr()
    nat.call("aegis_note_events", P, _stream())
    events = buf[: n_clips * max_events * NOTE_EVENT_DTYPE.itemsize].view(n_clips, max_events, NOTE_EVENT_DTYPE.itemsize)
    return {"events": events, "n_events": n_events, "max_events": max_events}


def note_events_to_list(events: torch.Tensor, n_events: torch.Tensor, clip: int) -> list:
    """One clip's records as the reference's list of dicts (keys and value types of midi_logic.py:71-104)."""
    n = int(n_events[clip])
    if n > events.shape[1]:
        raise nat.AegisNativeError(f"clip {clip}: {n} note events exceed the buffer of {events.shape[1]}")
    rec = events[clip, :n].cpu().numpy().reshape(-1).view(NOTE_EVENT_DTYPE) if n else np.zeros(0, NOTE_EVENT_DTYPE)
    out = []
    for r in rec:
        out.append({"note": int(r["note"]), "start": int(r["start"]), "end": int(r["end"]), "confidence": r["confidence"],
                    "velocity": int(r["velocity"]), "track": "main" if r["track"] else "safe", "rms_energy": r["rms_energy"],
                    "technique": TECHNIQUES[int(r["technique"])], "slope": float(r["slope"])})
    return out


# ----------------------------------------------------------------------------------------------
# K8
# ----------------------------------------------------------------------------------------------
FIN_EVENT_DTYPE = np.dtype([("note", "<i4"), ("start", "<i4"), ("end", "<i4"), ("velocity", "<i4"), ("track", "u1"),
                            ("technique", "u1"), ("slide", "u1"), ("harmonic_valid", "i1"), ("_pad", "u1", (4,)),
                            ("confidence", "<f8")])
FIN_ARTICULATIONS = (None, "normal", "bend", "vibrato", "noise")        # detect_articulation_bollinger's labels
FIN_SLIDES = (None, "normal", "slide_up", "slide_down")                 # detect_slides_macd's labels
KEY_NAMES = ("C", "C#", "D", "D#", "E", "F", "F#", "G", "G#", "A", "A#", "B")
KEY_MODES = ("major", "minor", "blues")


def note_events_financial(rake_mask: torch.Tensor, f0: torch.Tensor, voiced_flag: torch.Tensor, voiced_prob: torch.Tensor,
                          rms: torch.Tensor, *, sr: float, hop_length: int = 512, confidence_threshold: Optional[float] = None,
                          noise_gate_db: float = -40, sustain_ms: float = 50, min_note_duration_ms: float = 50,
                          use_harmonic_filter: bool = True, harmonic_tolerance: int = 1, slide_threshold: float = 0.3,
                          rsi_threshold: float = 70, max_events: Optional[int] = None) -> dict:
    """``get_midi_events_financial(use_financial=True)`` for a batch
    (aegis_engine_core_v2/midi_logic_financial.py:117-388): frame arrays [n_clips, T] in (``f0`` in Hz with NaN where
    unvoiced, as pYIN returns it), ``events`` (uint8 [n_clips, max_events, 32], records of ``FIN_EVENT_DTYPE``),
    ``n_events`` int32, ``threshold`` f64, ``key`` int32 (root | mode << 8, -1 when no note was out of scale) and
    ``key_confidence`` f64 out, all [n_clips].  Launches K8 prepare, K5 twice (f0 series, semitone series), K8 events.
    """
    if not f0.is_cuda:
        raise nat.AegisNativeError("expected CUDA tensors: the Aegis B200 path has no CPU implementation")
    f0 = f0.to(torch.float64).contiguous()
    n_clips, T = f0.shape
    dev = f0.device
    rake_mask = rake_mask.to(torch.uint8).contiguous()
    voiced_flag = voiced_flag.to(torch.uint8).contiguous()
    voiced_prob = voiced_prob.to(torch.float64).contiguous()
    if rms.dtype != torch.float32 or rms.stride(-1) != 1:
        rms = rms.float().contiguous()
    for name, t in (("rake_mask", rake_mask), ("voiced_flag", voiced_flag), ("voiced_prob", voiced_prob), ("rms", rms)):
        if tuple(t.shape) != (n_clips, T):
            raise ValueError(f"{name} must be [n_clips, T] like f0")
    min_frames, sustain_frames = note_frame_limits(sr, hop_length, sustain_ms, min_note_duration_ms)
    if max_events is None:   # an event that survives the duration filter spans min_frames + 1 frames
        max_events = T // (min_frames + 1) + 1
    P = nat.FinParams()
    P.rake_mask, P.f0, P.voiced_flag, P.voiced_prob = rake_mask.data_ptr(), f0.data_ptr(), voiced_flag.data_ptr(), voiced_prob.data_ptr()
    P.rms, P.rms_clip_stride = rms.data_ptr(), rms.stride(0)
    P.n_clips, P.n_frames, P.hop, P.sr = n_clips, T, hop_length, float(sr)
    f0_clean = torch.empty_like(f0)
    semitones = torch.empty_like(f0)
    scratch = torch.empty((int(nat.load().aegis_fin_scratch_bytes(n_clips, T)),), dtype=torch.uint8, device=dev)
    P.f0_clean, P.semitones, P.scratch = f0_clean.data_ptr(), semitones.data_ptr(), scratch.data_ptr()
    if T > 0 and n_clips > 0:
        nat.call("aegis_fin_prepare", P, _stream())
    events = torch.zeros((n_clips, max_events, FIN_EVENT_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    n_events = torch.zeros((n_clips,), dtype=torch.int32, device=dev)
    threshold = torch.empty((n_clips,), dtype=torch.float64, device=dev)
    key = torch.empty((n_clips,), dtype=torch.int32, device=dev)
    key_conf = torch.empty((n_clips,), dtype=torch.float64, device=dev)
    if T > 0 and n_clips > 0:
        a = trend_filters(f0_clean, want=("consensus", "boll_upper", "boll_lower"), boll_window=10, boll_num_std=2.0)
        m = trend_filters(semitones, want=("macd_line", "macd_hist"), macd_fast=5, macd_slow=20, macd_signal=9)
        P.trend, P.boll_upper, P.boll_lower = a["consensus"].data_ptr(), a["boll_upper"].data_ptr(), a["boll_lower"].data_ptr()
        P.macd_line, P.macd_hist = m["macd_line"].data_ptr(), m["macd_hist"].data_ptr()
        P.confidence_threshold = float("nan") if confidence_threshold is None else float(confidence_threshold)
        P.slide_threshold, P.rsi_threshold, P.noise_gate_db = float(slide_threshold), float(rsi_threshold), float(noise_gate_db)
        P.min_note_frames, P.sustain_frames, P.max_events = min_frames, sustain_frames, max_events
        P.use_harmonic_filter, P.harmonic_tolerance = int(bool(use_harmonic_filter)), int(harmonic_tolerance)
        P.events, P.n_events = events.data_ptr(), n_events.data_ptr()
        P.threshold_out, P.key_out, P.key_confidence_out = threshold.data_ptr(), key.data_ptr(), key_conf.data_ptr()
        nat.call("aegis_fin_events", P, _stream())
    else:
        threshold.fill_(0.5 if confidence_threshold is None else float(confidence_threshold))
        key.fill_(-1)
        key_conf.zero_()
    return {"events": events, "n_events": n_events, "max_events": max_events, "threshold": threshold, "key": key,
            "key_confidence": key_conf}


def fin_events_to_list(res: dict, clip: int) -> list:
    """One clip's records as the reference's list of dicts (keys of midi_logic_financial.py:222-231,361-384)."""
    events, n = res["events"], int(res["n_events"][clip])
    if n > events.shape[1]:
        raise nat.AegisNativeError(f"clip {clip}: {n} note events exceed the buffer of {events.shape[1]}")
    key = int(res["key"][clip])
    if key >= 0 and n == 0:
        raise IndexError("list index out of range")   # the reference indexes events[0] after removing every note
    rec = events[clip, :n].cpu().numpy().reshape(-1).view(FIN_EVENT_DTYPE) if n else np.zeros(0, FIN_EVENT_DTYPE)
    out = []
    for r in rec:
        art = FIN_ARTICULATIONS[int(r["technique"])]
        e = {"note": int(r["note"]), "start": int(r["start"]), "end": int(r["end"]), "confidence": float(r["confidence"]),
             "velocity": int(r["velocity"]), "track": "main" if r["track"] else "safe", "financial_artic": art,
             "financial_slide": FIN_SLIDES[int(r["slide"])], "technique": art}
        if r["harmonic_valid"] >= 0:
            e["harmonic_valid"] = bool(r["harmonic_valid"])
        out.append(e)
    if key >= 0 and out:
        out[0]["key_info"] = {"key": KEY_NAMES[key & 0xFF], "mode": KEY_MODES[key >> 8],
                              "confidence": float(res["key_confidence"][clip])}
    return out


# ----------------------------------------------------------------------------------------------
# K9
# ----------------------------------------------------------------------------------------------
def resample_poly(x: torch.Tensor, orig_sr: int, target_sr: int, *, n_channels: int = 1) -> torch.Tensor:
    """``librosa.resample(y, orig_sr=, target_sr=, res_type='polyphase')`` for a batch, fused with the PCM -> float
    conversion and ``librosa.to_mono``: ``x`` is [n_clips, n_frames * n_channels] float32 or int16 (channels
    interleaved, int16 scaled by 1/32768 as soundfile does); returns float32 [n_clips, ceil(n_frames * target / orig)],
    bit-identical to ``scipy.signal.resample_poly`` on the float32 mono signal.  Equal rates only convert and mix."""
    if not x.is_cuda:
        raise nat.AegisNativeError("expected a CUDA tensor: the Aegis B200 path has no CPU implementation")
    if int(orig_sr) != orig_sr or int(target_sr) != target_sr or orig_sr <= 0 or target_sr <= 0:
        raise ValueError("polyphase resampling needs positive integer sample rates")   # as librosa.resample does
    if x.dtype not in (torch.float32, torch.int16):
        raise ValueError("x must be float32 or int16")
    if x.dim() == 1:
        x = x[None]
    x = x.contiguous()
    n_clips, width = x.shape
    if n_channels < 1 or width % n_channels:
        raise ValueError("row length must be a multiple of n_channels")
    n_in = width // n_channels
    g = int(np.gcd(int(orig_sr), int(target_sr)))
    up, down = int(target_sr) // g, int(orig_sr) // g
    n_out = -(-n_in * up // down)
    out = torch.empty((n_clips, n_out), dtype=torch.float32, device=x.device)
    if n_clips == 0 or n_out == 0:
        return out
    if up == down == 1:     # identity filter: one tap of weight 1 at the output sample
        taps, n_pre_pad, n_pre_remove = np.ones(1, np.float32), 0, 0
    else:
        taps, n_pre_pad, n_pre_remove = tables.resample_poly_design(up, down)
    taps_d = _dev_tensor(("resample", up, down), x.device, lambda: taps)
    P = nat.ResampleParams()
    P.x, P.in_clip_stride, P.n_in = x.data_ptr(), x.stride(0), n_in
    P.in_format, P.n_channels, P.n_clips = (1 if x.dtype == torch.int16 else 0), n_channels, n_clips
    P.up, P.down, P.n_taps, P.taps = up, down, int(taps_d.numel()), taps_d.data_ptr()
    P.n_pre_pad, P.n_pre_remove = n_pre_pad, n_pre_remove
    P.out, P.out_clip_stride, P.n_out = out.data_ptr(), out.stride(0), n_out
    nat.call("aegis_resample_poly", P, _stream())
    return out


# ----------------------------------------------------------------------------------------------
# corpus synthesis
# ----------------------------------------------------------------------------------------------
def synth_events(n_clips: int, n_samples: int, events: dict, device, decay: float = 0.996) -> torch.Tensor:
    """Render Karplus-Strong / rake events (see corpus.plan_events) into [n_clips, n_samples] float32."""
    out = torch.zeros((n_clips, n_samples), dtype=torch.float32, device=device)
    ev = {k: torch.from_numpy(np.ascontiguousarray(v).view(np.int32) if v.dtype == np.uint32 else np.ascontiguousarray(v)).to(device)
          for k, v in events.items()}
    P = nat.SynthParams()
    P.out, P.clip_stride, P.n_events = out.data_ptr(), out.stride(0), int(ev["clip"].numel())
    P.ev_clip, P.ev_start, P.ev_len = ev["clip"].data_ptr(), ev["start"].data_ptr(), ev["length"].data_ptr()
    P.ev_period, P.ev_amp, P.ev_seed = ev["period"].data_ptr(), ev["amp"].data_ptr(), ev["seed"].data_ptr()
    P.decay = decay
    if P.n_events:
        nat.call("aegis_synth_ks", P, _stream())
    peak = out.abs().amax(dim=1, keepdim=True).clamp_min(1e-20)
    return out / peak * 0.9
