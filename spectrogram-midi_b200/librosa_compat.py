"""Drop-in replacements for the ``librosa.*`` calls on Aegis Engine's hot path.

Same names, argument meaning, dtypes, shapes and error behaviour as the librosa functions the
reference calls; numpy in, numpy out, computed by the sm_100a kernels behind
``libaegis_b200.so``.  Usage on the reference side::

    import spectrogram_midi_b200.librosa_compat as librosa      # instead of `import librosa`

Reference call sites: ``aegis_engine.py:24-26,63,67,70,190,216``, ``aegis_engine_core/worker.py:9-15``,
``aegis_engine_financial.py:45-51,63-69,154``; consumer-side helpers used by
``aegis_engine_core/midi_logic.py:17,43,51,69`` (``hz_to_midi``, ``amplitude_to_db``,
``util.softmask``) are plain numpy, exactly as in librosa.

There is no CPU fallback for the kernels: without a CUDA device / the built library these raise.
"""
from __future__ import annotations

import types
from typing import Optional

import numpy as np
import torch

from . import core, tables
from ._native import AegisNativeError

note_to_hz = tables.note_to_hz
note_to_midi = tables.note_to_midi
midi_to_hz = tables.midi_to_hz
hz_to_midi = tables.hz_to_midi


class ParameterError(ValueError):
    """Mirror of librosa.util.exceptions.ParameterError."""


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise AegisNativeError("no CUDA device: the Aegis B200 path has no CPU implementation")
    return torch.device("cuda", torch.cuda.current_device())


def _audio_to_device(y, *, what="y") -> torch.Tensor:
    y = np.asarray(y)
    if y.ndim != 1:
        raise ParameterError(f"{what} must be mono (1-d); got shape {y.shape}")
    if not np.issubdtype(y.dtype, np.floating):
        raise ParameterError("Audio data must be floating-point")
    if y.size and not np.isfinite(y).all():
        raise ParameterError("Audio buffer is not finite everywhere")
    return torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32)).to(_device())[None]


def _prepare(y, n_fft, hop_length, center, pad_mode):
    """Returns (device audio [1, N'], center flag for the kernels)."""
    if center and pad_mode not in ("constant", "zeros"):
        # any other numpy pad mode: pad on the host exactly as librosa does, then frame uncentred
        y = np.pad(np.asarray(y, dtype=np.float32), n_fft // 2, mode=pad_mode)
        return _audio_to_device(y), False
    return _audio_to_device(y), bool(center)


# ------------------------------------------------------------------------------------------------
# spectral front end
# ------------------------------------------------------------------------------------------------
def stft_magnitude(y, *, n_fft=2048, hop_length=512, center=True, pad_mode="constant"):
    """``np.abs(librosa.stft(y, ...))``: float32 [1 + n_fft//2, T] (the phase is never formed)."""
    yd, c = _prepare(y, n_fft, hop_length, center, pad_mode)
    if yd.shape[1] < n_fft and not c:
        raise ParameterError(f"Input signal length={yd.shape[1]} is too small for n_fft={n_fft}")
    return core.stft_features(yd, hop_length=hop_length, n_fft=n_fft, center=c)["mag"][0].cpu().numpy()


def melspectrogram(*, y=None, sr=22050, S=None, n_fft=2048, hop_length=512, win_length=None, window="hann",
                   center=True, pad_mode="constant", power=2.0, n_mels=128, **kwargs):
    """``librosa.feature.melspectrogram`` (aegis_engine.py:25): float32 [n_mels, T] mel power."""
    if S is not None or power != 2.0 or window != "hann" or win_length not in (None, n_fft) or kwargs:
        raise NotImplementedError("melspectrogram: only y=, power=2.0, window='hann', default mel options")
    yd, c = _prepare(y, n_fft, hop_length, center, pad_mode)
    out = core.stft_features(yd, sr=sr, hop_length=hop_length, n_fft=n_fft, center=c, n_mels=n_mels,
                             want_mag=False, want_mel=True)
    return out["mel"][0].cpu().numpy()


def rms(*, y=None, S=None, frame_length=2048, hop_length=512, center=True, pad_mode="constant"):
    """``librosa.feature.rms`` (aegis_engine.py:70): float32 [1, T]."""
    if S is not None:
        raise NotImplementedError("rms: only the time-domain form rms(y=...) is on the hot path")
    yd, c = _prepare(y, frame_length, hop_length, center, pad_mode)
    out = core.stft_features(yd, hop_length=hop_length, n_fft=frame_length, center=c, want_mag=False, want_rms=True)
    return out["rms"].cpu().numpy()


def power_to_db(S, *, ref=1.0, amin=1e-10, top_db=80.0):
    """``librosa.power_to_db`` (aegis_engine.py:26).  ``ref`` may be a scalar or a callable (np.max)."""
    S = np.asarray(S)
    if np.iscomplexobj(S):
        S = np.abs(S)
    if amin <= 0:
        raise ParameterError("amin must be strictly positive")
    if amin != 1e-10 or top_db != 80.0 or S.ndim != 2:
        raise NotImplementedError("power_to_db: kernels are built for 2-d input, amin=1e-10, top_db=80")
    ref_value = ref(S) if callable(ref) else np.abs(ref)
    dev = _device()
    Sd = torch.from_numpy(np.ascontiguousarray(S, dtype=np.float32)).to(dev)[None]
    smax = torch.tensor([float(S.max())], dtype=torch.float32, device=dev)
    refp = torch.tensor([float(ref_value)], dtype=torch.float32, device=dev)
    out = core.mel_post(Sd, smax, sr=22050, ref_power=refp, want_sdb=True, want_rake=False)
    return out["S_dB"][0].cpu().numpy()


def amplitude_to_db(S, *, ref=1.0, amin=1e-5, top_db=80.0):
    """``librosa.amplitude_to_db`` -- consumer-side helper (midi_logic.py:51); numpy, as in librosa."""
    magnitude = np.abs(np.asarray(S))
    ref_value = ref(magnitude) if callable(ref) else np.abs(ref)
    power = np.square(magnitude, out=magnitude.copy())
    log_spec = 10.0 * np.log10(np.maximum(amin**2, power))
    log_spec -= 10.0 * np.log10(np.maximum(amin**2, ref_value**2))
    if top_db is not None:
        log_spec = np.maximum(log_spec, log_spec.max() - top_db)
    return log_spec


# ------------------------------------------------------------------------------------------------
# pYIN
# ------------------------------------------------------------------------------------------------
def pyin(y, *, fmin, fmax, sr=22050, frame_length=2048, win_length=None, hop_length=None, n_thresholds=100,
         beta_parameters=(2, 18), boltzmann_parameter=2, resolution=0.1, max_transition_rate=35.92,
         switch_prob=0.01, no_trough_prob=0.01, fill_na=np.nan, center=True, pad_mode="constant"):
    """``librosa.pyin``: returns ``(f0 float64 [T], voiced_flag bool [T], voiced_prob float64 [T])``."""
    if fmin is None or fmax is None:
        raise ParameterError('both "fmin" and "fmax" must be provided')
    if win_length not in (None, frame_length // 2):
        raise NotImplementedError("pyin: win_length must be frame_length // 2")
    if hop_length is None:
        hop_length = frame_length // 4
    yd, c = _prepare(y, frame_length, hop_length, center, pad_mode)
    if yd.shape[1] == 0 or (not c and yd.shape[1] < frame_length):
        raise ParameterError("pyin: input too short")
    out = core.pyin_batch(
        yd, sr=sr, fmin=float(fmin), fmax=float(fmax), hop_length=int(hop_length), frame_length=int(frame_length),
        center=c, fill_na=fill_na, n_thresholds=int(n_thresholds), beta_parameters=tuple(beta_parameters),
        boltzmann_parameter=boltzmann_parameter, resolution=resolution, max_transition_rate=max_transition_rate,
        switch_prob=switch_prob, no_trough_prob=no_trough_prob)
    return (out["f0"][0].cpu().numpy(), out["voiced_flag"][0].cpu().numpy().astype(bool),
            out["voiced_prob"][0].cpu().numpy())


# ------------------------------------------------------------------------------------------------
# onsets (no reference call site; BASELINE north_star)
# ------------------------------------------------------------------------------------------------
def _onset_device(y, sr, hop_length, n_fft=2048, center=True):
    yd, c = _prepare(y, n_fft, hop_length, center, "constant")
    feat = core.stft_features(yd, sr=sr, hop_length=hop_length, center=c, want_mag=False, want_mel=True)
    return core.mel_post(feat["mel"], feat["mel_max"], sr=sr, hop_length=hop_length, want_sdb=False,
                         want_rake=False, want_onset=True, center=c)


def onset_strength(*, y=None, sr=22050, hop_length=512, n_fft=2048, center=True, **kwargs):
    if kwargs:
        raise NotImplementedError(f"onset_strength: unsupported options {sorted(kwargs)}")
    return _onset_device(y, sr, hop_length, n_fft, center)["onset_env"][0].cpu().numpy()


def onset_detect(*, y=None, sr=22050, onset_envelope=None, hop_length=512, units="frames", normalize=True, **kwargs):
    if units not in ("frames", "samples", "time"):
        raise ParameterError(f"Invalid unit type: {units}")
    if onset_envelope is None:
        post = _onset_device(y, sr, hop_length)
        env, mm = post["onset_env"], post["env_minmax"]
    else:
        e = np.ascontiguousarray(onset_envelope, dtype=np.float32)
        env = torch.from_numpy(e).to(_device())[None]
        mm = torch.tensor([[float(e.min()) if e.size else 0.0, float(e.max()) if e.size else 0.0]],
                          dtype=torch.float32, device=env.device)
    if env.shape[1] == 0:
        return np.array([], dtype=np.int64)
    pk = core.onset_peaks(env, mm, sr=sr, hop_length=hop_length, normalize=normalize, **kwargs)
    frames = np.flatnonzero(pk["peaks"][0].cpu().numpy())
    if units == "samples":
        return frames * hop_length
    if units == "time":
        return frames * hop_length / float(sr)
    return frames


# ------------------------------------------------------------------------------------------------
# librosa-shaped namespaces, so `librosa.feature.rms(...)` etc. keep working
# ------------------------------------------------------------------------------------------------
def _softmask(X, X_ref, *, power=1, split_zeros=False):
    """``librosa.util.softmask`` (numpy).  Like librosa's it has NO ``margin`` keyword, so the call at
    midi_logic.py:43 raises TypeError and the reference takes its raw-f0 branch (:47-49)."""
    X, X_ref = np.asarray(X, dtype=np.float64), np.asarray(X_ref, dtype=np.float64)
    if X.shape != X_ref.shape:
        raise ParameterError(f"Shape mismatch: {X.shape}!={X_ref.shape}")
    if np.any(X < 0) or np.any(X_ref < 0):
        raise ParameterError("X and X_ref must be non-negative")
    if power <= 0:
        raise ParameterError("power must be strictly positive")
    Z = np.maximum(X, X_ref)
    bad = Z < np.finfo(np.float64).tiny
    Z[bad] = 1
    mask = (X / Z) ** power
    ref_mask = (X_ref / Z) ** power
    good = ~bad
    mask[good] /= mask[good] + ref_mask[good]
    mask[bad] = 0.5 if split_zeros else 0.0
    return mask


feature = types.SimpleNamespace(melspectrogram=melspectrogram, rms=rms)
onset = types.SimpleNamespace(onset_strength=onset_strength, onset_detect=onset_detect)
util = types.SimpleNamespace(softmask=_softmask)


def resample(y, *, orig_sr, target_sr, res_type="soxr_hq", fix=True, scale=False, axis=-1, **kwargs):
    """``librosa.resample`` for ``res_type='polyphase'`` (= ``scipy.signal.resample_poly``) on the GPU (kernel K9),
    bit-identical to it for float32 input.  The default is librosa's (``'soxr_hq'``, libsoxr), which this image does not
    have and whose arithmetic cannot be pinned: it and every other ``res_type`` raise, so a caller must ask for
    ``'polyphase'`` by name -- nothing is substituted silently."""
    if res_type != "polyphase":
        raise NotImplementedError(f"res_type={res_type!r}: only 'polyphase' (scipy.signal.resample_poly) is built on the B200 path; "
                                  "pass res_type='polyphase' explicitly")
    y = np.asarray(y)
    if y.ndim != 1 or axis not in (-1, 0):
        raise NotImplementedError("mono signals only")
    if orig_sr == target_sr:
        return y
    ratio = float(target_sr) / orig_sr
    n_samples = int(np.ceil(y.shape[-1] * ratio))
    if n_samples < 1:
        raise ValueError(f"Input signal length={y.shape[-1]} is too small to resample from {orig_sr}->{target_sr}")
    out = core.resample_poly(torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32)).to(_device()), orig_sr, target_sr)[0]
    y_hat = out.cpu().numpy()
    if fix and len(y_hat) != n_samples:
        y_hat = np.pad(y_hat[:n_samples], (0, max(0, n_samples - len(y_hat))))
    if scale:
        y_hat = y_hat / np.sqrt(ratio)
    return np.asarray(y_hat, dtype=y.dtype if np.issubdtype(y.dtype, np.floating) else np.float32)


class ResampleDivergenceWarning(UserWarning):
    """The file's sample rate differs from the requested one and no ``res_type`` was given: the reference's
    ``librosa.load(file, sr=...)`` would convert with libsoxr's ``soxr_hq``, this path converts with ``'polyphase'``
    (``scipy.signal.resample_poly``).  The samples -- and everything derived from them -- differ from the reference's."""


def load(path, *, sr=22050, mono=True, offset=0.0, duration=None, res_type=None):
    """``librosa.load`` for RIFF WAV files (read with the standard library): PCM -> float32, channel mix-down and --
    when the file's rate differs from ``sr`` -- rate conversion.  16-bit files that need converting go to the GPU as
    int16 and are scaled, mixed and resampled there in one pass (K9).

    Rate conversion: librosa's default ``res_type`` is ``'soxr_hq'`` (libsoxr), which is not in this image and whose
    arithmetic cannot be restated; only ``'polyphase'`` is built.  ``res_type='polyphase'`` converts silently and
    bit-identically to librosa with that argument; any other explicit ``res_type`` raises ``NotImplementedError``; with
    no ``res_type`` (what the reference's call sites do, aegis_engine.py:24, aegis_engine_financial.py:45) a conversion
    that is actually needed warns with ``ResampleDivergenceWarning`` and uses ``'polyphase'`` -- never silently.
    ``offset`` / ``duration`` truncate to whole source frames as librosa does (``int(offset * sr_native)``).
    """
    import warnings
    import wave

    with wave.open(path, "rb") as w:
        file_sr, n_ch, width, n = w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()
        raw = w.readframes(n)
    start = int(offset * file_sr)
    stop = None if duration is None else start + int(duration * file_sr)
    need_resample = sr is not None and int(sr) != file_sr
    if need_resample and res_type is None:
        warnings.warn(f"{path}: {file_sr} Hz -> {sr} Hz converted with res_type='polyphase'; the reference's librosa.load "
                      "default is 'soxr_hq' (libsoxr, not available here): samples differ from the reference's. Pass "
                      "res_type='polyphase' to both to compare like with like.", ResampleDivergenceWarning, stacklevel=2)
        res_type = "polyphase"
    if need_resample and res_type != "polyphase":
        raise NotImplementedError(f"file is {file_sr} Hz, engine wants {sr} Hz and res_type={res_type!r}: only 'polyphase' is built")
    if need_resample and mono and width == 2:
        pcm = np.frombuffer(raw, dtype="<i2").reshape(-1, n_ch)[start:stop]
        if pcm.shape[0] == 0:
            return np.zeros(0, np.float32), int(sr)
        dev_pcm = torch.from_numpy(np.array(pcm, dtype=np.int16).reshape(1, -1)).to(_device())   # frombuffer views are read-only
        return core.resample_poly(dev_pcm, file_sr, int(sr), n_channels=n_ch)[0].cpu().numpy(), int(sr)
    if width == 2:
        data = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif width == 4:
        data = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    elif width == 1:
        data = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    else:
        raise NotImplementedError(f"{width * 8}-bit WAV")
    if n_ch > 1:
        data = data.reshape(-1, n_ch)
        data = data.mean(axis=1) if mono else data.T
    data = np.ascontiguousarray(data[..., start:stop], dtype=np.float32)
    if not need_resample:
        return data, file_sr
    if data.ndim != 1:
        raise NotImplementedError("resampling multi-channel output (mono=False)")
    return resample(data, orig_sr=file_sr, target_sr=int(sr), res_type=res_type), int(sr)
