"""ctypes binding of libaegis_b200.so (the C ABI declared in include/aegis_b200.h).

There is no CPU fallback: if the shared library is missing or fails to load, every entry point
raises ``AegisNativeError``.  The structures mirror the header field for field
(tests/test_abi.py compiles the header with gcc and compares sizes and offsets).
"""
from __future__ import annotations

import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AEGIS_B200_LIB") or os.path.join(PKG_DIR, "libaegis_b200.so")  # the override is for kernel A/B experiments
ABI_VERSION = 2

_f32p = C.c_void_p
_ptr = C.c_void_p
i32, i64, f64, f32 = C.c_int32, C.c_int64, C.c_double, C.c_float


class AegisNativeError(RuntimeError):
    pass


class StftParams(C.Structure):
    _fields_ = [
        ("y", _ptr), ("clip_stride", i64), ("n_samples", i64), ("n_clips", i32), ("hop", i32),
        ("pad", i32), ("n_frames", i32), ("window", _ptr), ("twiddle", _ptr),
        ("mag", _ptr), ("mag_clip_stride", i64), ("mag_row_stride", i32), ("n_mels", i32),
        ("mel", _ptr), ("mel_clip_stride", i64), ("mel_row_stride", i32), ("_reserved0", i32),
        ("mel_seg_start", _ptr), ("mel_rise_fall", _ptr),
        ("mel_max", _ptr), ("rms", _ptr), ("rms_clip_stride", i64),
    ]


class MelPostParams(C.Structure):
    _fields_ = [
        ("mel", _ptr), ("mel_clip_stride", i64), ("mel_row_stride", i32), ("n_mels", i32),
        ("n_clips", i32), ("n_frames", i32), ("mel_max", _ptr),
        ("s_db", _ptr), ("sdb_clip_stride", i64), ("sdb_row_stride", i32),
        ("rake_min_frames", i32), ("rake_max_frames", i32), ("onset_pad", i32),
        ("rake_ratio", f64), ("rake_mask", _ptr), ("onset_env", _ptr), ("env_minmax", _ptr),
        ("ref_power", _ptr), ("input_is_db", i32), ("_reserved0", i32),
    ]


class PeaksParams(C.Structure):
    _fields_ = [
        ("onset_env", _ptr), ("env_minmax", _ptr), ("n_clips", i32), ("n_frames", i32),
        ("pre_max", i32), ("post_max", i32), ("pre_avg", i32), ("post_avg", i32), ("wait", i32),
        ("normalize", i32), ("delta", f64), ("cand", _ptr), ("peaks", _ptr), ("n_peaks", _ptr),
    ]


class YinParams(C.Structure):
    _fields_ = [
        ("y", _ptr), ("clip_stride", i64), ("n_samples", i64), ("n_clips", i32), ("hop", i32),
        ("pad", i32), ("n_frames", i32), ("twiddle", _ptr), ("sr", f64), ("fmin", f64),
        ("min_period", i32), ("max_period", i32), ("n_pitch_bins", i32), ("bins_per_semitone", i32),
        ("n_thresholds", i32), ("max_cand", i32),
        ("thresholds", _ptr), ("beta_probs", _ptr), ("beta_cumsum", _ptr),
        ("boltz_fact", _ptr), ("boltz_exp", _ptr), ("no_trough_prob", f64),
        ("cand_bin", _ptr), ("cand_prob", _ptr), ("cand_count", _ptr), ("voiced_prob", _ptr),
        ("overflow", _ptr), ("cmnd_out", _ptr), ("block_sums", _ptr),
    ]


class ViterbiParams(C.Structure):
    _fields_ = [
        ("n_clips", i32), ("n_frames", i32), ("n_pitch_bins", i32), ("half_width", i32),
        ("n_variants", i32), ("n_interior_variants", i32), ("max_cand", i32), ("_reserved0", i32),
        ("cand_bin", _ptr), ("cand_prob", _ptr), ("cand_count", _ptr), ("voiced_prob", _ptr),
        ("lt_variants", _ptr), ("row_variant", _ptr), ("freqs", _ptr),
        ("log_tiny", f64), ("log_init_unvoiced", f64), ("fill_value", f64),
        ("backptr", _ptr), ("final_value", _ptr), ("states", _ptr), ("f0", _ptr), ("voiced_flag", _ptr),
    ]


class TrendParams(C.Structure):
    _fields_ = [
        ("x", _ptr), ("n_series", i32), ("n", i32), ("savgol_coeffs", _ptr),
        ("savgol_window", i32), ("sma_window", i32), ("ema_span", i32), ("boll_window", i32),
        ("macd_fast", i32), ("macd_slow", i32), ("macd_signal", i32), ("consensus_mask", i32),
        ("kalman_q", f64), ("kalman_r", f64), ("holt_alpha", f64), ("holt_beta", f64),
        ("boll_num_std", f64),
        ("compact", _ptr), ("scratch", _ptr), ("savgol", _ptr), ("kalman", _ptr), ("holt", _ptr),
        ("consensus", _ptr), ("consensus_conf", _ptr), ("sma", _ptr), ("ema", _ptr),
        ("boll_ma", _ptr), ("boll_upper", _ptr), ("boll_lower", _ptr),
        ("macd_line", _ptr), ("macd_sig", _ptr), ("macd_hist", _ptr),
    ]


class SynthParams(C.Structure):
    _fields_ = [
        ("out", _ptr), ("clip_stride", i64), ("n_events", i32), ("_reserved0", i32),
        ("ev_clip", _ptr), ("ev_start", _ptr), ("ev_len", _ptr), ("ev_period", _ptr),
        ("ev_amp", _ptr), ("ev_seed", _ptr), ("decay", f32), ("_reserved1", f32),
    ]


class GuitarParams(C.Structure):
    _fields_ = [
        ("s_db", _ptr), ("sdb_clip_stride", i64), ("sdb_row_stride", i32), ("n_mels", i32),
        ("n_clips", i32), ("n_frames", i32), ("mute_max_frames", i32), ("rake_frames", i32),
        ("fmin_hz", f64), ("f0", _ptr), ("voiced", _ptr), ("rake_in", _ptr),
        ("f0_out", _ptr), ("voiced_out", _ptr), ("rake_out", _ptr), ("mute_out", _ptr),
        ("distortion", _ptr), ("dist_work", _ptr),
    ]


class NoteEvent(C.Structure):
    _fields_ = [
        ("note", i32), ("start", i32), ("end", i32), ("velocity", i32), ("rms_energy", f32),
        ("track", C.c_uint8), ("technique", C.c_uint8), ("_pad", C.c_uint8 * 2),
        ("confidence", f64), ("slope", f64),
    ]


class NotesParams(C.Structure):
    _fields_ = [
        ("rake_mask", _ptr), ("f0", _ptr), ("voiced_flag", _ptr), ("voiced_prob", _ptr),
        ("rms", _ptr), ("rms_clip_stride", i64), ("pitch_index", _ptr), ("note_lut", _ptr),
        ("n_lut", i32), ("n_clips", i32), ("n_frames", i32), ("hop", i32),
        ("sr", f64), ("confidence_threshold", f64), ("noise_gate_db", f32),
        ("min_note_frames", i32), ("sustain_frames", i32), ("max_events", i32),
        ("events", _ptr), ("n_events", _ptr),
    ]


class FinEvent(C.Structure):
    _fields_ = [
        ("note", i32), ("start", i32), ("end", i32), ("velocity", i32),
        ("track", C.c_uint8), ("technique", C.c_uint8), ("slide", C.c_uint8), ("harmonic_valid", C.c_int8),
        ("_pad", C.c_uint8 * 4), ("confidence", f64),
    ]


class FinParams(C.Structure):
    _fields_ = [
        ("rake_mask", _ptr), ("f0", _ptr), ("voiced_flag", _ptr), ("voiced_prob", _ptr),
        ("rms", _ptr), ("rms_clip_stride", i64), ("n_clips", i32), ("n_frames", i32), ("hop", i32), ("_reserved0", i32),
        ("sr", f64), ("f0_clean", _ptr), ("semitones", _ptr), ("trend", _ptr), ("boll_upper", _ptr), ("boll_lower", _ptr),
        ("macd_line", _ptr), ("macd_hist", _ptr), ("confidence_threshold", f64), ("slide_threshold", f64),
        ("rsi_threshold", f64), ("noise_gate_db", f32), ("min_note_frames", i32), ("sustain_frames", i32),
        ("use_harmonic_filter", i32), ("harmonic_tolerance", i32), ("max_events", i32),
        ("events", _ptr), ("n_events", _ptr), ("threshold_out", _ptr), ("key_out", _ptr), ("key_confidence_out", _ptr),
        ("scratch", _ptr),
    ]


class ResampleParams(C.Structure):
    _fields_ = [
        ("x", _ptr), ("in_clip_stride", i64), ("n_in", i64), ("in_format", i32), ("n_channels", i32),
        ("n_clips", i32), ("up", i32), ("down", i32), ("n_taps", i32), ("taps", _ptr),
        ("n_pre_pad", i32), ("n_pre_remove", i32), ("out", _ptr), ("out_clip_stride", i64), ("n_out", i64),
    ]


class SmfOptions(C.Structure):
    _fields_ = [("hop", i32), ("midi_program", i32), ("sr", f64), ("vibrato_rate", f64), ("vibrato_depth", f64)]


ENTRY_POINTS = {
    "aegis_stft_fused": StftParams,
    "aegis_mel_post": MelPostParams,
    "aegis_onset_peaks": PeaksParams,
    "aegis_yin_candidates": YinParams,
    "aegis_viterbi": ViterbiParams,
    "aegis_trend_filters": TrendParams,
    "aegis_synth_ks": SynthParams,
    "aegis_guitar_filters": GuitarParams,
    "aegis_note_events": NotesParams,
    "aegis_fin_prepare": FinParams,
    "aegis_fin_events": FinParams,
    "aegis_resample_poly": ResampleParams,
}

_lib = None


def load() -> C.CDLL:
    """Load the library once; raise loudly if it is missing (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AegisNativeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback."
        )
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as e:
        raise AegisNativeError(f"cannot load {LIB_PATH}: {e}") from e
    lib.aegis_abi_version.restype = C.c_int
    lib.aegis_last_error.restype = C.c_char_p
    lib.aegis_device_sm_count.restype = C.c_int
    lib.aegis_note_events_bytes.restype = C.c_longlong
    lib.aegis_note_events_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.aegis_yin_workspace_bytes.restype = C.c_longlong
    lib.aegis_yin_workspace_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.aegis_fin_scratch_bytes.restype = C.c_longlong
    lib.aegis_fin_scratch_bytes.argtypes = [C.c_int, C.c_int]
    for name in ("aegis_smf_write_v1", "aegis_smf_write_v2"):
        fn = getattr(lib, name)
        fn.restype = C.c_longlong
        fn.argtypes = [C.c_void_p, C.c_int32, C.POINTER(SmfOptions), C.c_void_p, C.c_longlong]
    lib.aegis_musicxml_write.restype = C.c_longlong
    lib.aegis_musicxml_write.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_longlong]
    lib.aegis_tabs.restype = C.c_int
    lib.aegis_tabs.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    lib.aegis_guitar_blocks.restype = C.c_int
    lib.aegis_guitar_blocks.argtypes = [C.c_int]
    for name, struct in ENTRY_POINTS.items():
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(struct), C.c_void_p]
    if lib.aegis_abi_version() != ABI_VERSION:
        raise AegisNativeError(f"ABI mismatch: library {lib.aegis_abi_version()} != binding {ABI_VERSION}")
    _lib = lib
    return lib


def call(name: str, params: C.Structure, stream: int) -> None:
    lib = load()
    rc = getattr(lib, name)(C.byref(params), C.c_void_p(stream))
    if rc != 0:
        raise AegisNativeError(f"{name} failed ({rc}): {lib.aegis_last_error().decode(errors='replace')}")
