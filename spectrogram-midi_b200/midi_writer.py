"""Standard MIDI File output for note-event lists, in the native library (``aegis_smf_write_v1`` / ``_v2``).

``write_midi`` reproduces the file ``AegisEngine.extract_events`` saves through mido (aegis_engine.py:98-179): two
tracks (main / safe), a program change, note on / off with the hammer-on / pull-off velocity scaling, bend and vibrato
pitch-wheel curves, stable sort by tick, per-track delta times.  ``write_midi_financial`` reproduces the file of
``AegisFinancialEngine.audio_to_midi_financial`` (aegis_engine_financial.py:185-246).  mido itself is not needed
(and is not in this image: the byte layout follows the SMF 1.0 specification and mido's writer -- running status,
end-of-track appended -- and is checked by parsing the bytes back, not against mido).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _native as nat
from . import core

_V1_TECHNIQUES = {name: code for code, name in enumerate(core.TECHNIQUES)}
_V2_ARTICULATIONS = {name: code for code, name in enumerate(core.FIN_ARTICULATIONS)}
_V2_SLIDES = {name: code for code, name in enumerate(core.FIN_SLIDES)}


def records_v1(events) -> np.ndarray:
    """List of ``get_midi_events`` dicts -> ``core.NOTE_EVENT_DTYPE`` records."""
    rec = np.zeros(len(events), dtype=core.NOTE_EVENT_DTYPE)
    for r, e in zip(rec, events):
        r["note"], r["start"], r["end"], r["velocity"] = e["note"], e["start"], e["end"], e["velocity"]
        r["rms_energy"] = e.get("rms_energy", 0.0)
        r["track"] = 1 if e["track"] == "main" else 0
        r["technique"] = _V1_TECHNIQUES[e.get("technique")]
        r["confidence"], r["slope"] = e.get("confidence", 0.0), e.get("slope", 0.0)
    return rec


def records_v2(events) -> np.ndarray:
    """List of ``get_midi_events_financial`` dicts -> ``core.FIN_EVENT_DTYPE`` records."""
    rec = np.zeros(len(events), dtype=core.FIN_EVENT_DTYPE)
    for r, e in zip(rec, events):
        r["note"], r["start"], r["end"], r["velocity"] = e["note"], e["start"], e["end"], e["velocity"]
        r["track"] = 1 if e["track"] == "main" else 0
        r["technique"] = _V2_ARTICULATIONS.get(e.get("technique"), 0)
        r["slide"] = _V2_SLIDES.get(e.get("financial_slide"), 0)
        r["harmonic_valid"] = -1 if "harmonic_valid" not in e else int(bool(e["harmonic_valid"]))
        r["confidence"] = e.get("confidence", 0.0)
    return rec


def _smf(fn_name: str, rec: np.ndarray, opt: nat.SmfOptions) -> bytes:
    lib = nat.load()
    fn = getattr(lib, fn_name)
    rec = np.ascontiguousarray(rec)
    ptr = rec.ctypes.data_as(ctypes.c_void_p) if len(rec) else None
    need = fn(ptr, len(rec), ctypes.byref(opt), None, 0)
    if need < 0:
        raise ValueError(f"{fn_name}: {lib.aegis_last_error().decode(errors='replace')}")   # mido raises ValueError for bad data bytes
    buf = (ctypes.c_uint8 * need)()
    if fn(ptr, len(rec), ctypes.byref(opt), buf, need) != need:
        raise nat.AegisNativeError(f"{fn_name}: size changed between calls")
    return bytes(buf)


def smf_bytes(events, sr, hop_length, midi_program=27, vibrato_rate=5.0, vibrato_depth=0.3) -> bytes:
    """The v1 file as bytes; ``events`` is a list of dicts or an array of ``NOTE_EVENT_DTYPE`` records."""
    rec = events if isinstance(events, np.ndarray) else records_v1(events)
    opt = nat.SmfOptions(int(hop_length), int(midi_program), float(sr), float(vibrato_rate), float(vibrato_depth))
    return _smf("aegis_smf_write_v1", rec, opt)


def smf_bytes_financial(events, sr, hop_length) -> bytes:
    """The v2 file as bytes; ``events`` is a list of dicts or an array of ``FIN_EVENT_DTYPE`` records."""
    rec = events if isinstance(events, np.ndarray) else records_v2(events)
    return _smf("aegis_smf_write_v2", rec, nat.SmfOptions(int(hop_length), 0, float(sr), 0.0, 0.0))


def _save(data: bytes, output_mid):
    if hasattr(output_mid, "write"):       # file-like objects (BytesIO), as aegis_engine.py:175-178
        output_mid.write(data)
    else:
        with open(output_mid, "wb") as f:
            f.write(data)


def write_midi(events, output_mid, sr, hop_length, **kwargs):
    _save(smf_bytes(events, sr, hop_length, kwargs.get("midi_program", 27), kwargs.get("vibrato_rate", 5.0),
                    kwargs.get("vibrato_depth", 0.3)), output_mid)


def write_midi_financial(events, output_mid, sr, hop_length):
    _save(smf_bytes_financial(events, sr, hop_length), output_mid)
