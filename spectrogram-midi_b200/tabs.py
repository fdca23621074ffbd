"""Drop-in for ``aegis_engine_core/tabs.py``: tablature positions and the MusicXML export.

``generate_tabs`` (tabs.py:1-40) runs in the native library (``aegis_tabs``: one pass over the notes carrying the
fretting hand's "centre of gravity"), and so does ``export_musicxml`` (tabs.py:42-112, ``aegis_musicxml_write``): one
4/4 measure of quarter notes with string / fret technical marks and bend / slide / vibrato symbols, byte for byte the
document the reference's ElementTree code writes (pinned by golden files made with the real tabs.py).
"""
from __future__ import annotations

import ctypes
import numpy as np

from . import _native as nat


def tab_positions(notes) -> tuple:
    """(string, fret) int32 arrays for a sequence of MIDI notes; string 0 = unplayable in standard tuning."""
    notes = np.ascontiguousarray(notes, dtype=np.int32)
    strings = np.zeros(len(notes), dtype=np.int32)
    frets = np.zeros(len(notes), dtype=np.int32)
    rc = nat.load().aegis_tabs(notes.ctypes.data_as(ctypes.c_void_p), len(notes), strings.ctypes.data_as(ctypes.c_void_p),
                               frets.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        raise nat.AegisNativeError(f"aegis_tabs failed: {nat.load().aegis_last_error().decode(errors='replace')}")
    return strings, frets


def generate_tabs(events):
    strings, frets = tab_positions([e["note"] for e in events])
    return [{"time": e["start"], "string": int(s), "fret": int(f), "note": e["note"], "technique": e.get("technique"),
             "m_start": e["start"], "m_end": e["end"]}
            for e, s, f in zip(events, strings, frets) if s > 0]


_TECHNIQUE_CODE = {"vibrato": 1, "bend": 2, "slide": 3}


def musicxml_bytes(tab_data) -> bytes:
    """The MusicXML document of ``export_musicxml`` as bytes (native writer ``aegis_musicxml_write``): byte for byte
    what the reference's ``xml.etree.ElementTree`` tree serialises to."""
    n = len(tab_data)
    notes = np.ascontiguousarray([t["note"] for t in tab_data], dtype=np.int32)
    strings = np.ascontiguousarray([t["string"] for t in tab_data], dtype=np.int32)
    frets = np.ascontiguousarray([t["fret"] for t in tab_data], dtype=np.int32)
    tech = np.ascontiguousarray([_TECHNIQUE_CODE.get(t.get("technique"), 0) for t in tab_data], dtype=np.uint8)
    lib = nat.load()
    ptr = [a.ctypes.data_as(ctypes.c_void_p) if n else None for a in (notes, strings, frets, tech)]
    need = lib.aegis_musicxml_write(*ptr, n, None, 0)
    if need < 0:
        raise nat.AegisNativeError(f"aegis_musicxml_write failed: {lib.aegis_last_error().decode(errors='replace')}")
    buf = (ctypes.c_uint8 * need)()
    lib.aegis_musicxml_write(*ptr, n, buf, need)
    return bytes(buf)


def export_musicxml(tab_data, output_path):
    with open(output_path, "wb") as f:
        f.write(musicxml_bytes(tab_data))
    return output_path
