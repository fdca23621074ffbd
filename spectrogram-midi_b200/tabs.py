"""Drop-in for ``aegis_engine_core/tabs.py``: tablature positions and the MusicXML export.

``generate_tabs`` (tabs.py:1-40) runs in the native library (``aegis_tabs``: one pass over the notes carrying the
fretting hand's "centre of gravity"); ``export_musicxml`` (tabs.py:42-112) is document assembly with the standard
library, element for element what the reference writes (one 4/4 measure of quarter notes with string / fret
technical marks and bend / slide / vibrato symbols).
"""
from __future__ import annotations

import ctypes
import xml.etree.ElementTree as ET

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


_STEPS = "CCDDEFFGGAAB"
_SHARPS = {1, 3, 6, 8, 10}


def _sub(parent, tag, text=None, **attrib):
    el = ET.SubElement(parent, tag, **attrib)
    if text is not None:
        el.text = str(text)
    return el


def export_musicxml(tab_data, output_path):
    score = ET.Element("score-partwise", version="3.1")
    _sub(_sub(_sub(score, "part-list"), "score-part", id="P1"), "part-name", "Aegis Guitar")
    measure = _sub(_sub(score, "part", id="P1"), "measure", number="1")
    attr = _sub(measure, "attributes")
    _sub(attr, "divisions", 1)
    _sub(_sub(attr, "key"), "fifths", 0)
    time = _sub(attr, "time")
    _sub(time, "beats", 4)
    _sub(time, "beat-type", 4)
    clef = _sub(attr, "clef")
    _sub(clef, "sign", "G")
    _sub(clef, "line", 2)
    _sub(_sub(attr, "staff-details"), "staff-lines", 6)
    for t in tab_data:
        note = _sub(measure, "note")
        pitch = _sub(note, "pitch")
        pc = t["note"] % 12
        _sub(pitch, "step", _STEPS[pc])
        if pc in _SHARPS:
            _sub(pitch, "alter", 1)
        _sub(pitch, "octave", (t["note"] // 12) - 1)
        _sub(note, "duration", 1)
        _sub(note, "type", "quarter")
        notations = _sub(note, "notations")
        technical = _sub(notations, "technical")
        _sub(technical, "string", t["string"])
        _sub(technical, "fret", t["fret"])
        technique = t.get("technique")
        if technique == "bend":
            _sub(_sub(technical, "bend"), "bend-alter", 2)
        elif technique == "slide":
            _sub(notations, "slur", type="start", number="1")
        elif technique == "vibrato":
            _sub(technical, "hammer-on", type="start")
            _sub(_sub(notations, "ornaments"), "wavy-line", type="start", number="1")
    ET.ElementTree(score).write(output_path, encoding="UTF-8", xml_declaration=True)
    return output_path
