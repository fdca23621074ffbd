"""CPU restatement of the reference's MIDI export as abstract message lists.  TEST INFRASTRUCTURE ONLY.

Follows ``AegisEngine.extract_events`` (aegis_engine.py:98-172) and ``AegisFinancialEngine.audio_to_midi_financial``
(aegis_engine_financial.py:185-243) up to the point where they hand the tracks to ``mido.MidiFile.save``.

PARITY UNPINNED for the file bytes: mido (``requirements.txt``, version unpinned) is not in this image, so neither the
reference's writer nor mido's encoder can be run here.  What IS checked: the native writer's bytes are parsed back by
``read_smf`` below -- a reader written from the SMF 1.0 specification (chunks, variable-length deltas, running status,
meta events) -- and the decoded messages must equal these lists.

A message is ``(delta_ticks, kind, a, b)``: kind 'program' (a = program), 'on' / 'off' (a = note, b = velocity),
'pitch' (a = wheel value -8192..8191), 'name' (a = track name), 'end'.
"""
from __future__ import annotations

import math
import struct

import numpy as np


def v1_tracks(events, sr, hop_length, midi_program=27, vibrato_rate=5.0, vibrato_depth=0.3):
    """[main messages, safe messages] of aegis_engine.py:98-172."""
    secs_per_frame = hop_length / sr
    ticks_per_sec = 960            # mido.second2tick(1.0, ticks_per_beat=480, tempo=500000)
    timeline = []                  # (tick, order of insertion, track, kind, a, b)
    for e in events:
        st = int(e["start"] * secs_per_frame * ticks_per_sec)
        et = int(e["end"] * secs_per_frame * ticks_per_sec)
        tech = e.get("technique")
        vel = e["velocity"]
        vel = int(vel * 0.6) if tech == "hammer_on" else (int(vel * 0.5) if tech == "pull_off" else vel)
        timeline.append((st, e["track"], "on", e["note"], vel))
        timeline.append((et, e["track"], "off", e["note"], 0))
        span = et - st
        if tech == "bend":
            slope = e.get("slope", 0.0)
            top = int((1 if slope > 0 else -1) * (min(2.0, abs(slope) * 10) / 2.0) * 8191)
            for i in range(15):
                p = i / 15
                timeline.append((st + int(p * span), e["track"], "pitch", int(top * (1 - (1 - p) ** 2)), 0))
            timeline.append((et, e["track"], "pitch", 0, 0))
        elif tech == "vibrato":
            secs = span / ticks_per_sec
            n = max(10, min(20, int(secs * vibrato_rate * 4)))
            for i in range(n):
                phase = (i / n) * secs * vibrato_rate * 2 * np.pi
                timeline.append((st + int((i / n) * span), e["track"], "pitch", int(np.sin(phase) * 8191 * vibrato_depth), 0))
            timeline.append((et, e["track"], "pitch", 0, 0))
    timeline.sort(key=lambda m: m[0])          # stable, like list.sort in the reference
    tracks = {"main": [(0, "program", midi_program, 0)], "safe": [(0, "program", midi_program, 0)]}
    last = {"main": 0, "safe": 0}
    for tick, tr, kind, a, b in timeline:
        tracks[tr].append((tick - last[tr], kind, a, b))
        last[tr] = tick
    return [tracks["main"] + [(0, "end", 0, 0)], tracks["safe"] + [(0, "end", 0, 0)]]


def v2_tracks(events, sr, hop_length):
    """[main messages, safe messages] of aegis_engine_financial.py:199-243."""
    tracks = {"main": [(0, "name", "Aegis Financial - Main", 0)], "safe": [(0, "name", "Aegis Financial - Safe", 0)]}
    last = {"main": 0, "safe": 0}
    ms_per_tick = 500 / 480
    ms_per_frame = (hop_length / sr) * 1000
    for e in events:
        start_ticks = int(e["start"] * ms_per_frame / ms_per_tick)
        duration_ticks = int((e["end"] - e["start"]) * ms_per_frame / ms_per_tick)
        tr = e["track"]
        tracks[tr].append((start_ticks - last[tr], "on", e["note"], e["velocity"]))
        tracks[tr].append((duration_ticks, "off", e["note"], 0))
        last[tr] = start_ticks + duration_ticks
    return [tracks["main"] + [(0, "end", 0, 0)], tracks["safe"] + [(0, "end", 0, 0)]]


def read_smf(data: bytes):
    """(format, ticks_per_beat, [track message lists]) of a Standard MIDI File; raises ValueError on malformed input."""
    if data[:4] != b"MThd" or struct.unpack(">I", data[4:8])[0] != 6:
        raise ValueError("bad header chunk")
    fmt, n_tracks, tpb = struct.unpack(">HHH", data[8:14])
    pos = 14
    tracks = []
    for _ in range(n_tracks):
        if data[pos:pos + 4] != b"MTrk":
            raise ValueError("bad track chunk")
        length = struct.unpack(">I", data[pos + 4:pos + 8])[0]
        body = data[pos + 8:pos + 8 + length]
        if len(body) != length:
            raise ValueError("truncated track")
        pos += 8 + length
        msgs, i, status = [], 0, None

        def varlen():
            nonlocal i
            v = 0
            while True:
                byte = body[i]
                i += 1
                v = (v << 7) | (byte & 0x7F)
                if not byte & 0x80:
                    return v

        while i < len(body):
            delta = varlen()
            if body[i] == 0xFF:
                kind = body[i + 1]
                i += 2
                n = varlen()
                payload = body[i:i + n]
                i += n
                status = None
                if kind == 0x2F:
                    msgs.append((delta, "end", 0, 0))
                elif kind == 0x03:
                    msgs.append((delta, "name", payload.decode("latin1"), 0))
                else:
                    raise ValueError(f"unexpected meta event {kind:#x}")
                continue
            if body[i] & 0x80:
                status = body[i]
                i += 1
            if status is None:
                raise ValueError("data byte without running status")
            hi = status & 0xF0
            if status & 0x0F:
                raise ValueError("channel other than 0")
            if hi == 0xC0:
                msgs.append((delta, "program", body[i], 0))
                i += 1
            elif hi in (0x90, 0x80):
                msgs.append((delta, "on" if hi == 0x90 else "off", body[i], body[i + 1]))
                i += 2
            elif hi == 0xE0:
                msgs.append((delta, "pitch", (body[i] | (body[i + 1] << 7)) - 8192, 0))
                i += 2
            else:
                raise ValueError(f"unexpected status {status:#x}")
        if not msgs or msgs[-1][1] != "end":
            raise ValueError("track does not end with end-of-track")
        tracks.append(msgs)
    if pos != len(data):
        raise ValueError("trailing bytes")
    return fmt, tpb, tracks
