"""Serialisation after the path: tablature positions and MusicXML against the real
``aegis_engine_core/tabs.py`` (golden), Standard MIDI File bytes of the native writers parsed back with an
independent reader and compared with the restated message lists (mido is not in this image: bytes unpinned against
mido, see oracle/midi_messages.py).  Host code only -- no GPU needed, the library just has to load."""
import io
import os
import sys
import xml.etree.ElementTree as ET

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import midi_messages as MM  # noqa: E402
import spectrogram_midi_b200 as P  # noqa: E402

TECH = [None, "vibrato", "bend", "slide", "hammer_on", "pull_off"]


@pytest.fixture(scope="module")
def tabs_golden():
    with np.load(os.path.join(ROOT, "tests", "golden", "tabs_golden.npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def test_tabs_and_musicxml_match_reference(tabs_golden, tmp_path):
    g = tabs_golden
    skipped = 0
    for case in ("walk", "leaps", "short", "empty"):
        k = f"tabs/{case}"
        events = [{"note": int(n), "start": int(s), "end": int(e), "technique": TECH[int(t)]}
                  for n, s, e, t in zip(g[f"{k}/note"], g[f"{k}/start"], g[f"{k}/end"], g[f"{k}/technique"])]
        tab = P.tabs.generate_tabs(events)
        rows = np.array([[t["time"], t["string"], t["fret"], t["note"], TECH.index(t["technique"]), t["m_start"], t["m_end"]] for t in tab],
                        dtype=np.int64).reshape(-1, 7)
        np.testing.assert_array_equal(rows, g[f"{k}/tab"], err_msg=case)
        skipped += len(events) - len(tab)
        path = str(tmp_path / f"{case}.xml")
        assert P.tabs.export_musicxml(tab, path) == path
        assert open(path, "rb").read() == g[f"{k}/xml"].tobytes(), case
        root = ET.fromstring(P.tabs.musicxml_bytes(tab))      # and it is well-formed XML with one <note> per position
        assert root.tag == "score-partwise" and len(root.findall("./part/measure/note")) == len(tab)
    assert skipped > 10     # unplayable notes are dropped, as in the reference
    eng = P.engine.AegisEngine(22050)
    assert eng.generate_tabs(events) == tab     # the engine methods of aegis_engine.py:32-36


def _random_v1_events(rng, n):
    events, t = [], 0
    for _ in range(n):
        t += int(rng.integers(0, 25))
        end = t + int(rng.integers(0, 60))
        tech = TECH[int(rng.integers(0, 6))]
        events.append({"note": int(rng.integers(36, 90)), "start": t, "end": end, "velocity": int(rng.integers(0, 128)),
                       "track": "main" if rng.random() < 0.6 else "safe", "technique": tech, "confidence": 0.5,
                       "rms_energy": -20.0, "slope": float(rng.uniform(-0.4, 0.4)) if tech in ("bend", "slide", "vibrato") else 0.0})
        t = end + 1
    return events


@pytest.mark.parametrize("sr,hop", [(22050, 512), (44100, 512), (44100, 256)])
def test_v1_midi_file_decodes_to_the_reference_messages(sr, hop):
    rng = np.random.default_rng(sr + hop)
    for n, kw in ((0, {}), (1, {}), (40, {}), (150, {"midi_program": 30, "vibrato_rate": 6.5, "vibrato_depth": 0.8})):
        events = _random_v1_events(rng, n)
        data = P.midi_writer.smf_bytes(events, sr, hop, **kw)
        assert data[:14] == b"MThd\x00\x00\x00\x06\x00\x01\x00\x02\x01\xe0"     # format 1, two tracks, 480 ticks per beat
        fmt, tpb, tracks = MM.read_smf(data)
        want = MM.v1_tracks(events, sr, hop, **kw)
        assert (fmt, tpb, len(tracks)) == (1, 480, 2)
        assert tracks[0] == want[0] and tracks[1] == want[1]
        kinds = {m[1] for tr in tracks for m in tr}
        if n >= 40:
            assert kinds == {"program", "on", "off", "pitch", "end"}
    # file-like and path targets, as aegis_engine.py:175-178
    buf = io.BytesIO()
    P.midi_writer.write_midi(events, buf, sr, hop, **kw)
    assert buf.getvalue() == data


def test_v2_midi_file_decodes_to_the_reference_messages(tmp_path):
    rng = np.random.default_rng(3)
    for sr in (22050, 44100):
        events = [{k: v for k, v in e.items() if k != "slope"} for e in _random_v1_events(rng, 90)]
        for e in events:
            e["technique"] = [None, "normal", "bend", "vibrato", "noise"][int(rng.integers(0, 5))]
        path = str(tmp_path / f"v2_{sr}.mid")
        P.midi_writer.write_midi_financial(events, path, sr, 512)
        fmt, tpb, tracks = MM.read_smf(open(path, "rb").read())
        assert (fmt, tpb) == (1, 480) and tracks == MM.v2_tracks(events, sr, 512)
        assert tracks[0][0] == (0, "name", "Aegis Financial - Main", 0) and tracks[1][0][2] == "Aegis Financial - Safe"


def test_writer_rejects_what_mido_rejects():
    bad = [{"note": 140, "start": 0, "end": 5, "velocity": 90, "track": "main", "technique": None}]
    with pytest.raises(ValueError):
        P.midi_writer.smf_bytes(bad, 22050, 512)
    vib = [{"note": 60, "start": 0, "end": 50, "velocity": 90, "track": "main", "technique": "vibrato", "slope": 0.0}]
    with pytest.raises(ValueError):
        P.midi_writer.smf_bytes(vib, 22050, 512, vibrato_depth=1.5)      # pitch wheel beyond +-8191
    with pytest.raises(ValueError):
        MM.read_smf(b"MThd\x00\x00\x00\x06\x00\x01\x00\x01\x01\xe0MTrk\x00\x00\x00\x02\x00\x40")   # the reader is strict too


def test_running_status_and_varlen_layout():
    """Two notes far apart on one track: every byte of the track, by hand from the SMF specification."""
    # sr 1024 / hop 512: half a second per frame = 480 ticks per frame, exactly representable
    events = [{"note": 60, "start": 0, "end": 1, "velocity": 100, "track": "main", "technique": None},
              {"note": 62, "start": 21, "end": 22, "velocity": 101, "track": "main", "technique": None}]
    data = P.midi_writer.smf_bytes(events, 1024, 512)
    main = bytes([0x00, 0xC0, 27,               # program change
                  0x00, 0x90, 60, 100,          # note on at tick 0
                  0x83, 0x60, 0x80, 60, 0,      # note off 480 ticks later: 480 = 3 << 7 | 0x60; status changes, so it is written
                  0xCB, 0x00, 0x90, 62, 101,    # delta 9600 = 0x4B << 7 | 0x00
                  0x83, 0x60, 0x80, 62, 0,
                  0x00, 0xFF, 0x2F, 0x00])
    safe = bytes([0x00, 0xC0, 27, 0x00, 0xFF, 0x2F, 0x00])
    want = (b"MThd\x00\x00\x00\x06\x00\x01\x00\x02\x01\xe0" + b"MTrk" + len(main).to_bytes(4, "big") + main +
            b"MTrk" + len(safe).to_bytes(4, "big") + safe)
    assert data == want
    bend = [{"note": 60, "start": 0, "end": 3, "velocity": 100, "track": "safe", "technique": "bend", "slope": 0.1}]
    _, _, tracks = MM.read_smf(P.midi_writer.smf_bytes(bend, 1024, 512))
    wheel = [m for m in tracks[1] if m[1] == "pitch"]
    assert len(wheel) == 16 and wheel[0][2] == 0 and wheel[-1][2] == 0
    assert max(m[2] for m in wheel) == int(4095 * (1 - (1 - 14 / 15) ** 2)) == 4076
    raw = P.midi_writer.smf_bytes(bend, 1024, 512)
    # fifteen consecutive wheel messages share one status byte; the closing one follows the note-off and needs its own
    assert raw[14:].count(b"\xe0") == 2


def test_midi_files_of_the_golden_reference_events(golden, fin_golden):
    """The note events the REAL midi_logic.py / midi_logic_financial.py produced (committed golden vectors) written by
    the native writers: the files decode to the message lists of the restated export code."""
    from oracle import financial_events as FE

    n_v1 = n_v2 = 0
    for name in ("track22050", "track44100", "clip7"):
        k = f"midi/{name}"
        sr = int(golden[f"{k}/sr"][0])
        events = [{"note": int(r[0]), "start": int(r[1]), "end": int(r[2]), "velocity": int(r[3]), "track": "main" if r[4] else "safe",
                   "technique": TECH[int(r[5])], "confidence": float(f[0]), "rms_energy": float(f[1]), "slope": float(f[2])}
                  for r, f in zip(golden[f"{k}/events"], golden[f"{k}/event_float"])]
        _, tpb, tracks = MM.read_smf(P.midi_writer.smf_bytes(events, sr, 512))
        assert tpb == 480 and tracks == MM.v1_tracks(events, sr, 512)
        n_v1 += len(events)
    for name in [str(v) for v in fin_golden["fin/names"]]:
        k = f"fin/{name}"
        sr = int(fin_golden[f"{k}/args"][0])
        events = [{"note": int(r[0]), "start": int(r[1]), "end": int(r[2]), "velocity": int(r[3]), "track": "main" if r[4] else "safe",
                   "technique": FE.ARTIC[int(r[5])], "financial_slide": FE.SLIDE[int(r[6])], "confidence": float(c)}
                  for r, c in zip(fin_golden[f"{k}/events"], fin_golden[f"{k}/confidence"])]
        _, _, tracks = MM.read_smf(P.midi_writer.smf_bytes_financial(events, sr, 512))
        assert tracks == MM.v2_tracks(events, sr, 512), name
        n_v2 += len(events)
    assert n_v1 >= 10 and n_v2 > 200

