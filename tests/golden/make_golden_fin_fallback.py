"""Generate tests/golden/fin_fallback_golden.npz: ``get_midi_events_financial(use_financial=False)`` of the REAL
``aegis_engine_core_v2/midi_logic_financial.py`` (the fallback branch :178-196 + per-event detect_articulations_financial)
on seeded frame series.  Build container only (``/root/reference`` must exist):

    python tests/golden/make_golden_fin_fallback.py
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import make_golden as MG  # noqa: E402
import make_golden_financial_events as MF  # noqa: E402

TECH = [None, "normal", "bend", "vibrato", "noise", "slide"]


def rows(events):
    ints = np.array([[e["note"], e["start"], e["end"], e["velocity"], int(e["track"] == "main"),
                      TECH.index(e.get("technique")), int("technique" in e)] for e in events], dtype=np.int64).reshape(-1, 7)
    return ints, np.array([e["confidence"] for e in events], dtype=np.float64)


def main():
    MG._install_shims()
    pkg = types.ModuleType("refv2")
    pkg.__path__ = [f"{MG.REF}/aegis_engine_core_v2"]
    sys.modules["refv2"] = pkg
    for m in ("financial_filters", "financial_analysis", "harmonic_analysis"):
        MG._load(f"refv2.{m}", f"{MG.REF}/aegis_engine_core_v2/{m}.py", "refv2")
    logic = MG._load("refv2.midi_logic_financial", f"{MG.REF}/aegis_engine_core_v2/midi_logic_financial.py", "refv2")
    out, names = {}, []
    for name, seed, n, sr, kw in [("synth1", 1, 431, 22050, {}), ("synth2", 2, 1292, 22050, {}), ("synth3", 3, 900, 44100, {}),
                                  ("synth4_thr", 4, 700, 22050, {"confidence_threshold": 0.55, "sustain_ms": 120}),
                                  ("synth7_gate", 7, 600, 22050, {"noise_gate_db": -30, "min_note_duration_ms": 0})]:
        rake, f0, vf, vp, rms = MF.synthetic_frames(seed, n, sr)
        kw = dict(kw)
        thr = kw.pop("confidence_threshold", None)
        with contextlib.redirect_stdout(io.StringIO()):
            ev = logic.get_midi_events_financial(rake_mask=rake.copy(), f0=f0.copy(), voiced_flag=vf.copy(), active_probs=vp.copy(),
                                                 rms=rms.copy(), sr=sr, hop_length=512, confidence_threshold=thr,
                                                 use_financial=False, **kw)
        ints, conf = rows(ev)
        k = f"fb/{name}"
        names.append(name)
        out[f"{k}/rake_mask"], out[f"{k}/f0"], out[f"{k}/voiced_flag"], out[f"{k}/voiced_prob"], out[f"{k}/rms"] = rake, f0, vf, vp, rms
        out[f"{k}/args"] = np.array([sr, np.nan if thr is None else thr, kw.get("noise_gate_db", -40), kw.get("sustain_ms", 50),
                                     kw.get("min_note_duration_ms", 50)], dtype=np.float64)
        out[f"{k}/events"], out[f"{k}/confidence"] = ints, conf
        print(f"{name:14s} events {len(ev):3d} techniques {sorted(set(ints[:, 5].tolist()))} open-ended {int((ints[:, 6] == 0).sum())}")
    out["fb/names"] = np.array(names)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fin_fallback_golden.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
