"""Generate tests/golden/tabs_golden.npz from the REAL reference file ``aegis_engine_core/tabs.py`` (stdlib only).

Run in the build container only (``/root/reference`` must exist):

    python tests/golden/make_golden_tabs.py

Seeded note-event lists (in-range runs, big leaps, notes below E2 / above the 24th fret of the first string that the
reference skips, every technique label) go through ``generate_tabs`` and ``export_musicxml``; inputs, positions and
the XML bytes are stored side by side.
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as MG  # noqa: E402

TECH = [None, "vibrato", "bend", "slide", "hammer_on", "pull_off"]


def main():
    tabs = MG._load("ref_tabs", f"{MG.REF}/aegis_engine_core/tabs.py")
    rng = np.random.default_rng(77)
    out = {}
    for case, n in (("walk", 120), ("leaps", 80), ("short", 3), ("empty", 0)):
        if case == "walk":
            notes = np.clip(52 + np.cumsum(rng.integers(-3, 4, n)), 36, 92)
        else:
            notes = rng.integers(34, 95, n)
        start = np.cumsum(rng.integers(3, 40, n)) if n else np.zeros(0, int)
        events = [{"note": int(p), "start": int(s), "end": int(s + rng.integers(2, 30)), "technique": TECH[int(rng.integers(0, 6))]}
                  for p, s in zip(notes, start)]
        tab = tabs.generate_tabs(events)
        path = f"/tmp/_tabs_{case}.xml"
        tabs.export_musicxml(tab, path)
        k = f"tabs/{case}"
        out[f"{k}/note"] = np.array([e["note"] for e in events], dtype=np.int64)
        out[f"{k}/start"] = np.array([e["start"] for e in events], dtype=np.int64)
        out[f"{k}/end"] = np.array([e["end"] for e in events], dtype=np.int64)
        out[f"{k}/technique"] = np.array([TECH.index(e["technique"]) for e in events], dtype=np.int64)
        out[f"{k}/tab"] = np.array([[t["time"], t["string"], t["fret"], t["note"], TECH.index(t["technique"]), t["m_start"], t["m_end"]]
                                    for t in tab], dtype=np.int64).reshape(-1, 7)
        out[f"{k}/xml"] = np.frombuffer(open(path, "rb").read(), dtype=np.uint8)
        os.remove(path)
        print(case, len(events), "events ->", len(tab), "positions,", len(out[f"{k}/xml"]), "XML bytes")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tabs_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
