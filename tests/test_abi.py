"""CPU tests of the boundary: header <-> ctypes layout, exported symbols, no-oracle / no-fallback rules."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "aegis_b200.h")
PKG = os.path.join(ROOT, "spectrogram-midi_b200")

import spectrogram_midi_b200  # noqa: E402,F401
from spectrogram_midi_b200 import _native  # noqa: E402

STRUCTS = {
    "aegis_stft_params": _native.StftParams, "aegis_melpost_params": _native.MelPostParams,
    "aegis_peaks_params": _native.PeaksParams, "aegis_yin_params": _native.YinParams,
    "aegis_viterbi_params": _native.ViterbiParams, "aegis_trend_params": _native.TrendParams,
    "aegis_synth_params": _native.SynthParams, "aegis_guitar_params": _native.GuitarParams,
    "aegis_note_event": _native.NoteEvent, "aegis_notes_params": _native.NotesParams,
    "aegis_fin_event": _native.FinEvent, "aegis_fin_params": _native.FinParams,
    "aegis_resample_params": _native.ResampleParams, "aegis_smf_options": _native.SmfOptions,
}


@pytest.fixture(scope="module")
def lib_path():
    if not os.path.exists(_native.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    return _native.LIB_PATH


def test_ctypes_layout_matches_header(tmp_path):
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for cname, st in STRUCTS.items():
        lines.append(f'printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in st._fields_:
            lines.append(f'printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    seen = 0
    for ln in out:
        if not ln:
            continue
        cname, field, val = ln.split()
        st = STRUCTS[cname]
        if field == "size":
            assert ctypes.sizeof(st) == int(val), cname
        else:
            assert getattr(st, field).offset == int(val), f"{cname}.{field}"
        seen += 1
    assert seen == sum(len(s._fields_) + 1 for s in STRUCTS.values())


def test_every_header_field_is_bound():
    text = open(HEADER).read()
    for cname, st in STRUCTS.items():
        end = text.index("} " + cname + ";")
        body = text[text.rindex("typedef struct {", 0, end) + len("typedef struct {"):end]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                names.append(re.findall(r"[A-Za-z_][A-Za-z0-9_]*", part)[-1])
        assert names == [f for f, _ in st._fields_], cname


def test_library_exports_every_declared_symbol(lib_path):
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    declared = set(re.findall(r"\b(aegis_[a-z0-9_]+)\s*\(", text))
    assert {"aegis_stft_fused", "aegis_mel_post", "aegis_onset_peaks", "aegis_yin_candidates", "aegis_viterbi",
            "aegis_trend_filters", "aegis_synth_ks", "aegis_guitar_filters", "aegis_guitar_blocks", "aegis_note_events", "aegis_note_events_bytes", "aegis_abi_version",
            "aegis_last_error"} <= declared
    lib = ctypes.CDLL(lib_path)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/aegis_b200.h but not exported"
    lib.aegis_abi_version.restype = ctypes.c_int
    assert lib.aegis_abi_version() == _native.ABI_VERSION
    assert set(_native.ENTRY_POINTS) <= declared


def test_argument_errors_are_reported_not_thrown(lib_path):
    lib = _native.load()
    p = _native.StftParams()  # all-null params: must fail cleanly without touching a GPU
    rc = lib.aegis_stft_fused(ctypes.byref(p), None)
    assert rc != 0 and b"aegis_stft_fused" in lib.aegis_last_error()
    y = _native.YinParams()
    assert lib.aegis_yin_candidates(ctypes.byref(y), None) != 0


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "oracle." not in src.replace("oracle/", ""), f


def test_no_cpu_fallback_without_cuda():
    import numpy as np
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from spectrogram_midi_b200 import core, librosa_compat

    with pytest.raises(_native.AegisNativeError):
        librosa_compat.pyin(np.zeros(4096, np.float32), fmin=82.4, fmax=1046.5, sr=22050)
    with pytest.raises(_native.AegisNativeError):
        core.stft_features(torch.zeros(1, 4096))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_native.AegisNativeError):
        _native.load()


def test_host_side_frame_limits_and_note_table():
    """Host logic of the K6 / K7 bindings (no device work): the frame limits use the reference's expressions and the
    note table maps every distinct f0 value with int(round(hz_to_midi(f)))."""
    import numpy as np

    from spectrogram_midi_b200 import core, midi_logic, tables

    assert core.guitar_frame_limits(512, 22050) == (2, 1)        # int(50 / 23.2), int(30 / 23.2)
    assert core.guitar_frame_limits(512, 44100) == (4, 2)
    assert core.note_frame_limits(22050, 512) == (2, 2)          # int(0.05 * 22050 / 512) twice
    assert core.note_frame_limits(44100, 512) == (4, 4)
    assert core.NOTE_EVENT_DTYPE.itemsize == ctypes.sizeof(_native.NoteEvent) == 40
    for name in ("note", "start", "end", "velocity", "rms_energy", "track", "technique", "confidence", "slope"):
        assert core.NOTE_EVENT_DTYPE.fields[name][1] == getattr(_native.NoteEvent, name).offset, name
    assert core.FIN_EVENT_DTYPE.itemsize == ctypes.sizeof(_native.FinEvent) == 32
    for name in ("note", "start", "end", "velocity", "track", "technique", "slide", "harmonic_valid", "confidence"):
        assert core.FIN_EVENT_DTYPE.fields[name][1] == getattr(_native.FinEvent, name).offset, name
    f0 = np.array([0.0, 110.0, 110.0, np.nan, 440.0, 82.4068892282175 * 2 ** (5 / 120.0), -3.0])
    idx, lut = midi_logic._note_lut(f0)
    assert idx.dtype == np.uint16 and lut.dtype == np.int16 and len(lut) == 3
    pos = f0 > 0
    assert (idx[~pos] == 65535).all()
    want = np.array([int(round(float(tables.hz_to_midi(v)))) for v in f0[pos]])
    assert np.array_equal(lut[idx[pos]], want) and set(want) >= {45, 69}
