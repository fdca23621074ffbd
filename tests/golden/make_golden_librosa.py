"""Pin the librosa half of the oracle: run the REAL librosa on the seeded corpus and store its outputs.

librosa is the reference's third-party dependency for this path (requirements.txt:1, unpinned; call sites
aegis_engine.py:24-26,63,67,70, aegis_engine_core/worker.py:9-15, aegis_engine_financial.py:45-51,63-69,154).  It is not
installable in the build image (no network, not in the wheelhouse), so ``oracle/librosa_ref.py`` is a restatement whose
parity is UNPINNED.  The moment a machine has librosa, run

    python tests/golden/make_golden_librosa.py

and commit ``tests/golden/librosa_golden.npz``: ``tests/test_librosa_pin.py`` (skipped while the file is absent) then
compares the oracle with librosa itself -- STFT magnitude, mel power, dB image, RMS, pYIN (f0, voiced flags, voiced
probabilities), onset envelope and onset frames -- and records the librosa / numpy / scipy / numba versions.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

SIGNALS = {   # name -> (maker, sr); the same seeded signals the GPU parity tests use
    "track22050": ("test_track", (22050, 0), 22050),
    "track44100": ("test_track", (44100, 0), 44100),
    "bench22050": ("benchmark_signal", (22050, 0), 22050),
    "clip3": ("random_clip", (3, 6.0, 22050), 22050),
    "clip7": ("random_clip", (7, 12.0, 22050), 22050),
}


def probe():
    """(available, versions or reason) -- also used by __graft_entry__.smoke() to report what the GPU box has."""
    try:
        import librosa
    except Exception as e:   # ImportError, or a broken numba / soundfile underneath
        return False, f"{type(e).__name__}: {e}"
    import scipy

    vers = {"librosa": librosa.__version__, "numpy": np.__version__, "scipy": scipy.__version__}
    try:
        import numba

        vers["numba"] = numba.__version__
    except Exception:
        vers["numba"] = None
    return True, vers


def main():
    ok, info = probe()
    if not ok:
        print(f"librosa is not importable here ({info}): nothing written, the oracle stays unpinned")
        return 1
    import librosa

    import spectrogram_midi_b200  # noqa: F401
    from spectrogram_midi_b200 import corpus

    out = {"versions": np.array([f"{k}={v}" for k, v in info.items()])}
    for name, (maker, args, sr) in SIGNALS.items():
        y = np.asarray(getattr(corpus, maker)(*args), dtype=np.float32)
        out[f"{name}/y"] = y
        out[f"{name}/sr"] = np.array([sr])
        out[f"{name}/stft_mag"] = np.abs(librosa.stft(y, n_fft=2048, hop_length=512)).astype(np.float32)
        S = librosa.feature.melspectrogram(y=y, sr=sr, n_fft=2048, hop_length=512)
        out[f"{name}/mel"] = S
        out[f"{name}/S_dB"] = librosa.power_to_db(S, ref=np.max)
        out[f"{name}/rms"] = librosa.feature.rms(y=y, hop_length=512)[0]
        for tag, fmax in (("C6", "C6"), ("E6", "E6")):
            f0, vf, vp = librosa.pyin(y, fmin=librosa.note_to_hz("E2"), fmax=librosa.note_to_hz(fmax), sr=sr, hop_length=512)
            out[f"{name}/pyin_{tag}/f0"], out[f"{name}/pyin_{tag}/voiced_flag"], out[f"{name}/pyin_{tag}/voiced_prob"] = f0, vf, vp
        env = librosa.onset.onset_strength(y=y, sr=sr, hop_length=512)
        out[f"{name}/onset_env"] = env
        out[f"{name}/onset_frames"] = librosa.onset.onset_detect(onset_envelope=env, sr=sr, hop_length=512)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "librosa_golden.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays ({os.path.getsize(path) / 1024:.0f} KiB), versions {info}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
