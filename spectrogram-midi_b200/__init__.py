"""B200-native audio-analysis front end for Aegis Engine (avabag01-ai/spectrogram-midi).

The hot path behind ``aegis_engine.py`` / ``aegis_engine_core/worker.py`` -- STFT magnitude, mel dB,
rake mask, pYIN (YIN candidates + HMM Viterbi), RMS, onset strength / peaks, trend filters -- as
hand-written sm_100a CUDA kernels behind a C ABI (``include/aegis_b200.h``), with the reference's
Python call signatures on top.  See DESIGN.md / INTEGRATION.md.  Import as ``spectrogram_midi_b200``.
"""
__version__ = "0.1.0"

from . import tables, corpus  # noqa: F401  (pure numpy/scipy; importable without CUDA)


def __getattr__(name):  # torch / the CUDA library load lazily so `import` works on a CPU-only box
    import importlib

    if name in {"core", "batch", "engine", "librosa_compat", "vision", "worker", "financial_filters",
                "financial_analysis", "guitar_specific", "midi_logic", "midi_logic_financial", "midi_writer", "tabs", "distributed", "build", "_native"}:
        return importlib.import_module(f"{__name__}.{name}")
    if name in {"AegisEngine", "AegisFinancialEngine"}:
        return getattr(importlib.import_module(f"{__name__}.engine"), name)
    raise AttributeError(name)
