"""Engine facades with the reference's signatures, backed by the B200 kernels.

``AegisEngine`` mirrors ``aegis_engine.py:16-75,183-216`` (perception phase) and
``AegisFinancialEngine`` mirrors ``aegis_engine_financial.py:36-71``: same constructor arguments,
method names, keyword names/defaults, returned dict keys and dtypes.  The logic-filter phases
(``extract_events`` / ``audio_to_midi_financial``: note events, MIDI file) run on the GPU kernels K7 / K8 and the
library's own MIDI writer; ``generate_tabs`` / ``export_musicxml`` mirror ``aegis_engine_core/tabs.py``.
"""
from __future__ import annotations

import numpy as np
import torch

from . import batch, core, tables
from . import librosa_compat as librosa
from .vision import detect_rake_patterns
from .worker import _pyin_worker  # noqa: F401  (kept importable from here like aegis_engine.py:14)


def _as_audio(source, sr, start_time=0, end_time=None, res_type=None):
    """A numpy array is taken as audio already at ``sr``; a str/path goes through ``librosa_compat.load`` (which warns
    when it has to convert the rate without an explicit ``res_type``: the reference would use soxr_hq there)."""
    if isinstance(source, (str, bytes)) or hasattr(source, "__fspath__"):
        duration = (end_time - start_time) if end_time else None
        y, _ = librosa.load(source, sr=sr, offset=start_time, duration=duration, res_type=res_type)
        return y
    y = np.asarray(source, dtype=np.float32)
    s = int(start_time * sr)
    e = None if not end_time else s + int((end_time - start_time) * sr)
    return np.ascontiguousarray(y[s:e])


class AegisEngine:
    #: how files at another rate are converted to ``sample_rate``: ``None`` = warn (the reference's librosa.load uses
    #: soxr_hq, this path 'polyphase'), ``'polyphase'`` = convert silently (set it on both sides to compare like with like)
    res_type = None

    def __init__(self, sample_rate=44100, hop_length=512, n_fft=2048):
        self.sr = sample_rate
        self.hop_length = hop_length
        self.n_fft = n_fft

    # -- aegis_engine.py:22-27
    def load_audio(self, file_path, start_time=0, end_time=None):
        y = _as_audio(file_path, self.sr, start_time, end_time, self.res_type)
        if len(y) == 0:
            return y, np.zeros((128, 0), dtype=np.float32)
        S = librosa.feature.melspectrogram(y=y, sr=self.sr, n_fft=self.n_fft, hop_length=self.hop_length)
        S_dB = librosa.power_to_db(S, ref=np.max)
        return y, S_dB

    # -- aegis_engine.py:38-39
    def detect_rake_patterns(self, S_dB):
        return detect_rake_patterns(S_dB, self.hop_length, self.sr, 0.6)

    def _pyin(self, y):
        return librosa.pyin(y, fmin=librosa.note_to_hz("E2"), fmax=librosa.note_to_hz("C6"), sr=self.sr,
                            hop_length=self.hop_length)

    # -- aegis_engine.py:41-75
    def audio_to_midi(self, input_wav, output_mid, **kwargs):
        """Perception phase.  ``output_mid`` is ignored exactly as in the reference (:41-75 never use it)."""
        start_time, end_time = kwargs.get("start_time", 0), kwargs.get("end_time", None)
        rake_sensitivity = kwargs.get("rake_sensitivity", 0.6)
        y = _as_audio(input_wav, self.sr, start_time, end_time, self.res_type)
        if len(y) == 0:
            return None
        # one upload, every kernel on the device, one download (turbo_mode needs no special casing:
        # the serial full-clip result is what Turbo Mode approximates)
        yd = torch.from_numpy(y).to(librosa._device())[None]
        res = batch.analyze_batch(yd, sr=self.sr, hop_length=self.hop_length, rake_sensitivity=rake_sensitivity)
        return batch.to_host(res, 0, y=y)

    # -- aegis_engine.py:183-216
    def _parallel_pitch_tracking(self, y):
        """Turbo Mode.  The reference splits the clip into ``cpu_count()`` independent chunks (lossy at the
        seams); here the frame-parallel stages already use every SM, so the exact full-clip result is returned."""
        return self._pyin(np.asarray(y, dtype=np.float32))

    # -- aegis_engine.py:88-96: the note-event list extract_events builds before it writes the MIDI file
    def note_events(self, raw_data, **kwargs):
        """``get_midi_events`` on a perception result (same keyword arguments as ``extract_events``:
        ``confidence_threshold``, ``noise_gate_db``, ``sustain_ms``, ``min_note_duration_ms``), on the GPU (kernel K7)."""
        from .midi_logic import get_midi_events

        n = min(len(raw_data["f0"]), len(raw_data["rms"]), len(raw_data["rake_mask"]))   # aegis_engine.py:82-83
        return get_midi_events(raw_data["rake_mask"][:n], raw_data["f0"][:n], raw_data["voiced_flag"][:n],
                               raw_data["voiced_probs"][:n], raw_data["rms"][:n], self.sr, self.hop_length,
                               kwargs.pop("confidence_threshold", 0.70), **kwargs)

    # -- aegis_engine.py:77-181
    def extract_events(self, raw_data, output_mid, **kwargs):
        """Logic-filter phase: ``get_midi_events`` (K7) and, when ``output_mid`` is a path or a file-like object, the
        two-track MIDI file with bend / vibrato pitch-wheel curves (native writer, no mido)."""
        from . import midi_writer

        logic = {k: v for k, v in kwargs.items()
                 if k not in ("start_time", "end_time", "turbo_mode", "rake_sensitivity", "vibrato_rate", "vibrato_depth")}
        events = self.note_events(raw_data, **logic)
        if output_mid:
            midi_writer.write_midi(events, output_mid, self.sr, self.hop_length, **kwargs)
        return events

    # -- aegis_engine.py:32-36
    def generate_tabs(self, events):
        from .tabs import generate_tabs

        return generate_tabs(events)

    def export_musicxml(self, tab_data, xml_path):
        from .tabs import export_musicxml

        return export_musicxml(tab_data, xml_path)


class AegisFinancialEngine:
    """Perception methods of ``aegis_engine_financial.py:36-71`` (sr defaults to 22 050 there)."""

    res_type = None   # see AegisEngine.res_type

    def __init__(self, sample_rate=22050, hop_length=512, n_fft=2048):
        self.sr = sample_rate
        self.hop_length = hop_length
        self.n_fft = n_fft

    def load_audio(self, file_path, start_time=0, end_time=None):
        eng = AegisEngine(self.sr, self.hop_length, self.n_fft)
        eng.res_type = self.res_type
        return eng.load_audio(file_path, start_time, end_time)

    def detect_rake_patterns(self, S_dB, sensitivity=0.6):
        return detect_rake_patterns(S_dB, self.hop_length, self.sr, sensitivity)

    def pitch_tracking(self, y):
        return librosa.pyin(y, fmin=librosa.note_to_hz("E2"), fmax=librosa.note_to_hz("C6"), sr=self.sr,
                            hop_length=self.hop_length)

    def perception(self, input_wav, **kwargs):
        """Arrays ``audio_to_midi_financial`` hands to its note logic (:105-160): rake mask, pYIN with NaN
        f0, RMS, S_dB, plus the consensus trend of ``analyze_pitch_financial`` (financial_analysis.py:386-391).
        ``use_guitar_filters`` (default True, :128) applies ``apply_guitar_filters`` as :132-147 do."""
        y = _as_audio(input_wav, self.sr, kwargs.get("start_time", 0), kwargs.get("end_time", None), self.res_type)
        if len(y) == 0:
            return None
        yd = torch.from_numpy(y).to(librosa._device())[None]
        res = batch.analyze_batch(yd, sr=self.sr, hop_length=self.hop_length,
                                  rake_sensitivity=kwargs.get("rake_sensitivity", 0.6), with_sdb=True,
                                  with_trend=True, nan_to_num=False,
                                  with_guitar=kwargs.get("use_guitar_filters", True))
        host = batch.to_host(res, 0, y=y)
        host["sr"], host["hop_length"] = self.sr, self.hop_length  # financial_app_realtime.py:224-233
        return host

    def note_events(self, raw_data, confidence_threshold=None, **kwargs):
        """Phase 4 of ``audio_to_midi_financial`` (:150-171): the financial logic filter on a ``perception`` result;
        palm-muted frames are taken out of the voiced flags first (:147)."""
        from .midi_logic_financial import get_midi_events_financial
        voiced = np.asarray(raw_data["voiced_flag"], dtype=bool)
        if "mute_mask" in raw_data:
            voiced = voiced & ~np.asarray(raw_data["mute_mask"], dtype=bool)
        skip = ("confidence_threshold", "rake_sensitivity", "use_financial", "use_guitar_filters", "start_time", "end_time")
        return get_midi_events_financial(
            rake_mask=raw_data["rake_mask"], f0=raw_data["f0"], voiced_flag=voiced, active_probs=raw_data["voiced_probs"],
            rms=raw_data["rms"], sr=self.sr, hop_length=self.hop_length, confidence_threshold=confidence_threshold,
            use_financial=kwargs.get("use_financial", True), **{k: v for k, v in kwargs.items() if k not in skip})

    # -- aegis_engine_financial.py:73-246
    def audio_to_midi_financial(self, input_wav, output_mid, confidence_threshold=None, rake_sensitivity=0.6,
                                use_financial=True, **kwargs):
        """Whole v2 pipeline: perception (K1..K6), financial logic filter (K5 + K8), two-track MIDI file.  Returns
        ``output_mid``, or ``None`` when the audio is empty or no note is found, as the reference does."""
        from . import midi_writer

        raw = self.perception(input_wav, rake_sensitivity=rake_sensitivity, **kwargs)
        if raw is None:
            return None
        events = self.note_events(raw, confidence_threshold=confidence_threshold, use_financial=use_financial, **kwargs)
        if not events:
            return None
        midi_writer.write_midi_financial(events, output_mid, self.sr, self.hop_length)
        return output_mid

