"""CPU oracle for the Aegis audio-analysis hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it, and there only as the checker / the timed CPU baseline.  The product package
(``spectrogram-midi_b200/``) never imports this package and fails loudly when its CUDA library is
missing.

Parity status
-------------
* ``oracle/librosa_ref.py`` restates the subset of third-party **librosa** (un-vendored, unpinned:
  ``/root/reference/requirements.txt:1``; semantics followed: librosa >= 0.10, i.e.
  ``pad_mode='constant'``) that the reference calls on this path.  librosa is not installable in
  the build container (no network, not in the wheelhouse) and the reference pins no numerical
  result of pyin / melspectrogram / rms anywhere (it has no test-suite): **parity unpinned** for
  those functions.  What *is* pinned: the mel filterbank is cross-checked against two independent
  implementations that are themselves tested against librosa (``transformers.audio_utils`` and
  ``torchaudio``), the STFT against ``scipy.signal.stft``, the YIN difference function against
  its O(n^2) time-domain definition, and the Viterbi decoder against brute-force path
  enumeration (``tests/test_oracle.py``).
* ``oracle/reference_files.py`` restates the reference's *own* numpy code on the path
  (``aegis_engine_core/vision.py``, ``aegis_engine_core_v2/financial_filters.py``,
  ``aegis_engine_core_v2/financial_analysis.py``, ``aegis_engine_core/midi_logic.py``).  Those
  restatements are **pinned**: ``tests/golden/make_golden.py`` imports the real reference files
  from ``/root/reference`` in the build container and commits their outputs on seeded inputs as
  ``tests/golden/*.npz``; ``tests/test_oracle.py`` checks the restatement against them.
"""
