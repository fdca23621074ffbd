"""Import alias: ``import spectrogram_midi_b200`` loads the package in ``spectrogram-midi_b200/``.

The package directory keeps the repository's name (with its hyphen), which Python cannot import
directly; this module replaces itself in ``sys.modules`` with that package.
"""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "spectrogram-midi_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
