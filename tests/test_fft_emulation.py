"""CPU test of the device FFT passes: the header csrc/fft2048.cuh compiled for the host, its 128
"threads" run in a loop (in both orders, to expose intra-pass hazards) against numpy's FFT."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

import spectrogram_midi_b200  # noqa: E402,F401
from spectrogram_midi_b200 import tables  # noqa: E402


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    so = tmp_path_factory.mktemp("emul") / "fft_emul.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(so),
                    os.path.join(ROOT, "tests", "emulation", "fft_emul.cpp")], check=True)
    return ctypes.CDLL(str(so))


@pytest.mark.parametrize("fn", ["emul_fft2048", "emul_fft2048_inplace"])
@pytest.mark.parametrize("reverse", [0, 1])
def test_fft_passes_match_numpy(emul, fn, reverse):
    rng = np.random.default_rng(0)
    tw = tables.fft_twiddles()
    for trial in range(3):
        x = (rng.normal(size=2048) + 1j * rng.normal(size=2048)).astype(np.complex64)
        if trial == 2:
            x[:] = 0
            x[5] = 1.0  # impulse: exposes any index permutation error exactly
        out = np.zeros(2048, np.complex64)
        getattr(emul, fn)(x.ctypes.data_as(ctypes.c_void_p), tw.ctypes.data_as(ctypes.c_void_p),
                          out.ctypes.data_as(ctypes.c_void_p), reverse)
        ref = np.fft.fft(x.astype(np.complex128))
        assert np.abs(out - ref).max() <= 4e-7 * max(1.0, np.abs(ref).max())


def test_twiddle_table():
    tw = tables.fft_twiddles()
    assert tw.shape == (2048, 2) and tw.dtype == np.float32
    k = np.arange(2048)
    np.testing.assert_allclose(tw[:, 0] + 1j * tw[:, 1], np.exp(-2j * np.pi * k / 2048), atol=6e-8)
    assert tuple(tw[512]) == (0.0, -1.0) and tuple(tw[1024]) == (-1.0, 0.0)


@pytest.fixture(scope="module")
def remul(tmp_path_factory):
    so = tmp_path_factory.mktemp("remul") / "rfft_emul.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(so),
                    os.path.join(ROOT, "tests", "emulation", "rfft_emul.cpp")], check=True)
    return ctypes.CDLL(str(so))


@pytest.mark.parametrize("reverse", [0, 1])
def test_packed_real_fft_matches_numpy(remul, reverse):
    """csrc/rfft2048x2.cuh (two frames per warp, 32 x 32 complex transform + in-register conjugate-pair split)."""
    rng = np.random.default_rng(1)
    tw = tables.fft_twiddles()
    for trial in range(4):
        fr = rng.normal(size=(2, 2048)).astype(np.float32)
        if trial == 1:
            fr[1] = 0.0       # a silent partner stays exactly silent
        if trial == 2:
            fr[:] = 0
            fr[0, 7] = 1.0    # impulses: any index permutation error shows exactly
            fr[1, 1024] = 1.0
        if trial == 3:
            fr[0] *= 1e-4     # a quiet frame keeps its relative accuracy next to a loud one
        out = np.zeros((2, 1025), np.float32)
        rc = remul.emul_rfft2048x2(fr.ctypes.data_as(ctypes.c_void_p), tw.ctypes.data_as(ctypes.c_void_p),
                                   out.ctypes.data_as(ctypes.c_void_p), reverse)
        assert rc == 0, f"bin {rc - 1} not written exactly once"
        ref = 2.0 * np.abs(np.fft.rfft(fr.astype(np.float64), axis=1))
        for f in range(2):
            assert np.abs(out[f] - ref[f]).max() <= 4e-7 * max(np.abs(ref[f]).max(), 1e-30), (trial, f)
        if trial == 1:
            assert not out[1].any()
