import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    path = os.path.join(ROOT, "tests", "golden", "reference_golden.npz")
    with np.load(path, allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def guitar_golden():
    import numpy as np

    path = os.path.join(ROOT, "tests", "golden", "guitar_golden.npz")
    with np.load(path, allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def fin_golden():
    import numpy as np

    path = os.path.join(ROOT, "tests", "golden", "fin_events_golden.npz")
    with np.load(path, allow_pickle=False) as z:
        return {k: z[k] for k in z.files}

