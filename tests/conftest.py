import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_audio():
    import numpy as np
    from oracle import meyda_oracle as mo
    z = np.load(os.path.join(ROOT, "tests", "golden", "audio_pcm16.npz"))
    return {k: mo.pcm16_to_float(z[k]) for k in z.files}


@pytest.fixture(scope="session")
def golden_features():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_features.npz"))
