import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the reference tree at /root/reference (build container only)")


@pytest.fixture(scope="session")
def quad_golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "quadrotor_golden.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def veltab_golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "velocity_tables_golden.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return 0


@pytest.fixture(scope="session")
def maze_golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "maze_golden.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def geom_golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "maze_geom_golden.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def cont_golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "maze_continuous_golden.npz"), allow_pickle=False)
