import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def weights_seed0(tmp_path_factory):
    from irmv_detection_b200 import weights
    p = tmp_path_factory.mktemp("w") / "yolov8n_seed0.irmw"
    weights.write_random(str(p), seed=0)
    return str(p)


@pytest.fixture(scope="session")
def weights_seed1(tmp_path_factory):
    from irmv_detection_b200 import weights
    p = tmp_path_factory.mktemp("w") / "yolov8n_seed1.irmw"
    weights.write_random(str(p), seed=1)
    return str(p)


@pytest.fixture(scope="session")
def base_image():
    from irmv_detection_b200 import synth
    return synth.load_base()
