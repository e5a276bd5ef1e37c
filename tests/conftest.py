import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run under gpurun)")


@pytest.fixture(scope="session")
def lib_built():
    from bbocr_b200 import _lib
    _lib.build()
    return _lib


@pytest.fixture(scope="session")
def handle(lib_built):
    return lib_built.Handle(0)


@pytest.fixture(scope="session")
def states():
    from bbocr_b200 import weights
    return weights.calibrated_craft_state(), weights.calibrated_crnn_state()


@pytest.fixture(scope="session")
def oracle_reader(states):
    from bbocr_b200 import weights
    from oracle import easyocr_restated as E
    craft = E.CRAFT()
    craft.load_state_dict(weights.to_torch_state(states[0]))
    crnn = E.CRNN()
    crnn.load_state_dict(weights.to_torch_state(states[1]))
    return E.Reader(craft, crnn)


@pytest.fixture(scope="session")
def gpu_reader(lib_built, states):
    import bbocr_b200
    return bbocr_b200.Reader(["en"], gpu=True, verbose=False, precision="fp32", craft_state=states[0], crnn_state=states[1])
