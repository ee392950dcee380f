import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


TUBE_PARAMS = dict(dt=1e-3, theta=0.501, extrapolation="laplace", extrapolation_sub_type="constant")
TUBE_REGIONS = [
    dict(marker=1, kind="fluid", rho=1025.0, mu=3.5e-3),
    dict(marker=2, kind="solid", rho=1000.0, mu=344827.5862068966, lambda_=3103448.2758620703),
]


@pytest.fixture
def tube_params():
    return dict(TUBE_PARAMS), [dict(r) for r in TUBE_REGIONS]
