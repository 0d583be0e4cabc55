import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / 'tests' / 'golden'


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with `-m gpu` on the GPU box')


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN


@pytest.fixture(scope='session')
def default_hp():
    from deepcv_b200.yaml_config import benchmark_model_spec
    return benchmark_model_spec(ROOT / 'conf' / 'base' / 'parameters.yml', 'image_classifier', out_features=10)
