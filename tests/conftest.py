import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def golden_cases():
    import helpers
    return helpers.load_golden()


@pytest.fixture(scope="session")
def golden_indexes(tmp_path_factory, golden_cases):
    """builds every golden corpus once with the product's index writer -> {case name: path prefix}"""
    import helpers
    base = tmp_path_factory.mktemp("golden_idx")
    out = {}
    for case in golden_cases:
        prefix = str(base / case["name"])
        helpers.build_golden_index(case, prefix)
        out[case["name"]] = prefix
    return out


@pytest.fixture(scope="session")
def golden_indexes_crc(tmp_path_factory, golden_cases):
    """the same golden corpora written with dict=crc (FNV64 word ids instead of keywords in .spi)"""
    import helpers
    base = tmp_path_factory.mktemp("golden_idx_crc")
    out = {}
    for case in golden_cases:
        prefix = str(base / case["name"])
        helpers.build_golden_index(case, prefix, dict_crc=True)
        out[case["name"]] = prefix
    return out
