import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """GPU tests must fail loudly (not skip) on a GPU box; on a box without a GPU they are
    deselected by `-m "not gpu"`.  If someone runs them anyway without CUDA, skip."""
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "ref_golden.npz")
    return np.load(path, allow_pickle=False)


@pytest.fixture(scope="session")
def golden_tails():
    """WideAndDeep / FNN / InnerPNN trajectories from the real reference (tests/golden/make_golden_tails.py)."""
    path = os.path.join(ROOT, "tests", "golden", "ref_golden_tails.npz")
    return np.load(path, allow_pickle=False)


def state_from_golden(golden, prefix):
    """Collect {state_dict_key: ndarray} stored under `prefix/`."""
    pre = prefix.rstrip("/") + "/"
    return {k[len(pre):]: golden[k] for k in golden.files if k.startswith(pre)}


@pytest.fixture(scope="session")
def golden_sac():
    """Hybrid SAC networks / prioritized memory from the real reference (tests/golden/make_golden_sac.py)."""
    path = os.path.join(ROOT, "tests", "golden", "ref_golden_sac.npz")
    return np.load(path, allow_pickle=False)


@pytest.fixture(scope="session")
def golden_heads():
    """Hybrid TD3 (v10) / PPO network heads from the real reference (tests/golden/make_golden_heads.py)."""
    path = os.path.join(ROOT, "tests", "golden", "ref_golden_heads.npz")
    return np.load(path, allow_pickle=False)


@pytest.fixture(scope="session")
def golden_gp10():
    """generate_preds of hybrid_td3_main_per_v10.py from the real reference (tests/golden/make_golden_gp10.py)."""
    path = os.path.join(ROOT, "tests", "golden", "ref_golden_gp10.npz")
    return np.load(path, allow_pickle=False)


@pytest.fixture(scope="session")
def golden_td3():
    """Two learn steps of the reference's hybrid TD3 (v10) agent (tests/golden/make_golden_td3.py)."""
    path = os.path.join(ROOT, "tests", "golden", "ref_golden_td3.npz")
    return np.load(path, allow_pickle=False)


@pytest.fixture(scope="session")
def golden_ppo():
    """One learn() call of the reference's hybrid PPO agent (tests/golden/make_golden_ppo.py)."""
    path = os.path.join(ROOT, "tests", "golden", "ref_golden_ppo.npz")
    return np.load(path, allow_pickle=False)
