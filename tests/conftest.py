import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden():
    import torch
    g = torch.load(os.path.join(ROOT, "tests", "golden", "golden_v1.pt"), weights_only=False)
    # golden_v2: the script's other assemblies (text branch, averaged fusion, base heads, MultimodalModel,
    # AudioTextualModel, class-weighted CE) — `python oracle/make_golden.py v2`
    g2 = torch.load(os.path.join(ROOT, "tests", "golden", "golden_v2.pt"), weights_only=False)
    assert not set(g["cases"]) & set(g2["cases"])
    g["cases"].update(g2["cases"])
    g["module_trees"] = g2["module_trees"]
    # golden_v3: a batch whose rows lack different modalities — `python oracle/make_golden.py v3`
    g3 = torch.load(os.path.join(ROOT, "tests", "golden", "golden_v3.pt"), weights_only=False)
    assert not set(g["cases"]) & set(g3["cases"])
    g["cases"].update(g3["cases"])
    return g


@pytest.fixture(scope="session")
def golden_c1_epoch():
    import torch
    return torch.load(os.path.join(ROOT, "tests", "golden", "golden_c1_epoch.pt"), weights_only=False)


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The C-ABI library is built in-tree (git-ignored, shipped to the GPU box).  Build it if absent so the
    CPU suite can check that it loads and exports the header's symbols."""
    lib = os.path.join(ROOT, "multimodalaggressionrecognition_b200", "libmar.so")
    if not os.path.exists(lib):
        import __graft_entry__ as g
        g.build()
    yield


@pytest.fixture(scope="session")
def golden_alternating():
    import torch
    return torch.load(os.path.join(ROOT, "tests", "golden", "golden_alternating.pt"), weights_only=False)
