"""pytest configuration: markers, repo root on sys.path, golden-fixture loaders."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


class Case(dict):
    __getattr__ = dict.__getitem__


def _load_reallife():
    z = np.load(os.path.join(GOLDEN, "reallife.npz"))
    meta = json.load(open(os.path.join(GOLDEN, "reallife_meta.json")))
    cases = []
    for i, m in enumerate(meta["cases"]):
        c = Case(m)
        for a in ("signal", "fftRe", "fftIm", "magnitude", "phase"):
            c[a] = z[a][i]
        cases.append(c)
    windows = []
    for i, m in enumerate(meta["windows"]):
        w = Case(m)
        w["values"] = z[f"window_{i}"]
        windows.append(w)
    return cases, windows


def _load_fixtures():
    z = np.load(os.path.join(GOLDEN, "fixtures_v0_1.npz"))
    meta = json.load(open(os.path.join(GOLDEN, "fixtures_v0_1_meta.json")))
    cases = []
    for i, m in enumerate(meta["fftCases"]):
        c = Case(m)
        c["input"] = z[f"case_{i}_input"]
        c["fftRe"] = z[f"case_{i}_fftRe"]
        c["fftIm"] = z[f"case_{i}_fftIm"]
        cases.append(c)
    windows = []
    for i, m in enumerate(meta["windows"]):
        w = Case(m)
        w["values"] = z[f"window_{i}"]
        windows.append(w)
    return cases, windows


REALLIFE_CASES, REALLIFE_WINDOWS = _load_reallife()
FIXTURE_CASES, FIXTURE_WINDOWS = _load_fixtures()


def reallife(file=None, kind=None):
    return [c for c in REALLIFE_CASES if (file is None or c.file == file) and (kind is None or c.kind == kind)]


@pytest.fixture(scope="session")
def golden_reallife():
    return REALLIFE_CASES


@pytest.fixture(scope="session")
def golden_fixtures():
    return FIXTURE_CASES
