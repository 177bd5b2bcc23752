"""Shared fixtures.  GPU tests are marked ``@pytest.mark.gpu``; everything else runs on CPU."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def fixture_instance():
    """The reference's shipped instance (qmc/onebitdata1.mat) in the notebook's torch layout."""
    g = load_golden("fixture.npz")
    S = torch.from_numpy(g["S_true"])          # [R, I, J]
    C = torch.from_numpy(g["C_true"])          # [R, K]
    return {"S_true": S, "C_true": C, "golden": g}


@pytest.fixture(scope="session")
def nll_golden():
    g = load_golden("nll_cases.npz")
    K, I, J = 64, 51, 51
    bits = np.unpackbits(g["mask_bits"])[: K * I * J]
    Wx = torch.from_numpy(bits.astype(np.float32)).reshape(K, 1, I, J)
    return {"g": g, "Wx": Wx}


POINTS = {"p07": (0.7, 1.0), "p08": (0.8, 1.0), "zero": (0.0, 0.0), "p09c11": (0.9, 1.1)}


def nll_case_inputs(nll_golden, fixture_instance, tag):
    """Rebuild the inputs of one golden likelihood case ``<case>__<point>``."""
    g = nll_golden["g"]
    case, point = tag.split("__")
    fs, fc = POINTS[point]
    S = (fs * fixture_instance["S_true"]).unsqueeze(1).contiguous()
    C = (fc * fixture_instance["C_true"]).contiguous()
    Y = torch.from_numpy(g[f"{case}__Y"].astype(np.int64)).unsqueeze(1)
    bb = torch.from_numpy(g[f"{case}__bb"])
    sigma = float(g[f"{case}__sigma"])
    off = float(g[f"{case}__offset"])
    offset = None if np.isnan(off) else off
    return dict(S=S, C=C, Y=Y, Wx=nll_golden["Wx"], bb=bb, sigma=sigma, offset=offset,
                sentinels=offset is None,
                nll=float(g[f"{tag}__nll"]), gS=torch.from_numpy(g[f"{tag}__gS"]),
                gC=torch.from_numpy(g[f"{tag}__gC"]), Pmin_obs=float(g[f"{tag}__Pmin_obs"]),
                Pmin_all=float(g[f"{tag}__Pmin_all"]))


def all_case_tags():
    g = load_golden("nll_cases.npz")
    return [str(t) for t in g["case_points"]]


def lsq_case_names():
    return [str(n) for n in load_golden("lsq_cases.npz")["names"]]


def lsq_case_inputs(fixture_instance, name):
    """One golden least-squares case (tests/golden/make_golden_lsq.py): inputs in the reference's
    shapes plus the reference's own cost and gradients."""
    g = load_golden("lsq_cases.npz")
    K, I, J = 64, 51, 51
    off = float(g[name + "_offset"])
    Wx = torch.from_numpy(np.unpackbits(g[name + "_Wx"])[: K * I * J].astype(np.float32)).reshape(K, 1, I, J)
    return {
        "S": (float(g[name + "_scale"]) * fixture_instance["S_true"]).unsqueeze(1),
        "C": fixture_instance["C_true"].clone(),
        "Y": torch.from_numpy(g[name + "_Y"].astype(np.int64)),
        "Wx": Wx,
        "bb": torch.from_numpy(g[name + "_bb"]),
        "offset": None if np.isnan(off) else off,
        "obs_sub": g[name + "_obs_sub"], "cost": float(g[name + "_cost"]), "gS": g[name + "_gS"], "gC": g[name + "_gC"],
    }
