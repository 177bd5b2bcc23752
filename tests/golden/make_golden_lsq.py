#!/usr/bin/env python
"""Mint tests/golden/lsq_cases.npz by EXECUTING THE REFERENCE ITSELF (build container only; needs
/root/reference):   python tests/golden/make_golden_lsq.py

The masked least-squares baseline of qmc/qmc_dowjons.ipynb c1:84,108-114: quantise with the reference's
``quantize``, de-quantise with its ``get_quantized_obs_from_ordinal``, form ``get_tensor`` (+ log link),
``cost = torch.norm(Wx*(T_hat-Obs))**2`` and back-propagate.  Inputs come from the committed
``fixture.npz`` (the instance the reference ships) and seeded noise / masks.  Nothing here is imported
by the product.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
for name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.path[:0] = [os.path.join(REF, "qmc"), os.path.join(REF, "deep_prior")]
import quantization_model as qm_lin          # noqa: E402
import quantization_model_log as qm_log      # noqa: E402
import utils as ref_utils                    # noqa: E402

torch.set_num_threads(1)
fx = np.load(os.path.join(HERE, "fixture.npz"))
S_true = torch.from_numpy(fx["S_true"])      # [R,I,J]
C_true = torch.from_numpy(fx["C_true"])      # [R,K]
R, I, J = S_true.shape
K = C_true.shape[1]
T_true = qm_lin.get_tensor(S_true.unsqueeze(1), C_true)

out = {}
names = []


def case(name, bb, std, offset, scale, frac, seed):
    torch.manual_seed(seed)
    bb = torch.as_tensor(bb, dtype=torch.float32)
    if offset is None:
        Y = qm_lin.quantize(T_true, std, bb)
    else:
        Y = qm_log.quantize(T_true, std, bb, offset)
    Y = Y.unsqueeze(1)
    Wx = torch.bernoulli(frac * torch.ones(K, 1, I, J))
    Obs = qm_log.get_quantized_obs_from_ordinal(Y, bb, std)
    S = (scale * S_true).unsqueeze(1).clone().requires_grad_(True)
    C = C_true.clone().requires_grad_(True)
    T_hat = qm_log.get_tensor(S, C).unsqueeze(1)
    if offset is not None:
        T_hat = torch.log(T_hat + offset)
    cost = torch.norm(Wx * (T_hat - Obs)) ** 2
    cost.backward()
    names.append(name)
    out[name + "_bb"] = bb.numpy()
    out[name + "_offset"] = np.float64(np.nan if offset is None else float(offset))
    out[name + "_scale"] = np.float64(scale)
    out[name + "_Y"] = Y.numpy().astype(np.uint8)
    out[name + "_Wx"] = np.packbits(Wx.numpy().astype(np.uint8).reshape(-1))
    out[name + "_obs_sub"] = Obs.numpy().astype(np.float32)[::8, :, ::5, ::5]
    out[name + "_cost"] = np.float64(cost.item())
    out[name + "_gS"] = S.grad.numpy().astype(np.float32)
    out[name + "_gC"] = C.grad.numpy().astype(np.float32)
    print(name, "levels", int(Y.max()) + 1, "nobs", int(Wx.sum()), "cost", cost.item())


# log domain, the 7-level table and offset the dowjons notebook family uses (qmc/utils.py:43,50)
case("log7", ref_utils.QUANTIZATION_BOUNDARIES_7_ADJUSTED, 0.5, ref_utils.LOG_OFFSET_7_ADJUSTED, 0.8, 0.1, 1)
case("log7_dense", ref_utils.QUANTIZATION_BOUNDARIES_7_ADJUSTED, 3.0, ref_utils.LOG_OFFSET_7_ADJUSTED, 1.1, 0.5, 2)
# linear domain, 8 uniform levels over the data range (qmc/utils.py:18-19 pattern)
tmax = float(T_true.max())
case("lin8", torch.arange(9, dtype=torch.float32) * tmax / 8, 2 * tmax / 8, None, 0.7, 0.2, 3)
out["names"] = np.array(names)
np.savez_compressed(os.path.join(HERE, "lsq_cases.npz"), **out)
print("lsq_cases.npz", os.path.getsize(os.path.join(HERE, "lsq_cases.npz")) / 1024, "KiB")
