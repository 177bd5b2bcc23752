#!/usr/bin/env python
"""Mint the golden vectors under tests/golden/ by EXECUTING THE REFERENCE ITSELF.

Run in the build container only (it needs /root/reference, which does not exist on the GPU
box):   python tests/golden/make_golden.py

It imports /root/reference/qmc/quantization_model.py, quantization_model_log.py and utils.py
verbatim (matplotlib is stubbed because qmc/utils.py imports it at module scope; nothing we
call touches it), runs them on the only data the reference ships (qmc/onebitdata1.mat) and on
seeded inputs, and stores inputs and outputs as small .npz files.  The tests then check
(a) the oracle in oracle/qmc_oracle.py against these vectors on the CPU and (b) the CUDA path
against them on the GPU.  Nothing here is imported by the product.
"""
import os
import sys
import types

import numpy as np
import scipy.io as sio
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

for name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.path[:0] = [os.path.join(REF, "qmc"), os.path.join(REF, "deep_prior")]
# slf_dataset imports pandas/matplotlib only; fine with the stub
import quantization_model as qm_lin          # noqa: E402
import quantization_model_log as qm_log      # noqa: E402
import utils as ref_utils                    # noqa: E402

torch.set_num_threads(1)   # deterministic reduction order for the stored scalars


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path)/1024:.1f} KiB, {len(arrays)} arrays")


def f32(t):
    return np.ascontiguousarray(t.detach().cpu().numpy().astype(np.float32))


# ---------------------------------------------------------------------------------------
# 1. the shipped instance (qmc/onebitdata1.mat) in the notebook's torch layout
#    (qmc.ipynb c1:75-80: [I,J,K]->[K,I,J], [I,J,R]->[R,I,J], [K,R]->[R,K])
# ---------------------------------------------------------------------------------------
mat = sio.loadmat(os.path.join(REF, "qmc", "onebitdata1.mat"))
S_true = torch.from_numpy(mat["S_true"]).type(torch.float32).permute(2, 0, 1).contiguous()
C_true = torch.from_numpy(mat["C_true"]).type(torch.float32).permute(1, 0).contiguous()
T_true = torch.from_numpy(mat["T_true"]).type(torch.float32).permute(2, 0, 1).contiguous()
R, I, J = S_true.shape
K = C_true.shape[1]
T_ref = qm_lin.get_tensor(S_true.unsqueeze(1), C_true)
save("fixture.npz",
     S_true=f32(S_true), C_true=f32(C_true),
     T_true_sub=f32(T_true[::4, ::3, ::3]),
     get_tensor_sub=f32(T_ref[::4, ::3, ::3]),
     get_tensor_sum=np.float64(T_ref.double().sum().item()),
     get_tensor_abs_dev_from_T_true=np.float64((T_ref - T_true).abs().max().item()),
     nmse_07=np.float32(qm_lin.NMSE(qm_lin.get_tensor(0.7 * S_true.unsqueeze(1), C_true), T_true).item()),
     nmse_log_07=np.float32(qm_log.NMSE_LOG(qm_log.get_tensor(0.7 * S_true.unsqueeze(1), C_true), T_true,
                                            ref_utils.LOG_OFFSET_7_ADJUSTED).item()))

# ---------------------------------------------------------------------------------------
# 2. boundary tables of qmc/utils.py:11-51 (as the float32 tensors the notebooks build)
# ---------------------------------------------------------------------------------------
TABLES = {}
for key in dir(ref_utils):
    if key.startswith("QUANTIZATION_BOUNDARIES_"):
        TABLES[key] = torch.as_tensor(getattr(ref_utils, key), dtype=torch.float32)
consts = {k: np.float64(float(getattr(ref_utils, k))) for k in dir(ref_utils)
          if k.startswith(("SD_", "LOG_OFFSET_"))}
save("tables.npz", **{k: f32(v) for k, v in TABLES.items()}, **consts)

# ---------------------------------------------------------------------------------------
# 3. quantizer known answers: reference quantize() with noise_std = 0 evaluates the level
#    assignment on exactly the given input (X + randn*0 == X for every X incl. inf/NaN)
# ---------------------------------------------------------------------------------------
qz = {}
g = torch.Generator().manual_seed(7)
for key, bb in TABLES.items():
    n = bb.numel()
    lo, hi = bb[0].item(), bb[-1].item()
    span = hi - lo
    edges = [bb, torch.nextafter(bb, torch.tensor(float("inf"))),
             torch.nextafter(bb, torch.tensor(-float("inf"))),
             torch.tensor([float("nan"), float("inf"), -float("inf"), lo - span, hi + span, 0.0, -0.0])]
    rnd = lo - 0.1 * span + 1.2 * span * torch.rand(4096, generator=g)
    x = torch.cat(edges + [rnd]).to(torch.float32)
    y = qm_lin.quantize(x, 0.0, bb)
    assert y.max().item() <= n - 2
    qz[f"x__{key}"] = f32(x)
    qz[f"y__{key}"] = y.numpy().astype(np.int16)
# seeded end-to-end draws (noise drawn by the reference from torch's global CPU generator)
bb2 = torch.tensor([0.0, 5e-4, 1.0])
torch.manual_seed(11)
qz["seeded_lin_y"] = qm_lin.quantize(T_true, 1e-3, bb2).numpy().astype(np.uint8)
bb7 = TABLES["QUANTIZATION_BOUNDARIES_7_ADJUSTED"]
torch.manual_seed(12)
qz["seeded_log_y"] = qm_log.quantize(T_true, 0.5, bb7, offset=ref_utils.LOG_OFFSET_7_ADJUSTED).numpy().astype(np.uint8)
qz["seeded_lin_seed"] = np.int64(11)
qz["seeded_log_seed"] = np.int64(12)
qz["seeded_lin_std"] = np.float64(1e-3)
qz["seeded_log_std"] = np.float64(0.5)
save("quantize.npz", **qz)

# ---------------------------------------------------------------------------------------
# 4. likelihood / gradient known answers on the shipped instance
#    mask: per-entry Bernoulli(0.1) as qmc.ipynb c1:70-72, seed 1
# ---------------------------------------------------------------------------------------
torch.manual_seed(1)
Wx = torch.bernoulli(torch.ones((K, 1, I, J)) * 0.1)
noise = torch.randn(T_true.shape)
# (noise is not stored: only the levels Y derived from it are needed downstream)
nl = {"mask_bits": np.packbits(Wx.numpy().astype(np.uint8).reshape(-1)),
      "mask_seed": np.int64(1)}

bb8u = TABLES["QUANTIZATION_BOUNDARIES_8_BINS_UNIFORM"]
bb16 = TABLES["QUANTIZATION_BOUNDARIES_16_ADJUSTED"]
CASES = [
    # name, module, bb, sigma, offset
    ("lin2_s8e-3", qm_lin, bb2, 0.008, None),
    ("lin2_s1e-3", qm_lin, bb2, 1e-3, None),
    ("lin2_s1e-4", qm_lin, bb2, 1e-4, None),          # documented divergence regime at zero start
    ("lin8u_s2bw", qm_lin, bb8u, float(bb8u[1] - bb8u[0]) * 2, None),
    ("log7_s5", qm_log, bb7, 5.0, ref_utils.LOG_OFFSET_7_ADJUSTED),
    ("log7_s3", qm_log, bb7, 3.0, ref_utils.LOG_OFFSET_7_ADJUSTED),
    ("log7_s0.5", qm_log, bb7, 0.5, ref_utils.LOG_OFFSET_7_ADJUSTED),
    ("log7_s0.05", qm_log, bb7, 0.05, ref_utils.LOG_OFFSET_7_ADJUSTED),
    ("log16_s1", qm_log, bb16, 1.0, ref_utils.LOG_OFFSET_16_ADJUSTED),
]
POINTS = {"p07": (0.7, 1.0), "p08": (0.8, 1.0), "zero": (0.0, 0.0), "p09c11": (0.9, 1.1)}
names = []
for name, mod, bb, sigma, offset in CASES:
    # observation: the reference quantizer on the shared noise (noise_std applied by hand so
    # that both back ends see identical noisy values)
    if offset is None:
        noisy = T_true + noise * sigma
    else:
        noisy = torch.log(T_true + offset) + noise * sigma
    # level assignment on the given noisy values: quantize(.., 0.0, ..) of the linear file
    # (the loop is identical in quantization_model_log.py:15-21, which would re-apply the log)
    Y = qm_lin.quantize(noisy, 0.0, bb)
    Y4 = Y.unsqueeze(1)
    nl[f"{name}__bb"] = f32(bb)
    nl[f"{name}__sigma"] = np.float64(sigma)
    nl[f"{name}__offset"] = np.float64(np.nan if offset is None else offset)
    nl[f"{name}__Y"] = Y.numpy().astype(np.uint8)
    for pname, (fs, fc) in POINTS.items():
        S = (fs * S_true).unsqueeze(1).clone().requires_grad_(True)
        C = (fc * C_true).clone().requires_grad_(True)
        T_hat = mod.get_tensor(S, C).unsqueeze(1)
        if offset is not None:
            T_hat = torch.log(T_hat + offset)
        P = mod.prob_probit(Y4, T_hat, bb, sigma)
        nll = -torch.sum(Wx * torch.log(P))
        nll.backward()
        tag = f"{name}__{pname}"
        nl[f"{tag}__nll"] = np.float32(nll.item())
        nl[f"{tag}__gS"] = f32(S.grad[:, 0])
        nl[f"{tag}__gC"] = f32(C.grad)
        nl[f"{tag}__Pmin_obs"] = np.float32(P.detach()[Wx != 0].min().item())
        nl[f"{tag}__Pmin_all"] = np.float32(P.detach().min().item())
        nl[f"{tag}__P_sub"] = f32(P.detach()[::8, 0, ::5, ::5])
        names.append(tag)
        print(f"{tag:28s} nll={nll.item():.6g} Pmin_obs={nl[f'{tag}__Pmin_obs']:.3g} "
              f"Pmin_all={nl[f'{tag}__Pmin_all']:.3g} |gS|={S.grad.norm().item():.4g} |gC|={C.grad.norm().item():.4g}")
nl["case_points"] = np.array(names)
save("nll_cases.npz", **nl)

# ---------------------------------------------------------------------------------------
# 5. the small fry: one-bit BCE form, logistic CDF, mid-points, deterministic cost
# ---------------------------------------------------------------------------------------
torch.manual_seed(3)
T_s = 0.8 * T_true
target = (T_true > 5e-4).float()
misc = {
    "bce_probit": np.float32(qm_lin.NegLikelihood(5e-4, std=0.008)(T_s, target).item()),
    "bce_sigmoid": np.float32(qm_lin.NegLikelihood(5e-4, probit=False)(T_s, target).item()),
    "bce_probit_tail": np.float32(qm_lin.NegLikelihood(5e-4, std=1e-4)(T_s, target).item()),
    "F_sigmoid_x": f32(torch.linspace(-30, 30, 121)),
    "F_sigmoid_y": f32(qm_lin.F_sigmoid(torch.linspace(-30, 30, 121))),
    "F_probit_x": f32(torch.linspace(-0.05, 0.05, 201)),
    "F_probit_y": f32(qm_lin.F_probit(torch.linspace(-0.05, 0.05, 201), 0.008)),
    "midpoints": f32(qm_log.get_quantized_obs_from_ordinal(torch.arange(7), bb7, 0.5)),
    "determ_cost": np.float32(qm_lin.DeterministicCost(mean=5e-4)(0.8 * S_true.unsqueeze(1), C_true,
                                                                 2 * target - 1).item()),
    "outer_sub": f32(qm_lin.outer(S_true[0], C_true[0])[::8, ::5, ::5]),
}
save("misc.npz", **misc)

# ---------------------------------------------------------------------------------------
# 6. end-to-end: 25 alternating Adam iterations of the MLE loop, driven by the reference's
#    own functions (loop shape of qmc.ipynb c1:136-157 with S optimised directly, as in
#    backup/notebooks/onebit_lowrank.ipynb c1; no generator because its weights are not
#    shipped).  Linear domain, one-bit, sigma = 8e-3 (the notebook's std_probit, c1:58; at 1e-3 the
#    reference itself turns NaN after 8 iterations because an unobserved entry reaches P == 0), 10 % per-entry mask from above.
# ---------------------------------------------------------------------------------------
name, mod, bb, sigma, offset = CASES[0]
Y4 = torch.from_numpy(nl[f"{name}__Y"].astype(np.int64)).unsqueeze(1)
S = (0.7 * S_true).unsqueeze(1).clone().requires_grad_(True)
C = (0.9 * C_true).clone().requires_grad_(True)
optC = torch.optim.Adam([C], lr=0.005)
optS = torch.optim.Adam([S], lr=0.001)
lam = 1.0
trace = []
for it in range(25):
    optC.zero_grad()
    T_hat = mod.get_tensor(S.detach().clone(), C).unsqueeze(1)
    cost = -torch.sum(Wx * torch.log(mod.prob_probit(Y4, T_hat, bb, sigma))) + lam * torch.norm(C, "fro")
    cost.backward()
    optC.step()
    with torch.no_grad():
        C[C < 0] = 0
    optS.zero_grad()
    T_hat = mod.get_tensor(S, C.detach()).unsqueeze(1)
    cost = -torch.sum(Wx * torch.log(mod.prob_probit(Y4, T_hat, bb, sigma))) + lam * torch.norm(S, "fro")
    cost.backward()
    optS.step()
    with torch.no_grad():
        S[S < 0] = 0
    trace.append((cost.item(), mod.NMSE(mod.get_tensor(S, C), T_true).item()))
    print(it, trace[-1])
save("solver.npz", case=np.array(name), iters=np.int64(25), lam=np.float64(lam),
     lrC=np.float64(0.005), lrS=np.float64(0.001), s_scale=np.float64(0.7), c_scale=np.float64(0.9),
     trace=np.array(trace, dtype=np.float64), S_final=f32(S[:, 0]), C_final=f32(C))
