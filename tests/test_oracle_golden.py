"""CPU: pin the oracle (oracle/qmc_oracle.py) to outputs of the reference itself.

The vectors under tests/golden/ were produced by tests/golden/make_golden.py, which imports
and executes /root/reference/qmc/*.py.  The reference ships no tests of its own for this
path, so these are the known answers."""
import numpy as np
import pytest
import torch

from conftest import all_case_tags, load_golden, lsq_case_inputs, lsq_case_names, nll_case_inputs
from oracle import qmc_oracle as oc

torch.set_num_threads(1)


def test_get_tensor_matches_reference(fixture_instance):
    g = fixture_instance["golden"]
    S = fixture_instance["S_true"].unsqueeze(1)
    C = fixture_instance["C_true"]
    X = oc.get_tensor(S, C)
    assert X.shape == (64, 51, 51) and X.dtype == torch.float32
    np.testing.assert_array_equal(X[::4, ::3, ::3].numpy(), g["get_tensor_sub"])
    assert abs(X.double().sum().item() - float(g["get_tensor_sum"])) == 0.0
    # the vectorised statement performs the same fp32 operations in the same order
    np.testing.assert_array_equal(oc.get_tensor_vectorised(S, C).numpy(), X.numpy())
    # T_true of the shipped instance is exactly sum_r S_r o c_r
    assert np.abs(X[::4, ::3, ::3].numpy() - g["T_true_sub"]).max() <= 1e-8
    T_true = X  # stands in for T_true (max abs deviation stored in the golden file is 7.5e-9)
    n = oc.NMSE(oc.get_tensor(0.7 * S, C), T_true).item()
    assert abs(n - float(g["nmse_07"])) < 1e-6


@pytest.mark.parametrize("table", [k[3:] for k in load_golden("quantize.npz").files if k.startswith("x__")])
def test_assign_levels_bit_exact(table):
    q = load_golden("quantize.npz")
    bb = torch.from_numpy(load_golden("tables.npz")[table])
    x = torch.from_numpy(q[f"x__{table}"])
    want = q[f"y__{table}"].astype(np.int64)
    got = oc.assign_levels(x, bb).numpy()
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(oc.assign_levels_closed_form(x.numpy(), bb.numpy()), want)
    # the oracle must not mutate the caller's table (reference clones, quantization_model.py:15)
    assert bb[-1].item() != float("inf")


def test_seeded_quantize_draws_like_the_reference(fixture_instance):
    q = load_golden("quantize.npz")
    t = load_golden("tables.npz")
    T_true = oc.get_tensor(fixture_instance["S_true"].unsqueeze(1), fixture_instance["C_true"])
    # T_true differs from the .mat's stored T_true by <= 7.5e-9, which can flip a level on a
    # boundary tie; allow a handful of such entries out of 166,464.
    torch.manual_seed(int(q["seeded_lin_seed"]))
    y = oc.quantize(T_true, float(q["seeded_lin_std"]), torch.tensor([0.0, 5e-4, 1.0]))
    assert (y.numpy() != q["seeded_lin_y"]).sum() <= 8
    torch.manual_seed(int(q["seeded_log_seed"]))
    y = oc.quantize(T_true, float(q["seeded_log_std"]),
                    torch.from_numpy(t["QUANTIZATION_BOUNDARIES_7_ADJUSTED"]),
                    offset=float(t["LOG_OFFSET_7_ADJUSTED"]))
    assert (y.numpy() != q["seeded_log_y"]).sum() <= 8


@pytest.mark.parametrize("tag", all_case_tags())
def test_nll_and_grads_match_reference(tag, nll_golden, fixture_instance):
    c = nll_case_inputs(nll_golden, fixture_instance, tag)
    nll, gS, gC = oc.nll_and_grads(c["S"], c["C"], c["Y"], c["Wx"], c["bb"], c["sigma"],
                                   offset=c["offset"], sentinels=c["sentinels"])
    if np.isnan(c["nll"]):
        # the reference itself is NaN here (an entry with P == 0, mask applied by multiplication)
        assert torch.isnan(nll)
        return
    # same ops, same order, single thread: the fp32 port reproduces the reference bit for bit
    assert nll.item() == pytest.approx(c["nll"], rel=1e-6)
    np.testing.assert_allclose(gS[:, 0].numpy(), c["gS"].numpy(), rtol=1e-5, atol=1e-5 * float(c["gS"].abs().max()) + 1e-30)
    np.testing.assert_allclose(gC.numpy(), c["gC"].numpy(), rtol=1e-5, atol=1e-5 * float(c["gC"].abs().max()) + 1e-30)
    # the vectorised ("fair CPU") statement agrees too
    nll_v, gS_v, gC_v = oc.nll_and_grads(c["S"], c["C"], c["Y"], c["Wx"], c["bb"], c["sigma"],
                                         offset=c["offset"], sentinels=c["sentinels"], vectorised=True)
    assert nll_v.item() == pytest.approx(c["nll"], rel=1e-6)


@pytest.mark.parametrize("tag", all_case_tags())
def test_fp64_oracle_brackets_the_reference(tag, nll_golden, fixture_instance):
    """Where the fp32 reference is accurate (P >= 1e-5 on every entry) the independent float64
    statement agrees with it to ~1e-6; elsewhere it stays finite while the reference does not."""
    c = nll_case_inputs(nll_golden, fixture_instance, tag)
    nll64, gS64, gC64, pmin = oc.nll_and_grads_fp64(c["S"], c["C"], c["Y"], c["Wx"], c["bb"], c["sigma"],
                                                    offset=c["offset"], sentinels=c["sentinels"])
    assert np.isfinite(nll64)
    if np.isnan(c["nll"]) or c["Pmin_all"] < 1e-5:
        return
    assert nll64 == pytest.approx(c["nll"], rel=2e-6)
    gs = c["gS"].double().numpy()
    gc = c["gC"].double().numpy()
    # gradient tolerance = the north star's 1e-4: the reference's own fp32 rounding already
    # costs ~1.2e-5 in gC on the 8-level uniform table (heavy cancellation between entries)
    if np.linalg.norm(gs) > 0:
        assert np.linalg.norm(gS64[:, 0] - gs) / np.linalg.norm(gs) < 1e-4
        assert np.linalg.norm(gC64 - gc) / np.linalg.norm(gc) < 1e-4


def test_small_fry_matches_reference(fixture_instance):
    m = load_golden("misc.npz")
    t = load_golden("tables.npz")
    S = fixture_instance["S_true"].unsqueeze(1)
    C = fixture_instance["C_true"]
    T_true = oc.get_tensor(S, C)
    T_s = 0.8 * T_true
    target = (T_true > 5e-4).float()
    assert oc.neg_likelihood_bce(T_s, target, 5e-4, 0.008).item() == pytest.approx(float(m["bce_probit"]), rel=1e-5)
    assert oc.neg_likelihood_bce(T_s, target, 5e-4, probit=False).item() == pytest.approx(float(m["bce_sigmoid"]), rel=1e-5)
    assert oc.neg_likelihood_bce(T_s, target, 5e-4, 1e-4).item() == pytest.approx(float(m["bce_probit_tail"]), rel=1e-4)
    np.testing.assert_allclose(oc.F_sigmoid(torch.from_numpy(m["F_sigmoid_x"])).numpy(), m["F_sigmoid_y"], rtol=1e-6)
    np.testing.assert_allclose(oc.F_probit(torch.from_numpy(m["F_probit_x"]), 0.008).numpy(), m["F_probit_y"], rtol=1e-6, atol=1e-7)
    bb7 = torch.from_numpy(t["QUANTIZATION_BOUNDARIES_7_ADJUSTED"])
    np.testing.assert_array_equal(oc.get_quantized_obs_from_ordinal(torch.arange(7), bb7).numpy(), m["midpoints"])
    assert oc.deterministic_cost(0.8 * S, C, 2 * target - 1, mean=5e-4).item() == pytest.approx(float(m["determ_cost"]), rel=1e-5)
    np.testing.assert_array_equal(oc.outer_band_loop(S[0, 0], C[0])[::8, ::5, ::5].numpy(), m["outer_sub"])


@pytest.mark.parametrize("name", lsq_case_names())
def test_least_squares_baseline_matches_reference(name, fixture_instance):
    """oracle.masked_lsq against the reference's own run of qmc_dowjons.ipynb c1:84,108-114
    (same fp32 op sequence -> the same numbers to rounding of the reduction order)."""
    c = lsq_case_inputs(fixture_instance, name)
    obs = oc.get_quantized_obs_from_ordinal(c["Y"], c["bb"])
    np.testing.assert_array_equal(obs.numpy()[::8, :, ::5, ::5], c["obs_sub"])
    cost, gS, gC = oc.lsq_and_grads(c["S"], c["C"], c["Y"], c["Wx"], c["bb"], c["offset"])
    assert abs(cost.item() - c["cost"]) <= 1e-6 * abs(c["cost"])
    np.testing.assert_allclose(gS.numpy(), c["gS"], rtol=1e-5, atol=1e-6 * np.abs(c["gS"]).max())
    np.testing.assert_allclose(gC.numpy(), c["gC"], rtol=1e-5, atol=1e-6 * np.abs(c["gC"]).max())
    # the vectorised statement gives the same cost
    cost_v = oc.masked_lsq(c["S"], c["C"], c["Y"], c["Wx"], c["bb"], c["offset"], vectorised=True)
    assert abs(cost_v.item() - c["cost"]) <= 1e-6 * abs(c["cost"])


@pytest.mark.parametrize("sentinels", [False, True])
def test_logistic_model_fp64_statement_matches_the_composition(sentinels):
    """prob_sigmoid = the body of prob_probit with the reference's F_sigmoid; its tail-stable float64 statement
    (what the CUDA logistic epilogue is checked against) agrees with autograd through the literal composition.
    With the +-1e5 sentinels the literal form overflows exp() in the backward pass (NaN), like the probit
    tails: there only the NLL is compared."""
    g = torch.Generator().manual_seed(5)
    R, K, I, J = 3, 6, 5, 4
    S = (torch.rand(R, 1, I, J, generator=g) * 0.1 + 0.01).double()
    C = (torch.rand(R, K, generator=g) * 0.2 + 0.02).double()
    T = oc.get_tensor_vectorised(S.float(), C.float())
    bb = torch.linspace(T.min().item(), T.max().item(), 5)
    Y = oc.assign_levels(T + 0.002 * torch.randn(T.shape, generator=g), bb).reshape(K, 1, I, J)
    Wx = torch.bernoulli(torch.full((K, 1, I, J), 0.5), generator=g)
    scale = float(np.float32(0.003))
    Sd, Cd = S.clone().requires_grad_(True), C.clone().requires_grad_(True)
    Th = torch.einsum("rij,rk->kij", Sd[:, 0], Cd).unsqueeze(1)
    nll = -(Wx.double() * torch.log(oc.prob_sigmoid(Y, Th, bb.double(), scale, sentinels=sentinels))).sum()
    nll.backward()
    want = oc.logistic_nll_and_grads_fp64(S, C, Y, Wx, bb, scale, None, sentinels)
    assert nll.item() == pytest.approx(want[0], rel=1e-12)
    if not sentinels:
        np.testing.assert_allclose(Sd.grad.numpy(), want[1], rtol=1e-9, atol=1e-12 * np.abs(want[1]).max())
        np.testing.assert_allclose(Cd.grad.numpy(), want[2], rtol=1e-9, atol=1e-12 * np.abs(want[2]).max())
    else:
        assert np.isfinite(want[1]).all() and np.isfinite(want[2]).all()
