"""GPU results against fixtures produced by the UNMODIFIED reference (tests/golden/*.npz): the
reference itself never travels to the GPU box, its outputs do."""
import os

import numpy as np
import pytest

from util import rel_err, rel_err_rows

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["wals_k30", "wals_k64", "wals_k128"])
def test_wals_epochs_match_reference(name):
    from qmf_b200 import WalsEngineHandle, csr_from_coo
    g = np.load(os.path.join(GOLD, name + ".npz"))
    k = int(g["k"])
    uids, urp, uci, uv = csr_from_coo(g["u"], g["i"], g["v"])
    iids, irp, ici, iv = csr_from_coo(g["i"], g["u"], g["v"])
    assert np.array_equal(uids, g["rid0"]) and np.array_equal(iids, g["rid1"])     # index order bit-exact
    assert np.array_equal(urp, g["rp0"]) and np.array_equal(uci, g["ci0"])
    h = WalsEngineHandle(len(uids), len(iids), k)
    h.set_csr(0, urp, uci, uv)
    h.set_csr(1, irp, ici, iv)
    h.set_factors(1, g["Y0"])
    for e in range(len(g["losses"])):
        lu = h.half_step(0, float(g["alpha"]), float(g["lam"]))
        li = h.half_step(1, float(g["alpha"]), float(g["lam"]))
        assert rel_err_rows(h.get_factors(0), g["X"][e]) < 1e-9
        assert rel_err_rows(h.get_factors(1), g["Y"][e]) < 1e-9
        assert abs(lu - g["losses"][e, 0]) <= 1e-12 * abs(lu)
        assert abs(li - g["losses"][e, 1]) <= 1e-12 * abs(li)


def test_bpr_steps_match_reference():
    from qmf_b200.bpr import BprEngineHandle
    g = np.load(os.path.join(GOLD, "misc.npz"))
    P0, Q0, b0 = g["step_P0"], g["step_Q0"], g["step_b0"]
    lr, lu, li, lb = g["step_cfg"]
    h = BprEngineHandle(P0.shape[0], Q0.shape[0], P0.shape[1], use_biases=True)
    h.set_factors(0, P0); h.set_factors(1, Q0); h.set_biases(b0)
    h.update_triplets(g["step_u"], g["step_i"], g["step_j"], lr, lu, li, lb)
    assert rel_err_rows(h.get_factors(0), g["step_P"]) < 1e-13
    assert rel_err_rows(h.get_factors(1), g["step_Q"]) < 1e-13
    assert rel_err(h.get_biases(), g["step_b"]) < 1e-13


@pytest.mark.parametrize("name", ["bpr_k30", "bpr_k16_nobias"])
def test_bpr_eval_of_reference_factors(name):
    """eval losses (1e-12) and ranking metrics (EXACT) of the reference's own final factors"""
    from qmf_b200.bpr import BprEngineHandle
    from qmf_b200.evalrank import average_metric, eval_rank, labels_to_csr, user_metrics
    g = np.load(os.path.join(GOLD, name + ".npz"))
    cfg = dict(zip(g["cfg_keys"], g["cfg_vals"]))
    k, biases, nth = int(cfg["k"]), bool(cfg["biases"]), int(cfg["nthreads"])
    P, Q = g["P"], g["Q"]
    b = g["b"] if biases else None
    h = BprEngineHandle(P.shape[0], Q.shape[0], k, use_biases=biases)
    h.set_factors(0, P); h.set_factors(1, Q)
    if biases:
        h.set_biases(b)
    for test in (0, 1):
        got = h.eval_loss(g["eu%d" % test], g["ei%d" % test], g["ej%d" % test], nth)
        assert abs(got - g["losses"][-1, test]) <= 1e-12 * abs(got)
    rows = [np.flatnonzero(g["labels"][t] > 0).astype(np.int32) for t in range(len(g["test_users"]))]
    lp, li = labels_to_csr(rows)
    cnt, _ = eval_rank(P, Q, b, g["test_users"], lp, li)
    names = [str(x) for x in g["metric_names"]]
    per_user = [user_metrics(cnt[lp[t] + t: lp[t + 1] + t + 1], Q.shape[0], names) for t in range(len(rows))]
    for name, want in zip(names, g["metric_values"]):
        assert average_metric([m[name] for m in per_user], 4) == want, name
