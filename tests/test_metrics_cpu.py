"""Host side of the ranking evaluation (no GPU): qmfb_repeated_add must give the exact result of the
reference's one-addition-per-negative AUC loop (qmf/metrics/Metrics.cpp:87-95), and qmfb_rank_metrics the
reference's metric values from bucket counts."""
import numpy as np
import pytest


def _literal(s, t, c):
    for _ in range(c):
        s = s + t
    return s


def test_repeated_add_equals_the_literal_loop():
    from qmf_b200 import capi
    f = capi.lib.qmfb_repeated_add
    rng = np.random.default_rng(0)
    cases = []
    for _ in range(400):
        pos, neg = int(rng.integers(1, 60)), int(rng.integers(1, 5000))
        tp = int(rng.integers(1, pos + 1))
        cases.append((float(rng.random() * rng.choice([0.0, 1e-6, 0.3, 1.0])), tp / pos / neg, int(rng.integers(0, 20000))))
    # ties (t an odd multiple of half an ulp of s), binade crossings, absorbed increments, tiny accumulators
    cases += [(1.0, 2.0 ** -53, 5000), (1.0 + 2.0 ** -52, 2.0 ** -53, 5000), (1.0, 3 * 2.0 ** -53, 7000),
              (1.0 + 2.0 ** -52, 3 * 2.0 ** -53, 7001), (0.5, 2.0 ** -54 + 2.0 ** -53, 100000), (1.0, 2.0 ** -54, 1000),
              (0.0, 1e-9, 100000), (0.0, 2.0 ** -40, 1 << 20), (0.99999, 1e-7, 3000), (0.0, 1.0 / 3 / 7, 12345),
              (2.0 ** -30, 2.0 ** -31 + 2.0 ** -83, 50000), (0.75, 5 * 2.0 ** -54, 40000), (0.0, 0.1, 10)]
    for s, t, c in cases:
        assert f(s, t, c) == _literal(s, t, c), (s, t, c)


def test_repeated_add_long_runs_are_fast_and_exact():
    """10^9 additions in microseconds: checked against the closed form while the sum stays exact"""
    from qmf_b200 import capi
    f = capi.lib.qmfb_repeated_add
    assert f(0.0, 2.0 ** -40, 10 ** 9) == 10 ** 9 * 2.0 ** -40        # every partial sum is representable
    # composition: n + m additions == n then m additions
    t = 1.0 / 7 / 999983
    assert f(f(0.1, t, 123456789), t, 987654321) == f(0.1, t, 123456789 + 987654321)


@pytest.mark.parametrize("name", ["auc", "ap", "p@10", "r@3"])
def test_rank_metrics_match_the_oracle_on_dense_vectors(oracle_lib, name):
    import oracle
    from qmf_b200.evalrank import rank_metrics
    rng = np.random.default_rng(5)
    ni, nT = 300, 40
    kind, at = oracle.metric_kind(name)
    lp, cnt, want = [0], [], []
    for t in range(nT):
        scores = np.round(rng.normal(size=ni), 1)                  # rounded: many ties
        labels = (rng.random(ni) < 0.05).astype(np.float64)
        labels[rng.integers(0, ni)] = 1.0
        sp = np.sort(scores[labels > 0])
        c = np.zeros(len(sp) + 1, np.int32)
        for x in np.flatnonzero(labels <= 0):
            c[np.searchsorted(sp, scores[x], side="left")] += 1
        cnt.append(c)
        lp.append(lp[-1] + len(sp))
        want.append(oracle_lib.qmfo_metric_one(kind, at, np.ascontiguousarray(labels), np.ascontiguousarray(scores), ni))
    got = rank_metrics(name, np.concatenate(cnt), np.array(lp), ni)
    assert np.array_equal(got, np.array(want))
