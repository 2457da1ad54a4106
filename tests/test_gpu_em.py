"""GPU parity: stage ii (bootstrap + F redistribution) and stage iii (E-step, EM) vs the oracle."""
import numpy as np
import pytest

from colate_b200 import api, synth
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

RATE_RTOL = 1e-9   # north star: per-epoch coalescence rates within 1e-9 relative


def _block_stats(seed=1, rows=(30000, 20000)):
    sites = synth.make_sites(seed, list(rows), [2.5e8, 1.2e8][:len(rows)])
    gt = synth.make_genome(seed + 100, sites, 0.7)
    gr = synth.make_genome(seed + 200, sites, 0.7)
    return po.stage1(sites, gt, gr, seed=seed)


@pytest.mark.parametrize("bins,age", [("3,7,0.2", 0.0), ("3,7,0.1", 0.0), ("3,7,0.2", 250.0)])
def test_estep_matches_oracle(handle, bins, age):
    ep, _ = po.epochs_from_bins(bins, age, 28.0)
    ab = po.age_bins()
    rng = np.random.default_rng(0)
    for trial in range(6):
        rates = np.full(len(ep), 1 / 20000.) if trial == 0 else np.exp(rng.uniform(np.log(1e-7), np.log(1e-2), len(ep)))
        if trial == 3:
            rates[2] = 0.0
        for sh in (True, False):
            ll, num, den = handle.estep(sh, ep, rates, ab)
            for b in range(185):
                lo, no, do = po.estep(sh, ep, rates, ab[b])
                assert np.allclose(num[b], no, rtol=1e-11, atol=1e-300), (trial, sh, b)
                assert np.allclose(den[b], do, rtol=1e-10, atol=1e-9 * max(1.0, np.abs(do).max())), (trial, sh, b)
                assert ll[b] == pytest.approx(lo, rel=1e-12, abs=1e-12), (trial, sh, b)
                assert not np.isnan(num[b]).any() and (num[b] >= 0).all() and (den[b] >= 0).all()


def test_estep_reference_unit_test_closed_forms(handle):
    """The assertions of the reference's own unit test (include/test/test_aDNA.cpp:68-212): 21 epochs,
    7 constant rates 1e-7..1e-1, 92 age bins (C=5): no NaN, no negatives, and agreement with the oracle."""
    E = 21
    ypg = np.float32(28.0)
    ep = np.zeros(E)
    ep[1] = 1e3 / ypg
    log10 = np.float32(np.log(10))
    for e in range(2, E - 1):
        ep[e] = np.exp(log10 * (3.0 + 4.0 * (e - 1.0) / (E - 3.0))) / ypg
    ep[E - 1] = 1e8 / ypg
    t = np.exp(np.arange(92) / 5.0) / 10.0
    for f in range(1, 8):
        rates = np.full(E, 1e-7 * np.exp(np.log(10) * (f - 1)))
        for sh in (True, False):
            ll, num, den = handle.estep(sh, ep, rates, t)
            assert not np.isnan(num).any() and not np.isnan(den).any() and not np.isnan(ll).any()
            assert (num >= 0).all() and (den >= 0).all()
            for b in range(92):
                lo, no, do = po.estep(sh, ep, rates, t[b])
                # the reference's own tolerances here are 1e-3 (logl) and 0.1 abs-or-rel (num, denom):
                # at rates of 1e-7 denom[e] is a difference of nearly equal terms, so 1-ulp differences
                # between CUDA's and glibc's exp/log1p show up at ~1e-9 relative
                assert abs(ll[b] - lo) <= 1e-9 * max(1.0, abs(lo))
                assert np.allclose(num[b], no, rtol=1e-7, atol=1e-12)
                assert np.allclose(den[b], do, rtol=1e-6, atol=1e-6 * max(1.0, np.abs(do).max()))


@pytest.mark.parametrize("R,age", [(1, 0.0), (7, 0.0), (4, 250.0)])
def test_stage2_bit_exact(handle, R, age):
    o = _block_stats()
    nb = o["num_blocks"]
    st = api.mt_seed(5)
    w = api.draw_block_weights(st, R, nb)
    blk = np.stack([o["shared"], o["notshared"], o["shared_emp"], o["notshared_emp"]], axis=1)
    got = handle.stage2_bootstrap(w, blk, age)
    want = po.stage2(w, o, age)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("bins,age,R", [("3,7,0.2", 0.0, 3), ("3,7,0.1", 0.0, 2), ("3,7,0.2", 250.0, 2)])
def test_em_rates_and_iterations(handle, bins, age, R):
    o = _block_stats()
    nb = o["num_blocks"]
    st = api.mt_seed(5)
    w = api.draw_block_weights(st, R, nb)
    counts = po.stage2(w, o, age)
    ep, _ = po.epochs_from_bins(bins, age, 28.0)
    init = np.full(len(ep), 1 / 20000.)
    rates, iters, ll = handle.stage3_em(R, ep, init, counts)
    for r in range(R):
        ro, it, llo = po.em_run(ep, init, counts[r])
        assert iters[r] == it
        assert np.allclose(rates[r], ro, rtol=RATE_RTOL, atol=0), np.max(np.abs(rates[r] / ro - 1))
        assert ll[r] == pytest.approx(llo, rel=1e-10)


def test_em_short_run_tight(handle):
    """Few iterations: the per-iteration difference to glibc is at the 1e-13 level."""
    o = _block_stats()
    counts = po.stage2(np.ones((1, o["num_blocks"]), np.int32), o, 0.0)
    ep, _ = po.epochs_from_bins("3,7,0.2", 0.0, 28.0)
    init = np.full(len(ep), 1 / 20000.)
    rates, iters, ll = handle.stage3_em(1, ep, init, counts, max_iter=5)
    ro, it, llo = po.em_run(ep, init, counts[0], max_iter=5)
    assert iters[0] == it == 5
    assert np.allclose(rates[0], ro, rtol=1e-12, atol=0)
