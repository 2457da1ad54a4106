"""GPU parity: stage ii (bootstrap + F redistribution) and stage iii (E-step, EM) vs the oracle."""
import numpy as np
import pytest

from colate_b200 import api, synth
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

RATE_RTOL = 1e-9   # north star: per-epoch coalescence rates within 1e-9 relative


def _same(a, b):
    """bit-identical (NaNs compare equal)"""
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def test_device_libm_is_glibc(handle):
    """exp / log / log1p on the device follow glibc 2.39's FMA variants bit for bit."""
    assert api.libm_exact(), "host libm is not glibc 2.39/FMA: EM parity is then only approximate on this host"
    rng = np.random.default_rng(3)
    n = 1_000_000
    xs = {"exp": np.concatenate([-rng.random(n) * 800, (rng.random(n) - 0.5) * 1500, (rng.random(n) - 0.5) * 2e-15,
                                 rng.integers(0, 2**63, n, dtype=np.int64).view(np.float64), [0.0, -0.0, np.inf, -np.inf, np.nan, 709.78, -745.2]]),
          "log": np.concatenate([np.exp((rng.random(n) - 0.5) * 1400), 0.9 + rng.random(n) * 0.2, rng.random(n) * 1e-310,
                                 rng.integers(0, 2**63, n, dtype=np.int64).view(np.float64), [0.0, -0.0, 1.0, -1.0, np.inf, np.nan]]),
          "log1p": np.concatenate([-np.exp(-rng.random(n) * 760), np.exp(-rng.random(n) * 760), (rng.random(n) - 0.5) * 2,
                                   rng.random(n) * 1e6 - 0.999999, rng.integers(0, 2**63, n, dtype=np.int64).view(np.float64),
                                   [0.0, -0.0, -1.0, -2.0, np.inf, np.nan, 2.0**-29, 2.0**-54, 0.41421, 0.41422, -0.2929, -0.29289]])}
    for f, x in xs.items():
        got = handle.libm(f, x)
        want = po.libm(f, x)
        bad = ~((got.view(np.uint64) == want.view(np.uint64)) | (np.isnan(got) & np.isnan(want)))
        assert not bad.any(), (f, x[bad][:5], got[bad][:5], want[bad][:5])


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
                assert _same(num[b], no) and _same(den[b], do) and _same(ll[b], lo), (trial, sh, b)
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
                # (the reference's own tolerances here are 1e-3 on logl and 0.1 abs-or-rel on num/denom)
                assert _same(ll[b], lo) and _same(num[b], no) and _same(den[b], do), (f, sh, b)


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


@pytest.mark.parametrize("bins,age,R", [("3,7,0.2", 0.0, 3), ("3,7,0.1", 0.0, 2), ("3,7,0.2", 250.0, 2), ("3,7,0.05", 0.0, 1), ("2,8,0.3", 100.0, 5), ("4,6,0.5", 0.0, 2), ("1,8,0.02", 0.0, 1)])
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
        assert np.allclose(rates[r], ro, rtol=RATE_RTOL, atol=0)        # the north-star tolerance ...
        assert _same(rates[r], ro) and _same(ll[r], llo)               # ... is met with 0 ulp


@pytest.mark.parametrize("cluster", ["1", "2", "4", "8", "16"])
def test_em_cluster_sizes_bit_identical(handle, cluster, monkeypatch):
    """A replicate spread over 1, 2, 4, 8 or 16 SMs (thread-block cluster) gives the same bits."""
    monkeypatch.setenv("COLATE_EM_CLUSTER", cluster)
    o = _block_stats()
    counts = po.stage2(np.ones((1, o["num_blocks"]), np.int32), o, 0.0)
    ep, _ = po.epochs_from_bins("3,7,0.2", 0.0, 28.0)
    init = np.full(len(ep), 1 / 20000.)
    rates, iters, ll = handle.stage3_em(1, ep, init, counts, max_iter=40)
    ro, it, llo = po.em_run(ep, init, counts[0], max_iter=40)
    assert iters[0] == it and _same(rates[0], ro) and _same(ll[0], llo)


def test_em_short_run_tight(handle):
    """Few iterations."""
    o = _block_stats()
    counts = po.stage2(np.ones((1, o["num_blocks"]), np.int32), o, 0.0)
    ep, _ = po.epochs_from_bins("3,7,0.2", 0.0, 28.0)
    init = np.full(len(ep), 1 / 20000.)
    rates, iters, ll = handle.stage3_em(1, ep, init, counts, max_iter=5)
    ro, it, llo = po.em_run(ep, init, counts[0], max_iter=5)
    assert iters[0] == it == 5
    assert _same(rates[0], ro)


def test_em_throughput_mode_many_replicates(handle):
    """Enough replicates to fill the GPU: one CTA per replicate (k_em_cta), no cluster; spot-checked against the oracle."""
    o = _block_stats()
    R = 80
    w = api.draw_block_weights(api.mt_seed(11), R, o["num_blocks"])
    counts = po.stage2(w, o, 0.0)
    ep, _ = po.epochs_from_bins("3,7,0.2", 0.0, 28.0)
    init = np.full(len(ep), 1 / 20000.)
    rates, iters, ll = handle.stage3_em(R, ep, init, counts)
    for r in (0, 41, 79):
        ro, it, llo = po.em_run(ep, init, counts[r])
        assert iters[r] == it and _same(rates[r], ro) and _same(ll[r], llo)
    assert len({tuple(x) for x in rates}) > R // 2      # the replicates really differ


@pytest.mark.parametrize("R", [7, 24])
def test_em_mid_range_replicates(handle, R):
    """5..9 replicates run as clusters of 8 on k_em_split, 10 and more one CTA each on k_em_cta; spot-checked against the oracle."""
    o = _block_stats()
    w = api.draw_block_weights(api.mt_seed(12), R, o["num_blocks"])
    counts = po.stage2(w, o, 0.0)
    ep, _ = po.epochs_from_bins("3,7,0.2", 0.0, 28.0)
    init = np.full(len(ep), 1 / 20000.)
    rates, iters, ll = handle.stage3_em(R, ep, init, counts, max_iter=300)
    for r in (0, R // 2, R - 1):
        ro, it, llo = po.em_run(ep, init, counts[r], max_iter=300)
        assert iters[r] == it and _same(rates[r], ro) and _same(ll[r], llo)
    assert len({tuple(x) for x in rates}) > R // 2


def test_em_1000_replicates_three_kernels_bitwise(handle, monkeypatch):
    """BASELINE.json configs[2]'s EM: 1000 block-bootstrap replicates at E = 43 (--bins 3,7,0.1).  Every replicate's
    iteration count, rates and log-likelihood agree BITWISE between the three EM kernels -- k_em_cta (one CTA per
    replicate, work spread over the CTA: the default here), k_em (thread per task) and k_em_split (clusters of 8) --
    and 20 replicates are checked against the oracle (coal.cpp:3675-3827 + coal_EM)."""
    o = _block_stats()
    R = 1000
    w = api.draw_block_weights(api.mt_seed(13), R, o["num_blocks"])
    counts = po.stage2(w, o, 0.0)
    ep, _ = po.epochs_from_bins("3,7,0.1", 0.0, 28.0)
    assert len(ep) == 43
    init = np.full(len(ep), 1 / 20000.)
    res = {}
    for name, env in (("cta", {}), ("task", {"COLATE_EM_KERNEL": "task"}), ("split", {"COLATE_EM_CLUSTER": "8"})):
        with monkeypatch.context() as m:
            for k, v in env.items():
                m.setenv(k, v)
            res[name] = handle.stage3_em(R, ep, init, counts)
    for name in ("task", "split"):
        assert np.array_equal(res["cta"][1], res[name][1]), name                                  # iterations, all 1000
        assert np.array_equal(res["cta"][0].view(np.int64), res[name][0].view(np.int64)), name    # rates, bit patterns
        assert np.array_equal(res["cta"][2].view(np.int64), res[name][2].view(np.int64)), name    # log-likelihoods
    rates, iters, ll = res["cta"]
    assert iters.min() >= 1001 and len({tuple(x) for x in rates}) > R // 2
    for r in list(range(0, R, 53)) + [R - 1]:
        ro, it, llo = po.em_run(ep, init, counts[r])
        assert iters[r] == it and _same(rates[r], ro) and _same(ll[r], llo), r


@pytest.mark.parametrize("kernel", ["cta", "task"])
@pytest.mark.parametrize("bins,age", [("3,7,0.2", 0.0), ("3,7,0.1", 250.0), ("2,8,0.3", 100.0), ("4,6,0.5", 0.0)])
def test_em_one_cta_kernels_on_other_grids(handle, monkeypatch, kernel, bins, age):
    """Both one-CTA-per-replicate kernels on ancient / coarse / wide epoch grids (the last epoch holding ages, epochs
    without ages, ep_null > 0), 40 replicates each, against the oracle."""
    monkeypatch.setenv("COLATE_EM_KERNEL", kernel)
    o = _block_stats()
    R = 40
    w = api.draw_block_weights(api.mt_seed(14), R, o["num_blocks"])
    counts = po.stage2(w, o, age)
    ep, _ = po.epochs_from_bins(bins, age, 28.0)
    init = np.full(len(ep), 1 / 20000.)
    rates, iters, ll = handle.stage3_em(R, ep, init, counts, max_iter=150)
    for r in (0, 17, 39):
        ro, it, llo = po.em_run(ep, init, counts[r], max_iter=150)
        assert iters[r] == it and _same(rates[r], ro) and _same(ll[r], llo), r


def test_em_begin_end_pipeline_equals_the_blocking_calls(handle):
    """colate_stage3_em_begin / _end: the EM of pair A runs on the handle's EM stream while pair B is uploaded and taken
    through stage i; results are those of the blocking calls, bit for bit; the calls that would overwrite the counts of
    the EM in flight are refused."""
    ep, _ = api.epochs_from_bins("3,7,0.2")
    ri = np.full(len(ep), 1 / 20000.0)
    pairs = []
    for seed in (1, 2, 3):
        sites = synth.make_sites(seed, [40000, 25000], [2.5e8, 1.2e8])
        pairs.append((sites, synth.make_genome(seed + 100, sites, 0.7), synth.make_genome(seed + 200, sites, 0.7)))

    def front(p):
        handle.load(*p)
        s1 = handle.stage1(api.mt_seed(7), fetch=False)
        return s1, api.draw_block_weights(s1.mt_state, 2, s1.num_blocks)

    want = []
    for p in pairs:                                   # blocking reference run
        s1, w = front(p)
        handle.stage2_bootstrap_dev(w, None, s1.num_blocks, 0.0)
        want.append(handle.stage3_em(2, ep, ri))
    got = []
    inflight = False
    for p in pairs:
        s1, w = front(p)                              # upload + stage i of this pair under the EM of the previous one
        if inflight:
            with pytest.raises(api._lib.ColateError) as e:
                handle.stage2_bootstrap_dev(w, None, s1.num_blocks, 0.0)
            assert e.value.code == -7
            got.append(handle.stage3_em_end())
        handle.stage2_bootstrap_dev(w, None, s1.num_blocks, 0.0)
        handle.stage3_em_begin(2, ep, ri)
        inflight = True
    got.append(handle.stage3_em_end())
    with pytest.raises(api._lib.ColateError):
        handle.stage3_em_end()                        # nothing in flight any more
    for (r0, i0, l0), (r1, i1, l1) in zip(want, got):
        assert _same(r0, r1) and _same(i0, i1) and _same(l0, l1)
