"""GPU parity: stage i (parse_tmptmp, coal.cpp:2071-2321) through the C-ABI vs the oracle."""
import numpy as np
import pytest

from colate_b200 import api, synth
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def _compare_stage1(o, s1, rel=1e-12):
    assert s1.num_blocks == o["num_blocks"]
    assert s1.n_used == o["n_used_total"]
    # integer tallies: bit-exact
    assert np.array_equal(s1.block_tallies[:, 0], o["n_shared"])
    assert np.array_equal(s1.block_tallies[:, 1], o["n_notshared"])
    assert np.array_equal(s1.block_tallies[:, 2], o["n_emp"])
    # fp64 histograms: bit-exact (the device replays the reference's additions in its order and rounding)
    for v, k in enumerate(("shared", "notshared", "shared_emp", "notshared_emp")):
        assert np.array_equal(s1.block_stats[:, v], o[k]), (k, np.abs(s1.block_stats[:, v] - o[k]).max())
    # generator state after the stage: identical subsequent stream
    a = np.zeros(64, np.uint32); b = np.array([po.lib().oracle_mt_next(o["rng"]) for _ in range(64)], np.uint32)
    st = s1.mt_state.copy()
    api.lib().colate_mt_generate(st, 64, a)
    assert np.array_equal(a, b)


def test_mt_stream_matches_std_mt19937(handle):
    st = api.mt_seed(1)
    for word0, n, k in ((0, 5000, 3), (1600, 40000, 3), (200 * 12345, 30000, 4), (0, 700000, 7), (200 * 1000003, 4000, 5), (400, 0, 3)):
        got, after = handle.mt_stream(st, word0, n, k)
        want = np.zeros(n + 50, np.uint32)
        po.lib().oracle_mt_words(1, word0, n + 50, want)
        assert np.array_equal(got, want[:n]), (word0, n, k)
        nxt = np.zeros(50, np.uint32)
        api.lib().colate_mt_generate(after, 50, nxt)
        assert np.array_equal(nxt, want[n:]), (word0, n, k)


@pytest.mark.parametrize("split", [1, 2, 4, 8])
def test_mt_stream_jump_split_factors(handle, split, monkeypatch):
    """A jump split over P CTAs (the parts of the polynomial meet in global atomics on the zeroed destination window,
    kernels_mt.cu: k_jump) gives the same windows for every P: the stream behind a 5-level jump tree against std::mt19937."""
    monkeypatch.setenv("COLATE_JUMP_SPLIT", str(split))
    st = api.mt_seed(7)
    for word0, n, k in ((0, 200 * 8 * 300, 3), (200 * 777, 200 * 16 * 37, 4)):
        got, after = handle.mt_stream(st, word0, n, k)
        want = np.zeros(n + 50, np.uint32)
        po.lib().oracle_mt_words(7, word0, n + 50, want)
        assert np.array_equal(got, want[:n]), (split, word0, n, k)
        nxt = np.zeros(50, np.uint32)
        api.lib().colate_mt_generate(after, 50, nxt)
        assert np.array_equal(nxt, want[n:]), (split, word0, n, k)


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("masks", [False, True])
def test_stage1_small_weird(handle, seed, masks):
    sites = synth.make_sites(seed, [4000, 3000, 5000], [2.5e8, 1.2e8, 0.9e8], weird=0.1)
    gt = synth.make_genome(seed + 100, sites, 0.7, weird=0.1)
    gr = synth.make_genome(seed + 200, sites, 0.6, weird=0.1)
    tm = rm = None
    if masks:
        tm = [synth.make_mask(seed * 10 + c, int(L) if c != 1 else int(L) // 2, 0.3) for c, L in enumerate(sites.chrom_len)]
        rm = [synth.make_mask(seed * 20 + c, int(L), 0.2) for c, L in enumerate(sites.chrom_len)]
    o = po.stage1(sites, gt, gr, seed=seed, tmask=tm, rmask=rm)
    handle.load(sites, gt, gr, tm, rm)
    s1 = handle.stage1(api.mt_seed(seed))
    _compare_stage1(o, s1)


def test_stage1_config1_shape(handle):
    """chr1-shaped input at 1/5 scale (200 k rows; the oracle finishes in a second)."""
    sites = synth.make_sites(1, [200000], [2.49e8])
    gt = synth.make_genome(101, sites, 0.7)
    gr = synth.make_genome(201, sites, 0.7)
    o = po.stage1(sites, gt, gr, seed=1)
    handle.load(sites, gt, gr)
    s1 = handle.stage1(api.mt_seed(1))
    _compare_stage1(o, s1)
    assert s1.n_used > 20000


def test_stage1_edge_cases(handle):
    # empty chromosome in the middle, chromosome without records, genome with a single record
    sites = synth.make_sites(5, [300, 0, 200, 100], [1e8, 1e8, 4e7, 9e7])
    gt = synth.make_genome(6, sites, 0.9)
    gr = synth.make_genome(7, sites, 0.9)
    keep = gr.chrom != 2
    gr = synth.Genome(gr.chrom[keep], gr.bp[keep], gr.anc[keep], gr.der[keep], gr.aaf[keep], gr.daf[keep])
    o = po.stage1(sites, gt, gr, seed=9)
    handle.load(sites, gt, gr)
    _compare_stage1(o, handle.stage1(api.mt_seed(9)))
    one = synth.Genome(gt.chrom[:1], gt.bp[:1], gt.anc[:1], gt.der[:1], gt.aaf[:1], gt.daf[:1])
    o = po.stage1(sites, one, gr, seed=9)
    handle.load(sites, one, gr)
    s1 = handle.stage1(api.mt_seed(9))
    assert s1.n_used == 0 == o["n_used_total"]
    _compare_stage1(o, s1)


def test_stage1_age_range_error(handle):
    sites = synth.make_sites(1, [500], [1e8])
    sites.age_end[:] = 2e7  # bin >= 185: the reference writes out of bounds here
    gt = synth.make_genome(2, sites, 1.0)
    gr = synth.make_genome(3, sites, 1.0)
    assert po.stage1(sites, gt, gr, seed=1)["num_blocks"] == -2
    handle.load(sites, gt, gr)
    with pytest.raises(api._lib.ColateError) as e:
        handle.stage1(api.mt_seed(1))
    assert e.value.code == -2


def test_config2_whole_genome_full_size(handle):
    """BASELINE.json configs[1] at its full size (22 autosomes, 10,000,000 rows, two ~1x genomes):
    every block histogram, tally and the generator state bit-exact against the oracle, then the
    whole path (stage ii + EM to convergence, E = 43) against the oracle's rates."""
    sites = synth.make_sites(1, synth.rows_for_genome(10_000_000), synth.AUTOSOME_LEN)
    gt = synth.make_genome(101, sites, 0.7)
    gr = synth.make_genome(201, sites, 0.7)
    o = po.stage1(sites, gt, gr, seed=1)
    handle.load(sites, gt, gr)
    s1 = handle.stage1(api.mt_seed(1))
    _compare_stage1(o, s1)
    assert s1.n_used > 1_000_000 and s1.num_blocks > 100
    # one tally per sample: 100 draws per used row end up in the not-shared histogram
    assert int(s1.block_tallies[:, 1].sum()) == 100 * s1.n_used
    w = api.draw_block_weights(s1.mt_state, 1, s1.num_blocks)
    counts = handle.stage2_bootstrap(w, s1.block_stats, 0.0)
    assert np.array_equal(counts, po.stage2(w, o, 0.0))
    ep, _ = po.epochs_from_bins("3,7,0.1", 0.0, 28.0)
    init = np.full(len(ep), 1 / 20000.)
    rates, iters, ll = handle.stage3_em(1, ep, init)          # counts: the device-resident result of stage ii
    ro, it, llo = po.em_run(ep, init, counts[0])
    assert iters[0] == it and np.array_equal(rates[0], ro) and ll[0] == llo


def test_config4_adna_masks(handle):
    """BASELINE.json configs[3] shape (aDNA-like: ~0.5x coverage, mostly single reads, P/N masks with
    40 % / 20 % N in runs, target 7000 years old) on chromosomes 1-3 at the whole-genome row density."""
    lens = synth.AUTOSOME_LEN[:3]
    rows = [int(10_000_000 * L / sum(synth.AUTOSOME_LEN)) for L in lens]
    sites = synth.make_sites(4, rows, lens)
    gt = synth.make_genome(104, sites, 0.35, mean_extra_reads=0.1)
    gr = synth.make_genome(204, sites, 0.35, mean_extra_reads=0.1)
    tm = [synth.make_mask(40 + c, int(L), 0.4) for c, L in enumerate(lens)]
    rm = [synth.make_mask(50 + c, int(L), 0.2) for c, L in enumerate(lens)]
    o = po.stage1(sites, gt, gr, seed=1, tmask=tm, rmask=rm)
    handle.load(sites, gt, gr, tm, rm)
    s1 = handle.stage1(api.mt_seed(1))
    _compare_stage1(o, s1)
    assert s1.n_used > 10_000
    age = 7000.0 / 28.0   # --target_age 7000 --years_per_gen 28 (coal.cpp:3110-3118)
    w = api.draw_block_weights(s1.mt_state, 2, s1.num_blocks)
    counts = handle.stage2_bootstrap(w, s1.block_stats, age)
    assert np.array_equal(counts, po.stage2(w, o, age))
    ep, ep_null = po.epochs_from_bins("3,7,0.1", age, 28.0)
    assert ep_null > 0
    init = np.full(len(ep), 1 / 20000.)
    rates, iters, ll = handle.stage3_em(2, ep, init)
    for r in range(2):
        ro, it, llo = po.em_run(ep, init, counts[r])
        assert iters[r] == it and np.array_equal(rates[r], ro) and ll[r] == llo


def test_config4_adna_masks_whole_genome(handle):
    """BASELINE.json configs[3] at FULL size: 22 autosomes, 10 M rows, ~0.5x genomes of mostly single reads, P/N masks with
    40 % (target) / 20 % (reference) N in runs of 1-50 kb over every chromosome: stage i bit for bit (histograms, tallies,
    generator state), stage ii at age 250 generations."""
    lens = synth.AUTOSOME_LEN
    sites = synth.make_sites(4, synth.rows_for_genome(10_000_000), lens)
    gt = synth.make_genome(104, sites, 0.35, mean_extra_reads=0.1)
    gr = synth.make_genome(204, sites, 0.35, mean_extra_reads=0.1)
    tm = [synth.make_mask(40 + c, int(L), 0.4) for c, L in enumerate(lens)]
    rm = [synth.make_mask(50 + c, int(L), 0.2) for c, L in enumerate(lens)]
    o = po.stage1(sites, gt, gr, seed=1, tmask=tm, rmask=rm)
    handle.load(sites, gt, gr, tm, rm)
    del tm, rm
    s1 = handle.stage1(api.mt_seed(1))
    _compare_stage1(o, s1)
    assert s1.n_used > 30_000 and s1.num_blocks > 100
    age = 7000.0 / 28.0
    w = api.draw_block_weights(s1.mt_state, 3, s1.num_blocks)
    assert np.array_equal(handle.stage2_bootstrap(w, s1.block_stats, age), po.stage2(w, o, age))
    handle.set_mask(0, None); handle.set_mask(1, None)


def test_config5_all_pairs_eight_genomes(handle):
    """BASELINE.json configs[4] at 1 M rows x 8 genomes: all 28 ordered pairs through the batched driver (joins cached per
    genome, one generator stream for all pairs, one EM launch); a third of the pairs against the oracle, bit for bit."""
    from colate_b200 import pairs as pairs_mod
    sites = synth.make_sites(12, synth.rows_for_genome(1_000_000), synth.AUTOSOME_LEN)
    genomes = [synth.make_genome(400 + g, sites, 0.7) for g in range(8)]
    handle.set_sites(sites.site_off, sites.pos, sites.age_begin, sites.age_end, sites.meta())
    for g, G in enumerate(genomes):
        handle.set_genome(g, G.chrom, G.bp, G.aaf, G.daf, G.anc.astype(np.uint16) | (G.der.astype(np.uint16) << 8))
        handle.set_mask(g, None)
    res = pairs_mod.all_pairs(handle, len(genomes), seed=5, bins="3,7,0.1", max_iter=40)
    assert res["pairs"].shape == (28, 2)
    ep, _ = po.epochs_from_bins("3,7,0.1", 0.0, 28.0)
    init = np.full(len(ep), 1 / 20000.)
    for p in range(0, 28, 3):
        i, j = res["pairs"][p]
        o = po.stage1(sites, genomes[i], genomes[j], seed=5)
        assert res["num_blocks"][p] == o["num_blocks"] and res["n_used"][p] == o["n_used_total"]
        w = po.draw_block_weights(o["rng"], 1, o["num_blocks"])
        assert np.array_equal(res["counts"][p], po.stage2(w, o, 0.0)[0]), (i, j)
        ro, it, llo = po.em_run(ep, init, res["counts"][p], max_iter=40)
        assert res["iters"][p] == it and np.array_equal(res["rates"][p], ro) and res["ll"][p] == llo, (i, j)


def test_config5_all_pairs_small(handle):
    """BASELINE.json configs[4] (all pairs of N genomes over one mutation set) at small scale: every
    ordered pair (target=i, reference=j), i<j, through the batched driver == that pair run alone == oracle."""
    from colate_b200 import pairs as pairs_mod
    sites = synth.make_sites(11, [30000, 20000], [2.4e8, 1.3e8])
    genomes = [synth.make_genome(300 + g, sites, 0.7) for g in range(5)]
    handle.set_sites(sites.site_off, sites.pos, sites.age_begin, sites.age_end, sites.meta())
    for g, G in enumerate(genomes):
        handle.set_genome(g, G.chrom, G.bp, G.aaf, G.daf, G.anc.astype(np.uint16) | (G.der.astype(np.uint16) << 8))
        handle.set_mask(g, None)
    res = pairs_mod.all_pairs(handle, len(genomes), seed=3, bins="3,7,0.2", max_iter=60)
    assert res["pairs"].shape == (10, 2)
    ep, _ = po.epochs_from_bins("3,7,0.2", 0.0, 28.0)
    init = np.full(len(ep), 1 / 20000.)
    for p, (i, j) in enumerate(res["pairs"]):
        o = po.stage1(sites, genomes[i], genomes[j], seed=3)
        assert res["num_blocks"][p] == o["num_blocks"] and res["n_used"][p] == o["n_used_total"]
        w = po.draw_block_weights(o["rng"], 1, o["num_blocks"])
        assert np.array_equal(res["counts"][p], po.stage2(w, o, 0.0)[0]), (i, j)
        ro, it, llo = po.em_run(ep, init, res["counts"][p], max_iter=60)
        assert res["iters"][p] == it and np.array_equal(res["rates"][p], ro) and res["ll"][p] == llo, (i, j)
    # asymmetric in target / reference: the swapped pair is a different estimate
    swapped = pairs_mod.all_pairs(handle, len(genomes), seed=3, bins="3,7,0.2", pairs=[(1, 0)], max_iter=60)
    assert not np.array_equal(swapped["rates"][0], res["rates"][0])


def test_stage1_fuzz_shapes(handle):
    """Random shapes: chromosome counts, empty chromosomes, tiny and mid-size row counts, coverages from
    almost nothing to every site, odd rows, masks on either / both genomes -- bit-exact every time."""
    rng = np.random.default_rng(2024)
    for case in range(40):
        n_chr = int(rng.integers(1, 6))
        rows = [int(rng.choice([0, 1, 2, 37, 500, 4000, 20000], p=[.08, .05, .05, .12, .3, .3, .1])) for _ in range(n_chr)]
        lens = [float(rng.choice([3e7, 9e7, 2.4e8])) for _ in range(n_chr)]
        seed = int(rng.integers(1, 1 << 30))
        sites = synth.make_sites(seed % 1000 + 7, rows, lens, weird=float(rng.choice([0.0, 0.1, 0.3])))
        cov_t, cov_r = float(rng.choice([0.05, 0.5, 1.0])), float(rng.choice([0.05, 0.7, 1.0]))
        gt = synth.make_genome(seed % 977 + 1, sites, cov_t, mean_extra_reads=float(rng.choice([0.0, 1.0, 4.0])), weird=0.1)
        gr = synth.make_genome(seed % 991 + 2, sites, cov_r, weird=0.1)
        tm = rm = None
        if rng.random() < 0.4:
            tm = [synth.make_mask(case * 10 + c, int(L), float(rng.choice([0.1, 0.6])), run_lo=100, run_hi=200000) for c, L in enumerate(lens)]
        if rng.random() < 0.4:
            rm = [synth.make_mask(case * 10 + 5 + c, int(L), 0.3, run_lo=100, run_hi=200000) for c, L in enumerate(lens)]
        o = po.stage1(sites, gt, gr, seed=seed, tmask=tm, rmask=rm)
        handle.load(sites, gt, gr, tm, rm)
        if o["num_blocks"] < 0:
            with pytest.raises(api._lib.ColateError):
                handle.stage1(api.mt_seed(seed))
            continue
        _compare_stage1(o, handle.stage1(api.mt_seed(seed)))


def test_fast_bin_index_near_every_threshold(handle):
    """k_sample's table-free bin index (lg2.approx + one FMA, fixed-point in the float's mantissa) around every
    age-bin threshold of coal.cpp:2265: wherever it does not defer to the exact table it IS the exact bin."""
    thr10 = np.zeros(po_nthr(), dtype=np.float64)
    assert api.lib().colate_test_bin_thresholds(thr10) == 0
    j = np.arange(-3000, 3001, dtype=np.float64)
    ages = (thr10[1:, None] / 10.0) * (1.0 + j[None, :] * 1e-7)                 # +-3e-4 relative = +-3e-3 in 10 ln(10 a)
    ulps = np.concatenate([np.nextafter(thr10[1:] / 10.0, np.inf), np.nextafter(thr10[1:] / 10.0, 0), thr10[1:] / 10.0])
    a = np.concatenate([ages.ravel(), ulps, [0.0, 1e-300, 1e-30, 0.05, 1e9, 1e30]])
    fast, exact = handle.bin_fast(a)
    sure = fast >= 0
    assert np.array_equal(fast[sure], exact[sure])
    # the deferral window is narrow (< 1.3e-4 in t on either side of a threshold) ...
    far = np.abs(j) > 200                                                       # |dt| > 2e-4
    assert sure[: ages.size].reshape(ages.shape)[:, far].all()
    # ... and contains every age within 4e-5 of a threshold (the documented error bound of the fast index)
    near = np.abs(j) < 40
    assert not sure[: ages.size].reshape(ages.shape)[:, near].any()
    want = np.array([0, 0, 0, 0, 185, 185])                                     # far outside the grid: bin 0 / "out of range"
    assert np.array_equal(exact[-6:], want) and ((fast[-6:] == want) | (fast[-6:] == -1)).all()


def po_nthr():
    return 186


def test_fast_bin_index_every_float(handle):
    """All 2.7e8 floats from 2^-7 to 2^25 (every age the bins distinguish): no unflagged disagreement with the exact
    table, and the measured error of the fixed-point t stays inside the documented 4.7e-5."""
    flagged, bad, worst = handle.bin_sweep(0x3C000000, 0x4C000000)
    n = 0x4C000000 - 0x3C000000
    assert bad == 0
    assert worst < 4.7e-5, worst
    assert 1e-5 < flagged / n < 1e-3


@pytest.mark.parametrize("world", [2, 3])
def test_stage1_chromosome_shards_on_one_gpu(handle, world):
    """The multi-GPU decomposition (dist.py) replayed on one GPU: each shard samples at its offset of the reference's
    generator stream (colate_stage1_sample with used_rank_base > 0: the stream chunk starts before the shard's first
    row and the tile-ordered layout begins mid-chunk).  Concatenated blocks == the oracle's single run, bit for bit."""
    from colate_b200 import dist as cdist
    sites = synth.make_sites(7, [9000, 2500, 7000, 100, 6000], [2.0e8, 0.6e8, 1.5e8, 0.3e8, 1.2e8])
    gt = synth.make_genome(107, sites, 0.7)
    gr = synth.make_genome(207, sites, 0.7)
    o = po.stage1(sites, gt, gr, seed=3)
    parts = cdist.split_chromosomes(np.diff(sites.site_off), world)
    seed_state = api.mt_seed(3)
    used_base = block_base = 0
    stats, tallies, state = [], [], None
    for lo, hi in parts:
        s0, s1 = int(sites.site_off[lo]), int(sites.site_off[hi])
        handle.set_sites(sites.site_off[lo:hi + 1] - sites.site_off[lo], sites.pos[s0:s1], sites.age_begin[s0:s1],
                         sites.age_end[s0:s1], sites.meta()[s0:s1])
        for slot, g in ((0, gt), (1, gr)):
            first, end = api.chr_ranges(len(sites.chr_names), g.chrom)          # the seek is emulated on the whole file
            al = g.anc.astype(np.uint16) | (g.der.astype(np.uint16) << 8)
            api.check(api.lib().colate_set_genome(handle._h, slot, g.n, api.ptr(np.ascontiguousarray(first[lo:hi])),
                                                  api.ptr(np.ascontiguousarray(end[lo:hi])), api.ptr(g.bp), api.ptr(g.aaf),
                                                  api.ptr(g.daf), api.ptr(al), 0))
        handle.set_mask(0, None); handle.set_mask(1, None)
        used_chr, blocks_chr = handle.stage1_flags()
        n_local = int(np.sum(blocks_chr))
        st, tl, state = handle.stage1_sample(seed_state, used_base, block_base, n_local)
        stats.append(st[:n_local]); tallies.append(tl[:n_local])
        used_base += int(np.sum(used_chr)); block_base += n_local
    assert used_base == o["n_used_total"] and block_base == o["num_blocks"]
    stats, tallies = np.concatenate(stats), np.concatenate(tallies)
    for v, k in enumerate(("shared", "notshared", "shared_emp", "notshared_emp")):
        assert np.array_equal(stats[:, v], o[k]), k
    assert np.array_equal(tallies[:, 1], o["n_notshared"])
    a = np.zeros(64, np.uint32); b = np.array([po.lib().oracle_mt_next(o["rng"]) for _ in range(64)], np.uint32)
    api.lib().colate_mt_generate(state.copy(), 64, a)                           # the last shard holds the final generator state
    assert np.array_equal(a, b)


def test_stream_cache_serves_shorter_and_longer_requests(handle):
    """colate_set_stream_cache: pairs with fewer / more used rows than the stream left in HBM, a different seed and a
    switch back -- histograms and the generator state after the stage are those of the uncached calls."""
    sites = synth.make_sites(21, [6000, 5000], [2.0e8, 1.1e8])
    cover = (0.7, 0.3, 0.9, 0.5, 0.95)                              # used rows differ by far more than the cache's headroom
    genomes = [synth.make_genome(400 + g, sites, c) for g, c in enumerate(cover)]
    handle.set_sites(sites.site_off, sites.pos, sites.age_begin, sites.age_end, sites.meta())
    for g, G in enumerate(genomes):
        handle.set_genome(g, G.chrom, G.bp, G.aaf, G.daf, G.anc.astype(np.uint16) | (G.der.astype(np.uint16) << 8))
        handle.set_mask(g, None)
    calls = [(0, 1, 5), (1, 2, 5), (2, 4, 5), (3, 1, 5), (0, 2, 6), (2, 4, 5), (1, 3, 5)]
    plain = [handle.stage1(api.mt_seed(s), target_slot=i, reference_slot=j) for i, j, s in calls]
    assert len({p.n_used for p in plain}) > 3
    handle.set_stream_cache(True)
    try:
        cached = [handle.stage1(api.mt_seed(s), target_slot=i, reference_slot=j) for i, j, s in calls]
    finally:
        handle.set_stream_cache(False)
    for a, b in zip(plain, cached):
        assert a.n_used == b.n_used and a.num_blocks == b.num_blocks
        assert np.array_equal(a.block_stats, b.block_stats) and np.array_equal(a.block_tallies, b.block_tallies)
        assert np.array_equal(a.mt_state, b.mt_state)


def test_rejection_sampling_rows_match_the_reference(handle):
    """Rows with age_begin > 0 whose interval reaches past the age grid: every sample beyond it is redrawn (coal.cpp:2279-2294),
    two more engine words each.  Device == compiled reference (fixture) == oracle: histograms, tallies and the generator
    state after the stage."""
    from helpers import load, dataset_from
    z = load("stage1_reject.npz")
    sites, gt, gr = dataset_from(z)
    handle.load(sites, gt, gr)
    s1 = handle.stage1(api.mt_seed(int(z["seed"])))
    assert s1.num_blocks == int(z["ref_num_blocks"])
    for v, k in enumerate(("shared", "notshared", "shared_emp", "notshared_emp")):
        assert np.array_equal(s1.block_stats[:, v], z[f"ref_{k}"]), k
    o = po.stage1(sites, gt, gr, seed=int(z["seed"]))
    _compare_stage1(o, s1)
    assert handle.extra_words() > 0 and handle.extra_words() % 2 == 0
    want = api.mt_seed(int(z["seed"]))
    n = 200 * s1.n_used + handle.extra_words()
    api.lib().colate_mt_generate(want, n, np.zeros(n, np.uint32))
    assert np.array_equal(want, s1.mt_state)
    assert (s1.block_tallies[:, 1].sum(), s1.block_tallies[:, 0].sum() <= 100 * s1.n_used) == (100 * s1.n_used, True)


@pytest.mark.parametrize("seed", [41, 42])
def test_rejection_sampling_random(handle, seed):
    sites = synth.make_sites(seed, [30000, 20000, 9000], [2.4e8, 1.3e8, 6e7], weird=0.05)
    deep = synth.add_deep_rows(sites, seed + 7, 0.004)
    gt = synth.make_genome(seed + 100, sites, 0.7)
    gr = synth.make_genome(seed + 200, sites, 0.7)
    o = po.stage1(sites, gt, gr, seed=seed)
    handle.load(sites, gt, gr)
    s1 = handle.stage1(api.mt_seed(seed))
    _compare_stage1(o, s1)
    assert handle.extra_words() > 0
    # without such rows the call consumes exactly 200 words per used row
    sites2 = synth.make_sites(seed, [30000, 20000, 9000], [2.4e8, 1.3e8, 6e7], weird=0.05)
    handle.load(sites2, gt, gr)
    handle.stage1(api.mt_seed(seed))
    assert handle.extra_words() == 0


def test_rows_the_reference_cannot_process_are_refused(handle):
    """age_begin <= 0 with an interval past the grid (the reference writes out of bounds, coal.cpp:2269); age_begin itself
    past the grid (its rejection loop never ends): COLATE_ERR_AGE_RANGE from the flag pass, as in the oracle."""
    for ab, ae in ((0.0, 2e7), (1e7, 3e7), (-5.0, 1.2e7)):
        sites = synth.make_sites(5, [3000], [2.4e8])
        gt = synth.make_genome(105, sites, 0.9)
        gr = synth.make_genome(205, sites, 0.9)
        k = np.nonzero(sites.meta() & 1)[0]
        sites.age_begin[k[::7]] = ab; sites.age_end[k[::7]] = ae
        assert po.stage1(sites, gt, gr, seed=1)["num_blocks"] == -2
        handle.load(sites, gt, gr)
        with pytest.raises(api._lib.ColateError) as e:
            handle.stage1(api.mt_seed(1))
        assert e.value.code == -2


# ---- the bcf / bam front-ends' weighting variant from pre-decoded per-row counts -------------------------
def _load_pileup(handle, sites, t_counts, r_counts, tm=None, rm=None):
    handle.set_sites(sites.site_off, sites.pos, sites.age_begin, sites.age_end, sites.meta())
    handle.set_pileup(0, t_counts)
    handle.set_pileup(1, r_counts)
    handle.set_mask(0, None if tm is None else api.mask_bits_from_seq(tm, sites.site_off, sites.pos))
    handle.set_mask(1, None if rm is None else api.mask_bits_from_seq(rm, sites.site_off, sites.pos))


@pytest.mark.parametrize("tag", ["plain", "masked"])
def test_pileup_front_end_matches_reference_parse_onebambam(handle, tag):
    """colate_set_pileup + front_end = 1 against the REFERENCE's parse_onebambam outputs (coal.cpp:1799-2069 run on synthetic
    reads, tests/golden/stage1_bambam.npz) and against the oracle's tallies, bit for bit, generator state included."""
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import bambam_masks, load, sites_from
    z = load("stage1_bambam.npz")
    sites = sites_from(z)
    tm, rm = bambam_masks(z) if tag == "masked" else (None, None)
    handle.set_option("front_end", 1)
    try:
        _load_pileup(handle, sites, z["t_counts"], z["r_counts"], tm, rm)
        s1 = handle.stage1(api.mt_seed(int(z["seed"])))
        assert s1.num_blocks == int(z[f"ref_{tag}_num_blocks"])
        for v, k in enumerate(("shared", "notshared", "shared_emp", "notshared_emp")):
            assert np.array_equal(s1.block_stats[:, v], z[f"ref_{tag}_{k}"]), k
        o = po.stage1_pileup(sites, z["t_counts"], z["r_counts"], seed=int(z["seed"]), tmask=tm, rmask=rm)
        _compare_stage1(o, s1)
        if tag == "plain":        # stage ii with the 1e3 normalisation + EM: the .colate_mat / .coal the reference CLI wrote
            import tempfile
            for R in (1, 3):
                st = s1.mt_state.copy()
                w = api.draw_block_weights(st, R, s1.num_blocks)
                counts = handle.stage2_bootstrap(w, s1.block_stats, 0.0)
                assert np.array_equal(counts, po.stage2(w, o, 0.0, norm_1e3=True))
                ep, _ = api.epochs_from_bins("3,7,0.2")
                rates, iters, _ = handle.stage3_em(R, ep, np.full(len(ep), 1 / 20000.0), None, 100000)
                with tempfile.TemporaryDirectory() as d:
                    api.write_colate_mat(d + "/o.colate_mat", counts)
                    api.write_coal(d + "/o.coal", ep, rates)
                    g = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"bambam_R{R}")
                    assert open(d + "/o.colate_mat").read() == open(g + ".colate_mat").read()
                    assert open(d + "/o.coal").read() == open(g + ".coal").read()
    finally:
        handle.set_option("front_end", 0)
        handle.set_mask(0, None); handle.set_mask(1, None)


def test_pileup_front_end_larger_random(handle):
    """200 k rows, random pileups (0-6 reads of up to three alleles per row and genome): device == oracle; a pass with rejoin
    (the join columns recomputed from the stored counts) gives the same bits; tmp/tmp weights on the same counts differ."""
    rng = np.random.default_rng(5)
    sites = synth.make_sites(9, [120000, 80000], [2.4e8, 1.3e8], weird=0.05)
    ai = np.searchsorted(np.frombuffer(b"ACGT", np.uint8), sites.anc.clip(65, 84))
    di = np.searchsorted(np.frombuffer(b"ACGT", np.uint8), sites.der.clip(65, 84))

    def pile(p_cov):
        c = np.zeros((sites.n, 4), np.int32)
        cov = rng.random(sites.n) < p_cov
        n_reads = rng.integers(1, 7, sites.n)
        for k in range(6):
            live = cov & (k < n_reads)
            u = rng.random(sites.n)
            col = np.where(u < 0.3, di % 4, np.where(u < 0.93, ai % 4, rng.integers(0, 4, sites.n)))
            np.add.at(c, (np.nonzero(live)[0], col[live]), 1)
        return c

    tc, rc = pile(0.6), pile(0.7)
    o = po.stage1_pileup(sites, tc, rc, seed=3)
    assert o["n_used_total"] > 10000
    handle.set_option("front_end", 1)
    try:
        _load_pileup(handle, sites, tc, rc)
        s1 = handle.stage1(api.mt_seed(3))
        _compare_stage1(o, s1)
        handle.set_option("rejoin", 1)
        s2 = handle.stage1(api.mt_seed(3))
        handle.set_option("rejoin", 0)
        assert np.array_equal(s1.block_stats, s2.block_stats)
        handle.set_option("front_end", 0)
        s3 = handle.stage1(api.mt_seed(3))
        assert s3.n_used == s1.n_used and not np.array_equal(s3.block_stats[:, 1], s1.block_stats[:, 1])
    finally:
        handle.set_option("front_end", 0)
        handle.set_option("rejoin", 0)


def test_pileup_front_end_refuses_rows_the_reference_overruns_on(handle):
    """age_begin > 0 with an interval past the age grid: parse_onebambam has no bound check (coal.cpp:2034-2039)."""
    sites = synth.make_sites(4, [3000], [2.4e8])
    synth.add_deep_rows(sites, 5, 0.05)
    c = np.zeros((sites.n, 4), np.int32)
    c[:, :] = 2
    ai = np.searchsorted(np.frombuffer(b"ACGT", np.uint8), sites.anc.clip(65, 84)) % 4
    di = np.searchsorted(np.frombuffer(b"ACGT", np.uint8), sites.der.clip(65, 84)) % 4
    c[:] = 0
    c[np.arange(sites.n), ai] = 1
    c[np.arange(sites.n), di] += 1
    assert po.stage1_pileup(sites, c, c, seed=1)["num_blocks"] == -2
    handle.set_option("front_end", 1)
    try:
        _load_pileup(handle, sites, c, c)
        with pytest.raises(api._lib.ColateError) as e:
            handle.stage1(api.mt_seed(1))
        assert e.value.code == -2
    finally:
        handle.set_option("front_end", 0)


def test_device_pileup_from_reads_matches_reference_bam_parser(handle):
    """The decoder's counting loop on the device (colate_pileup_*): the synthetic reads of the bam/bam fixture, regenerated from
    their seed, piled up at the .mut rows == the pileup the REFERENCE's bam_parser held at every row (fixture t_counts / r_counts:
    length / mapping-quality / base-quality / mismatch filters, the three bases at either end of a read that never count), and
    stage i on those slots == the reference's parse_onebambam."""
    import os, sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here); sys.path.insert(0, os.path.join(here, "golden"))
    import make_golden
    from helpers import load
    z = load("stage1_bambam.npz")
    sites, lens, genome, reads_t, reads_r = make_golden.bambam_inputs(int(z["seed"]))
    assert np.array_equal(sites.pos, z["pos"])

    def soa(reads):
        per = []
        for c in range(len(lens)):
            rd = [r for r in reads if r[0] == c]
            ln = np.array([len(r[4]) for r in rd], np.int32)
            off = np.concatenate([[0], np.cumsum(ln)[:-1]]).astype(np.int64) if len(rd) else np.zeros(0, np.int64)
            per.append((np.array([r[1] for r in rd], np.int32), np.array([r[2] for r in rd], np.uint8), ln, off,
                        np.frombuffer(b"".join(r[4] for r in rd), np.uint8), np.concatenate([r[5] for r in rd]) if rd else np.zeros(0, np.uint8)))
        return per

    handle.set_sites(sites.site_off, sites.pos, sites.age_begin, sites.age_end, sites.meta())
    handle.set_mask(0, None); handle.set_mask(1, None)
    handle.set_option("front_end", 1)
    try:
        got_t = handle.pileup_from_reads(0, soa(reads_t), genome, fetch=True)
        got_r = handle.pileup_from_reads(1, soa(reads_r), genome, fetch=True)
        assert np.array_equal(got_t, z["t_counts"]) and np.array_equal(got_r, z["r_counts"])
        assert (got_t.sum(1) > 0).sum() > 1000
        s1 = handle.stage1(api.mt_seed(int(z["seed"])))
        assert s1.num_blocks == int(z["ref_plain_num_blocks"])
        for v, k in enumerate(("shared", "notshared", "shared_emp", "notshared_emp")):
            assert np.array_equal(s1.block_stats[:, v], z[f"ref_plain_{k}"]), k
        # unsorted reads are refused like the reference does (htslib.cpp:411-414)
        bad = soa(reads_t)
        p = bad[0][0].copy(); p[5], p[6] = p[6] + 1000, p[5]
        bad[0] = (p,) + bad[0][1:]
        with pytest.raises(api._lib.ColateError) as e:
            handle.pileup_from_reads(0, bad, genome)
        assert e.value.code == -6
    finally:
        handle.set_option("front_end", 0)


@pytest.mark.parametrize("tag", ["plain", "refg", "refg_masked"])
def test_row_counts_front_end_matches_reference_parse_vcfvcf(handle, tag):
    """colate_set_row_counts + front_end = 1 (the bcf front-ends on pre-decoded per-row (AAF, DAF)) against the REFERENCE's
    parse_vcfvcf outputs (coal.cpp:907-1228 run on synthetic genotype records, tests/golden/stage1_vcfvcf.npz): histograms and
    generator state bit for bit; the decoder half is the oracle's restatement (pyoracle.decode_vcfvcf)."""
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import bambam_masks, load, vcfvcf_counts
    z = load("stage1_vcfvcf.npz")
    sites, tc, rc = vcfvcf_counts(z, tag != "plain")
    tm, rm = bambam_masks(z) if tag == "refg_masked" else (None, None)
    handle.set_option("front_end", 1)
    try:
        handle.set_sites(sites.site_off, sites.pos, sites.age_begin, sites.age_end, sites.meta())
        handle.set_row_counts(0, tc[:, 0], tc[:, 1])
        handle.set_row_counts(1, rc[:, 0], rc[:, 1])
        handle.set_mask(0, None if tm is None else api.mask_bits_from_seq(tm, sites.site_off, sites.pos))
        handle.set_mask(1, None if rm is None else api.mask_bits_from_seq(rm, sites.site_off, sites.pos))
        s1 = handle.stage1(api.mt_seed(int(z["seed"])))
        assert s1.num_blocks == int(z[f"ref_{tag}_num_blocks"])
        for v, k in enumerate(("shared", "notshared", "shared_emp", "notshared_emp")):
            assert np.array_equal(s1.block_stats[:, v], z[f"ref_{tag}_{k}"]), k
        o = po.stage1_pileup(sites, tc, rc, seed=int(z["seed"]), tmask=tm, rmask=rm)
        _compare_stage1(o, s1)
    finally:
        handle.set_option("front_end", 0)
        handle.set_mask(0, None); handle.set_mask(1, None)
