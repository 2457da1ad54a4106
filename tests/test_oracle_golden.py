"""CPU: the oracle (oracle/colate_oracle.c) against fixtures generated from the UNMODIFIED reference
(tests/golden/make_golden.py), and -- when oracle/_ref is present -- against the reference itself."""
import os
import tempfile

import numpy as np
import pytest

from colate_b200 import synth
from oracle import pyoracle as po
from helpers import GOLDEN, bambam_masks, dataset_from, load, same, sites_from, vcfvcf_counts


def test_random_streams_match_libstdcxx():
    z = load("random_ref.npz")
    a = np.zeros(4000, np.uint32); po.lib().oracle_mt_words(1, 0, 4000, a)
    assert same(a, z["words_seed1"])
    a = np.zeros(1000, np.uint32); po.lib().oracle_mt_words(123456789, 10**6, 1000, a)
    assert same(a, z["words_seed123456789_skip1e6"])
    u = np.zeros(2000); po.lib().oracle_uniform_real_n(7, 3, 2000, u)
    assert same(u, z["real_seed7_skip3"])
    for nb in (1, 9, 105, 500):
        i = np.zeros(3000, np.int32); po.lib().oracle_uniform_int_n(3, 11, nb, 3000, i)
        assert same(i, z[f"int_{nb}"])


@pytest.mark.parametrize("tag", ["nomask", "mask"])
def test_stage1_matches_reference_parse_tmptmp(tag):
    z = load("stage1_small.npz")
    sites, gt, gr = dataset_from(z)
    tm = rm = None
    if tag == "mask":
        # the fixture stores the masks as pass bits at the row positions: rebuild tiny per-chromosome
        # sequences that give the same answers (mask longer than every position)
        def seqs(bits):
            out = []
            for c in range(len(sites.chr_names)):
                lo, hi = int(sites.site_off[c]), int(sites.site_off[c + 1])
                L = int(sites.pos[lo:hi].max()) + 2 if hi > lo else 2
                s = np.full(L, ord("P"), np.uint8)
                idx = np.arange(lo, hi)
                ok = (bits[idx >> 5] >> (idx & 31)) & 1
                s[sites.pos[lo:hi][ok == 0] - 1] = ord("N")
                out.append(s.tobytes())
            return out
        tm, rm = seqs(z["tmask_bits"]), seqs(z["rmask_bits"])
    o = po.stage1(sites, gt, gr, seed=int(z["seed"]), tmask=tm, rmask=rm)
    assert o["num_blocks"] == int(z[f"ref_{tag}_num_blocks"])
    for k in ("shared", "notshared", "shared_emp", "notshared_emp"):
        assert same(o[k], z[f"ref_{tag}_{k}"]), k            # bit-exact, same summation order
    assert same(o["rng"].words(), z[f"ref_{tag}_mt"])         # generator state after the stage
    # integer tallies are consistent with the fp64 vectors
    assert o["n_notshared"].sum() == 100 * o["n_used_total"]
    assert (o["n_shared"] <= o["n_notshared"]).all()


@pytest.mark.parametrize("tag", ["plain", "masked"])
def test_stage1_pileup_matches_reference_parse_onebambam(tag):
    """the oracle's bam/bam front-end on pre-decoded pileups against the reference's own parse_onebambam
    (coal.cpp:1799-2069) run on synthetic reads (fixture: make_golden.py bambam; the pileups are the reference bam_parser's)."""
    z = load("stage1_bambam.npz")
    sites = sites_from(z)
    tm, rm = bambam_masks(z) if tag == "masked" else (None, None)
    o = po.stage1_pileup(sites, z["t_counts"], z["r_counts"], seed=int(z["seed"]), tmask=tm, rmask=rm)
    assert o["num_blocks"] == int(z[f"ref_{tag}_num_blocks"])
    assert o["n_used_total"] > 100
    for k in ("shared", "notshared", "shared_emp", "notshared_emp"):
        assert same(o[k], z[f"ref_{tag}_{k}"]), k
    assert same(o["rng"].words(), z[f"ref_{tag}_mt"])
    # the filter cases the fixture is meant to hold: three alleles seen, reads without either allele, uncovered rows
    for cnt in (z["t_counts"], z["r_counts"]):
        assert ((cnt > 0).sum(1) >= 3).any() and (cnt.sum(1) == 0).any()


@pytest.mark.parametrize("R", [1, 3])
def test_bambam_cli_chain_matches_reference_cli(R, tmp_path):
    """The oracle's whole bam/bam chain -- stage i on pileups, block bootstrap, F redistribution, the 1e3 normalisation of
    coal.cpp:3453-3463, EM -- against the <out>.colate_mat and <out>.coal the reference CLI wrote for the same inputs
    (--target_bam / --reference_bam on synthetic reads, make_golden.py bambam)."""
    from colate_b200 import api
    z = load("stage1_bambam.npz")
    sites = sites_from(z)
    o = po.stage1_pileup(sites, z["t_counts"], z["r_counts"], seed=int(z["seed"]))
    w = po.draw_block_weights(o["rng"], R, o["num_blocks"])
    counts = po.stage2(w, o, 0.0, norm_1e3=True)
    api.write_colate_mat(str(tmp_path / "o.colate_mat"), counts)             # the product's host writer (no GPU involved)
    assert open(tmp_path / "o.colate_mat").read() == open(os.path.join(GOLDEN, f"bambam_R{R}.colate_mat")).read()
    ep, _ = po.epochs_from_bins("3,7,0.2")
    rates = np.stack([po.em_run(ep, np.full(len(ep), 1 / 20000.0), counts[r])[0] for r in range(R)])
    po.write_coal(str(tmp_path / "o.coal"), ep, rates)
    assert open(tmp_path / "o.coal").read() == open(os.path.join(GOLDEN, f"bambam_R{R}.coal")).read()


@pytest.mark.parametrize("tag", ["plain", "refg", "refg_masked"])
def test_vcfvcf_front_end_matches_reference_parse_vcfvcf(tag):
    """SURVEY.md 8(f) N3, the bcf/bcf front-end: the decoder half of parse_vcfvcf restated (cursor over the genotype records, allele
    match / flip, biallelic test, "alt not reported" records, --ref_genome fall-back: coal.cpp:997-1137) feeding the SAME count-based
    stage i as the bam front-end (weights (N_target - DAF_target) * DAF_ref / (N_ref * 100), coal.cpp:1164-1197) against the
    reference's own parse_vcfvcf run on synthetic genotype records (fixture: make_golden.py vcfvcf)."""
    z = load("stage1_vcfvcf.npz")
    sites, tc, rc = vcfvcf_counts(z, tag != "plain")
    tm, rm = bambam_masks(z) if tag == "refg_masked" else (None, None)
    o = po.stage1_pileup(sites, tc, rc, seed=int(z["seed"]), tmask=tm, rmask=rm)
    assert o["num_blocks"] == int(z[f"ref_{tag}_num_blocks"])
    assert o["n_used_total"] > 100
    for k in ("shared", "notshared", "shared_emp", "notshared_emp"):
        assert same(o[k], z[f"ref_{tag}_{k}"]), k
    assert same(o["rng"].words(), z[f"ref_{tag}_mt"])


def test_pileup_restatement_matches_reference_bam_parser():
    """The oracle's pileup of decoded reads (incl. the reference's first-read quirk) against the counts the reference's own
    bam_parser held at every row (fixture, make_golden.py bambam); the reads are regenerated from the fixture's seed."""
    import sys
    sys.path.insert(0, GOLDEN)
    import make_golden
    z = load("stage1_bambam.npz")
    sites, lens, genome, reads_t, reads_r = make_golden.bambam_inputs(int(z["seed"]))
    for reads, key in ((reads_t, "t_counts"), (reads_r, "r_counts")):
        for c in range(len(lens)):
            lo, hi = int(sites.site_off[c]), int(sites.site_off[c + 1])
            got = po.pileup_from_reads(sites.pos[lo:hi], [(r[1], r[2], r[4], r[5]) for r in reads if r[0] == c], genome[c])
            assert same(got, z[key][lo:hi]), (key, c)


def test_estep_matches_reference_coal_EM():
    z = load("estep_ref.npz")
    ab = po.age_bins()
    for tag in ("b02", "b01", "anc"):
        ep, rates = z[f"{tag}_epochs"], z[f"{tag}_rates"]
        for i in range(rates.shape[0]):
            for s in (0, 1):
                for b in range(0, 185, 3):
                    ll, num, den = po.estep(s == 0, ep, rates[i], ab[b])
                    assert same(ll, z[f"{tag}_ll"][i, s, b]) and same(num, z[f"{tag}_num"][i, s, b]) and same(den, z[f"{tag}_den"][i, s, b])


def test_epoch_grids():
    z = load("estep_ref.npz")
    for tag, bins, age in (("b02", "3,7,0.2", 0.0), ("b01", "3,7,0.1", 0.0), ("anc", "3,7,0.2", 250.0)):
        ep, null = po.epochs_from_bins(bins, age, 28.0)
        assert same(ep, z[f"{tag}_epochs"]) and null == int(z[f"{tag}_ep_null"])
    ep, _ = po.epochs_from_bins("3,7,0.2")
    assert len(ep) == 23 and ep[0] == 0 and ep[1] == 0          # first boundary dropped
    assert len(po.epochs_from_bins("3,7,0.1")[0]) == 43
    a, ypg = po.ages("7000", "0", 28.0)
    assert a == 250.0 and ypg == 28.0
    assert po.epochs_from_bins("3,7,0.2", a, ypg)[1] == 5


@pytest.mark.parametrize("name", ["bins02_R1", "bins02_R3", "ancient", "bins01_R1"])
def test_whole_path_coal_text_identical(name):
    """oracle stage i -> ii -> iii -> .coal text == the reference CLI's file, byte for byte."""
    z = load("cli_small.npz")
    sites, gt, gr = dataset_from(z)
    args = [str(x) for x in z[f"{name}_args"]]
    opt = dict(zip(args[::2], args[1::2]))
    R = int(opt.get("--num_bootstraps", 1))
    o = po.stage1(sites, gt, gr, seed=int(z["seed"]))
    assert o["num_blocks"] == int(z[f"{name}_num_blocks"])
    w = po.draw_block_weights(o["rng"], R, o["num_blocks"])
    age, ypg = po.ages(opt.get("--target_age"), opt.get("--reference_age"), float(opt["--years_per_gen"]) if "--years_per_gen" in opt else None)
    counts = po.stage2(w, o, age)
    ep, null = po.epochs_from_bins(opt["--bins"], age, ypg)
    rates = np.stack([po.em_run(ep, np.full(len(ep), 1 / 20000.), counts[r])[0] for r in range(R)])
    with tempfile.TemporaryDirectory() as d:
        po.write_coal(os.path.join(d, "o.coal"), ep, rates, age > 0, null)
        assert open(os.path.join(d, "o.coal")).read() == open(os.path.join(GOLDEN, f"cli_{name}.coal")).read()


def test_site_meta_filter():
    L = po.lib()
    ok = L.oracle_site_meta(0, 1, 1.0, 2.0, b"A/G")
    assert ok == (1 | (ord("A") << 8) | (ord("G") << 16))
    for args in ((1, 1, 1.0, 2.0, b"A/G"), (0, 2, 1.0, 2.0, b"A/G"), (0, 1, 2.0, 2.0, b"A/G"), (0, 1, -3.0, -1.0, b"A/G"),
                 (0, 1, 1.0, 2.0, b"AT/G"), (0, 1, 1.0, 2.0, b"NA"), (0, 1, 1.0, 2.0, b"A/"), (0, 1, 1.0, 2.0, b"/G"),
                 (0, 1, 1.0, 2.0, b"N/G"), (0, 1, 1.0, 2.0, b"A/1x")):
        assert L.oracle_site_meta(*args) == 0, args
    assert L.oracle_site_meta(0, 1, -1.0, 0.0, b"0/1") != 0     # age_end >= 0 passes, '0'/'1' codes allowed


def test_bin_index():
    L = po.lib()
    assert L.oracle_bin_of_double_age(0.0) == 0                 # log(0) -> INT_MIN+1 -> max(0, .)
    assert L.oracle_bin_of_double_age(1e-9) == 0
    ab = po.age_bins()
    for b in range(1, 185):                                       # age_bin[b] = exp((b-1)/10)/10 sits in bin b
        assert L.oracle_bin_of_double_age(ab[b] * (1 + 1e-9)) == b
    if po.ref_available():
        rng = np.random.default_rng(5)
        for a in np.exp(rng.uniform(np.log(1e-3), np.log(5e7), 20000)):
            assert L.oracle_bin_of_double_age(a) == po.ref().ref_bin_of_double_age(a)
            assert L.oracle_bin_of_float_age(np.float32(a)) == po.ref().ref_bin_of_float_age(np.float32(a))


def test_stage1_rejection_sampling_matches_reference():
    """Rows with age_begin > 0 whose interval reaches past the age grid: the reference redraws every sample whose bin
    reaches 185 (coal.cpp:2279-2294), two more engine words per redraw.  Histograms AND generator state after the
    stage equal the compiled reference's (fixture from tests/golden/make_golden.py reject)."""
    z = load("stage1_reject.npz")
    sites, gt, gr = dataset_from(z)
    o = po.stage1(sites, gt, gr, seed=int(z["seed"]))
    assert o["num_blocks"] == int(z["ref_num_blocks"])
    for k in ("shared", "notshared", "shared_emp", "notshared_emp"):
        assert same(o[k], z[f"ref_{k}"]), k
    assert same(o["rng"].words(), z["ref_mt"])
    # the stage consumed more than 200 words per used row: redraws happened
    plain = po.mt_seed(int(z["seed"]))
    for _ in range(200 * o["n_used_total"]):
        po.lib().oracle_mt_next(plain)
    assert not same(plain.words(), z["ref_mt"])
    # rows the reference cannot process stay errors: age_begin <= 0 with an interval past the grid (out-of-bounds write,
    # coal.cpp:2269), age_begin itself past the grid (the rejection loop never ends)
    for ab, ae in ((0.0, 2e7), (1e7, 3e7)):
        s2, g2t, g2r = dataset_from(z)
        used = np.nonzero(s2.meta() & 1)[0]
        s2.age_begin[used] = ab; s2.age_end[used] = ae
        assert po.stage1(s2, g2t, g2r, seed=1)["num_blocks"] == -2


@pytest.mark.skipif(not po.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("seed", [21, 22])
def test_stage1_random_vs_live_reference(seed):
    sites = synth.make_sites(seed, [2500, 1800], [2.5e8, 1.2e8], weird=0.15)
    gt = synth.make_genome(seed + 100, sites, 0.5, mean_extra_reads=0.3, weird=0.15)   # aDNA-like: sparse, N mostly 1
    gr = synth.make_genome(seed + 200, sites, 0.8, weird=0.15)
    with tempfile.TemporaryDirectory() as d:
        synth.write_dataset(d, sites, {"t": gt, "r": gr})
        r = po.ref_parse_tmptmp(d, sites.chr_names, "syn", "t", "r", seed=seed)
    o = po.stage1(sites, gt, gr, seed=seed)
    assert o["num_blocks"] == r["num_blocks"]
    for k in ("shared", "notshared", "shared_emp", "notshared_emp"):
        assert same(o[k], r[k])
    assert same(o["rng"].words(), r["mt"])
