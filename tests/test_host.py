"""CPU: host side of the product (readers, seek emulation, RNG, epoch grid, writers, CLI) and the C ABI surface.
No compute entry point is exercised here -- those need a GPU and fail loudly without one."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from colate_b200 import _lib, api, synth
from oracle import pyoracle as po
from helpers import GOLDEN, dataset_from, load, same

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "colate_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(colate_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    L = C.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/colate_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert b"sm_100a" in api.lib().colate_version()


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.ColateError) as e:
        api.Handle(0)
    assert e.value.code == -4 and "no CPU fallback" in str(e.value)


def test_product_does_not_touch_the_oracle():
    for dp, _, files in os.walk(os.path.join(ROOT, "colate_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "oracle/" not in txt.replace("oracle/_ref", ""), os.path.join(dp, f)


def test_mt19937_window_generator(built):
    z = load("random_ref.npz")
    w = api.mt_seed(1)
    out = np.zeros(4000, np.uint32)
    api.lib().colate_mt_generate(w, 4000, out)
    assert same(out, z["words_seed1"])
    w = api.mt_seed(123456789)
    burn = np.zeros(10**6, np.uint32); api.lib().colate_mt_generate(w, 10**6, burn)
    out = np.zeros(1000, np.uint32); api.lib().colate_mt_generate(w, 1000, out)
    assert same(out, z["words_seed123456789_skip1e6"])
    # odd chunk sizes keep the window consistent
    w = api.mt_seed(1)
    parts = []
    for n in (1, 5, 623, 624, 625, 7, 1248, 867):
        o = np.zeros(n, np.uint32); api.lib().colate_mt_generate(w, n, o); parts.append(o)
    assert same(np.concatenate(parts), z["words_seed1"][:sum(len(p) for p in parts)])


def test_characteristic_polynomial_and_jumps(built):
    terms = np.zeros(1000, np.int32)
    n = api.lib().colate_test_charpoly_terms(terms, 1000)
    assert n == 134 and terms[0] == 0                       # 135 terms with the leading t^19937
    for q in (0, 3, 9, 17, 20):
        w = api.mt_seed(77)
        jw = np.zeros(624, np.uint32)
        assert api.lib().colate_test_jump_window_host(w, q, jw) == 0
        a = np.zeros(700, np.uint32); api.lib().colate_mt_generate(jw, 700, a)
        b = np.zeros(700, np.uint32); po.lib().oracle_mt_words(77, 200 << q, 700, b)
        assert same(a, b), q


def test_block_weights_match_libstdcxx_uniform_int(built):
    for R, nb, burn in ((1, 9, 0), (5, 9, 777), (1000, 105, 200 * 31), (3, 1, 5), (4, 500, 0)):
        w = api.mt_seed(11)
        tmp = np.zeros(max(burn, 1), np.uint32); api.lib().colate_mt_generate(w, burn, tmp)
        W = api.draw_block_weights(w, R, nb)
        g = po.mt_seed(11)
        for _ in range(burn):
            po.lib().oracle_mt_next(g)
        assert same(W, po.draw_block_weights(g, R, nb))
        assert (W.sum(axis=1) == nb).all()
        nxt = np.zeros(8, np.uint32); api.lib().colate_mt_generate(w, 8, nxt)
        assert same(nxt, np.array([po.lib().oracle_mt_next(g) for _ in range(8)], np.uint32))


def test_bin_thresholds_reproduce_the_bin_index(built):
    thr = np.zeros(186)
    assert api.lib().colate_test_bin_thresholds(thr) == 0    # host log() monotone around every threshold
    rng = np.random.default_rng(0)
    a = np.concatenate([np.exp(rng.uniform(np.log(1e-4), np.log(9e6), 200000)), thr[1:] / 10, np.nextafter(thr[1:] / 10, 0), [0.0]])
    x10 = 10 * a
    want = np.array([po.lib().oracle_bin_of_double_age(v) for v in a])
    got = np.searchsorted(thr[1:], x10, side="right")
    assert same(got, want)


def test_repeated_rounded_addition_matches_the_plain_loop(built):
    """exact_sum.cuh (k_replay's slow path): acc += w, c times, each addition rounded (coal.cpp:2269, 2291-2292)."""
    f = api.lib().colate_test_add_repeated
    rng = np.random.default_rng(5)

    def plain(acc, w, c):
        acc = np.float64(acc)
        for _ in range(c):
            acc = np.float64(acc + np.float64(w))
        return float(acc)

    cases = []
    for _ in range(4000):
        w = float(rng.integers(0, 3) * rng.integers(0, 40) / (rng.integers(1, 60) * 100.0))   # weights like num / (reads * 100)
        acc = float(rng.choice([0.0, w, rng.uniform(0, 1e-3), rng.uniform(0, 4), np.exp(rng.uniform(-5, 12))]))
        cases.append((acc, w, int(rng.integers(1, 101))))
    for e in range(-6, 14):       # sums that step over a power of two, with and without landing on it
        top = 2.0 ** e
        for w in (top / 64, top / 64 * (1 + 2.0 ** -30), 0.3 * top, top / 3, top * 2.0 ** -52, top * 2.0 ** -53, top * 1.5 * 2.0 ** -53):
            for back in (1, 2, 7, 63, 64, 65):
                cases.append((float(np.float64(top) - np.float64(back) * np.float64(w)), float(w), 100))
                cases.append((float(np.nextafter(top, 0)), float(w), 37))
    for acc, w, c in cases:
        assert f(acc, w, c) == plain(acc, w, c), (acc, w, c)


def test_libm_port_is_this_hosts_libm(built):
    assert api.libm_exact()


def test_straight_line_log1p_of_the_em_folds(built):
    """glm::log1p_wide (all branches of fdlibm's log1p for 0 < x < 1 as selects, used by the one-fold-per-lane
    logsumexp of k_em_cta) == this host's log1p wherever log1p_wide_ok says so, and that is nearly everywhere."""
    if not api.libm_exact():
        pytest.skip("this host's libm is not glibc 2.39 / FMA")
    import math
    rng = np.random.default_rng(3)
    xs = [np.exp(-rng.uniform(0, 45, 200000)), rng.uniform(0.40, 0.43, 100000), rng.uniform(0.9999, 1.0, 50000),
          np.exp(-rng.uniform(0, 700, 50000)), rng.uniform(0, 1, 200000),
          np.array([2.0 ** -54, 2.0 ** -29, np.nextafter(2.0 ** -54, 0), np.nextafter(2.0 ** -29, 0), math.sqrt(2) - 1, 1.0, 0.5, 0.41421356, 0.41422])]
    edges = np.array([0x3c8fffff, 0x3c900000, 0x3e1fffff, 0x3e200000, 0x3fda8279, 0x3fda827a], dtype=np.uint64) << np.uint64(32)
    xs.append(np.concatenate([(edges + np.uint64(k)).view(np.float64) for k in (0, 1, 0xffffffff)]))
    x = np.concatenate(xs)
    x = x[(x > 0) & (x <= 1)]
    y = np.zeros_like(x); ok = np.zeros(x.shape[0], np.int32)
    api.lib().colate_test_log1p_wide(x.shape[0], np.ascontiguousarray(x), y, ok)
    want = np.array([math.log1p(float(v)) for v in x])      # libm's log1p (numpy's array log1p is its own SIMD code)
    good = ok == 1
    assert np.array_equal(y[good].view(np.int64), want[good].view(np.int64))
    assert good.mean() > 0.9 and good[x < 0.41].all()
    assert not ok[x == 1.0].any()


def test_ages_epochs_and_age_bins_match_oracle(built):
    assert same(api.age_bins(), po.age_bins())
    for ta, ra, ypg in ((None, None, None), ("7000", "0", 28.0), ("1e4", "23000", 29.5), ("0", "500", None)):
        assert api.ages(ta, ra, ypg) == po.ages(ta, ra, ypg)
        age, y = api.ages(ta, ra, ypg)
        for bins in ("3,7,0.2", "3,7,0.1", "2.5,6.5,0.25", "4,7,0.3"):
            e1, n1 = api.epochs_from_bins(bins, age, y)
            e2, n2 = po.epochs_from_bins(bins, age, y)
            assert same(e1, e2) and n1 == n2
    with pytest.raises(_lib.ColateError):
        api.epochs_from_bins("3,7")


def test_coal_file_roundtrip(built):
    ep, _ = api.epochs_from_bins("3,7,0.2")
    ep = ep[1:]          # [0, 0, ...] of a non-ancient run trips the reference's own assert (coal.cpp:3546-3549) when fed back
    rates = np.abs(np.random.default_rng(1).normal(1e-4, 3e-5, (2, len(ep))))
    with tempfile.TemporaryDirectory() as d:
        api.write_coal(os.path.join(d, "a.coal"), ep, rates)
        po.write_coal(os.path.join(d, "b.coal"), ep, rates)
        assert open(os.path.join(d, "a.coal")).read() == open(os.path.join(d, "b.coal")).read()
        with pytest.raises(_lib.ColateError):
            api.write_coal(os.path.join(d, "bad.coal"), np.concatenate([[0.0], ep]), np.ones((1, len(ep) + 1)))
            api.epochs_from_coal_file(os.path.join(d, "bad.coal"))
        e2, r2 = api.epochs_from_coal_file(os.path.join(d, "a.coal"))
        # --coal goes through stof (float) for the epochs and operator>> (double) for the rates, like the reference
        line = open(os.path.join(d, "a.coal")).read().split("\n")[1]
        assert same(e2, po.epochs_from_coal_line(line))
        assert np.allclose(r2, rates[0], rtol=1e-5)
        api.write_bin(os.path.join(d, "a.bin"), ep, rates, np.array([1001, 1274], np.int32))
        raw = open(os.path.join(d, "a.bin"), "rb").read()
        assert raw[:8] == b"COLATEB1" and np.frombuffer(raw[8:16], np.int32).tolist() == [2, len(ep)]
        assert same(np.frombuffer(raw[16 + 8 * len(ep):16 + 8 * len(ep) * 3], np.float64).reshape(2, -1), rates)


def _brute_chr_ranges(n_chr, chrom):
    cur, nxt, n = -1, 0, len(chrom)
    first, end = [], []
    for c in range(n_chr):
        while not (cur >= 0 and chrom[cur] == c):
            if nxt >= n:
                break
            cur = nxt; nxt += 1
        if cur >= 0 and chrom[cur] == c:
            e = cur + 1
            while e < n and chrom[e] == c:
                e += 1
            first.append(cur); end.append(e); cur, nxt = e - 1, e
        else:
            first.append(-1); end.append(-1)
    return np.array(first), np.array(end)


def test_chr_ranges_emulate_the_sequential_seek(built):
    rng = np.random.default_rng(2)
    cases = [np.array([0, 0, 1, 1, 2]), np.array([1, 1, 2, 2]), np.array([2, 0, 0, 1]), np.array([], dtype=np.int32),
             np.array([0, 0, 3, 3, 1, 0, 2]), np.array([1]), np.array([5, 5, 5])]
    cases += [np.sort(rng.integers(0, 4, 30)) for _ in range(5)] + [rng.integers(0, 5, 25) for _ in range(20)]
    for ch in cases:
        f, e = api.chr_ranges(3, ch.astype(np.int32))
        bf, be = _brute_chr_ranges(3, ch)
        assert same(f, bf) and same(e, be), ch


def _colate_in_image(chrom_ids, names, rng):
    """Record stream with the given chromosome id per record (ids index `names`)."""
    out = bytearray()
    for k, c in enumerate(chrom_ids):
        nm = names[int(c)].encode()
        out += np.array([len(nm)], dtype="<i4").tobytes() + nm + np.array([1000 + 3 * k], dtype="<i4").tobytes()
        out += bytes([65 + int(rng.integers(0, 4)), 67]) + np.array([int(rng.integers(0, 5)), int(rng.integers(0, 5))], dtype="<i4").tobytes()
    return bytes(out)


def test_colate_in_run_finder_matches_the_sequential_reader(built, tmp_path):
    """colate_ingest_colate_in's host half: runs found by galloping + run-level chromosome seek == the record-by-record
    decode + colate_chr_ranges whenever every record carries its run's header (what the device verifies)."""
    rng = np.random.default_rng(11)
    names_all = ["1", "2", "3", "10", "X", "chr22"]          # widths 19, 19, 19, 20, 19, 23
    chr_list = ["1", "2", "3", "10"]                         # "X" and "chr22" are not in --chr
    cases = [np.repeat([0, 1, 2, 3], [5, 1, 7, 3]), np.repeat([1, 2], [4, 4]), np.repeat([4, 0, 5, 3], [3, 6, 2, 9]),
             np.array([], dtype=np.int64), np.array([2]), np.repeat([0, 3, 0], [8, 2, 8]), np.repeat([0, 1, 0, 1], [600, 7, 500, 1])]
    cases += [np.repeat(rng.integers(0, 6, 6), rng.integers(1, 40, 6)) for _ in range(30)]
    cases += [rng.integers(0, 6, 40) for _ in range(10)]                     # interleaved: mostly runs of length 1
    verified = 0
    for ch in cases:
        img = _colate_in_image(ch, names_all, rng)
        for cut in (0, 5):                                                   # a truncated last record ends the stream
            image = img[:len(img) - cut] if len(img) > cut else img
            path = str(tmp_path / "x.colate.in")
            open(path, "wb").write(image)
            rc, bp, aaf, daf, al = api.read_colate_in(path, chr_list)
            f0, e0 = api.chr_ranges(len(chr_list), rc)
            n_rec, runs, f1, e1 = api.colate_in_runs(image, chr_list)
            # the device's check, here on the host: every record of a run starts with the run's header
            ok = True
            for off, w, cid, n in runs:
                hl = w - 14
                first = image[off:off + hl]
                ok &= all(image[off + j * w: off + j * w + hl] == first for j in range(n))
            if not ok:
                continue
            verified += 1
            assert n_rec == rc.shape[0]
            assert same(np.repeat(runs[:, 2], runs[:, 3]).astype(np.int32), rc)
            assert same(f0, f1) and same(e0, e1), (ch, cut)
    assert verified >= 60
    # a clean file: one run per chromosome, found with O(log n) probes per run
    ch = np.repeat([0, 1, 2, 3], [5000, 3000, 1, 4000])
    n_rec, runs, _, _ = api.colate_in_runs(_colate_in_image(ch, names_all, rng), chr_list)
    assert n_rec == 12001 and runs.shape[0] == 4 and same(runs[:, 3], np.array([5000, 3000, 1, 4000]))


def test_readers_roundtrip_against_written_files(built):
    z = load("stage1_small.npz")
    sites, gt, gr = dataset_from(z)
    with tempfile.TemporaryDirectory() as d:
        synth.write_dataset(d, sites, {"t": gt, "r": gr})
        metas = sites.meta()
        for c, nm in enumerate(sites.chr_names):
            lo, hi = int(sites.site_off[c]), int(sites.site_off[c + 1])
            pos, ab, ae, meta = api.read_mut(os.path.join(d, f"syn_chr{nm}.mut"))
            assert same(pos, sites.pos[lo:hi]) and same(ab, sites.age_begin[lo:hi]) and same(ae, sites.age_end[lo:hi])
            assert same(meta, metas[lo:hi])
        subprocess.run(["gzip", "-k", os.path.join(d, "syn_chr1.mut")], check=True)
        os.remove(os.path.join(d, "syn_chr1.mut"))                      # Mutations::Read falls back to <file>.gz
        pos, _, _, _ = api.read_mut(os.path.join(d, "syn_chr1.mut"))
        assert same(pos, sites.pos[:int(sites.site_off[1])])
        for nm, g in (("t", gt), ("r", gr)):
            rc, bp, aaf, daf, al = api.read_colate_in(os.path.join(d, nm + ".colate.in"), sites.chr_names)
            assert same(rc, g.chrom) and same(bp, g.bp) and same(aaf, g.aaf) and same(daf, g.daf)
            assert same(al, g.anc.astype(np.uint16) | (g.der.astype(np.uint16) << 8))
        with pytest.raises(_lib.ColateError):
            api.read_mut(os.path.join(d, "missing.mut"))
        # site meta from the row filter == the oracle's restatement, row by row
        for i in range(0, sites.n, 7):
            mt = {0: chr(sites.anc[i]) + "/" + chr(sites.der[i]), 1: chr(sites.anc[i]) + "T/" + chr(sites.der[i]), 2: "NA"}[int(sites.odd[i])]
            a = api.lib().colate_site_meta(int(sites.flipped[i]), int(sites.n_branch[i]), float(sites.age_begin[i]), float(sites.age_end[i]), mt.encode())
            b = po.lib().oracle_site_meta(int(sites.flipped[i]), int(sites.n_branch[i]), float(sites.age_begin[i]), float(sites.age_end[i]), mt.encode())
            assert a == b == metas[i]


def test_mut_reader_pinned_to_the_reference_reader(built, tmp_path):
    """colate_read_mut against Mutations::Read ITSELF (include/src/mutations.cpp:56-283, run through oracle/ref_probe.cpp;
    fixture tests/golden/mut_reader.npz from make_golden.py mut_reader): 3643 rows of number spellings std::stoi / std::stof
    accept (signs, blanks, leading zeros, trailing garbage, exponents, inf / nan / hex, float midpoints with all their digits)
    give the same positions, age bit patterns and row filter; every line the reference dies on -- a field std::stoi or
    std::stof throws on (no digits, out of range, ERANGE overflow AND underflow), missing fields -- is refused here too."""
    z = load("mut_reader.npz")
    p = str(tmp_path / "good.mut")
    open(p, "wb").write(z["text"].tobytes())
    pos, ab, ae, meta = api.read_mut(p)
    assert same(pos, z["pos"]) and same(meta, z["meta"])
    assert same(ab.view(np.uint32), z["age_begin"].view(np.uint32)) and same(ae.view(np.uint32), z["age_end"].view(np.uint32))
    assert int((meta & 1).sum()) > 300
    head = z["text"].tobytes().split(b"\n")[0] + b"\n"
    for line, died in zip(z["bad"], z["ref_died"]):
        open(p, "wb").write(head + bytes(line))
        with pytest.raises(_lib.ColateError) as e:
            api.read_mut(p)
        assert e.value.code == -5
        # (one case the reference survives by reading past the end of the line -- age_end without its ';' -- is refused as well)
        assert died or bytes(line).count(b";") == 9
    if po.ref_available():      # live: the fixture is what the compiled reference says today
        open(p, "wb").write(z["text"].tobytes())
        rows = po.ref_read_mut(p)
        assert same(rows["pos"], pos) and same(rows["age_begin"].view(np.uint32), ab.view(np.uint32)) and same(po.ref_meta(rows), meta)


def test_make_tmp_from_a_table_matches_the_reference_cli(built, tmp_path):
    """`Colate --mode make_tmp --target_table` (maketmp_table, coal.cpp:2682-2808) -- the .colate.in the host
    writes is byte-identical to the file the reference CLI wrote from the same inputs (fixture maketmp_table.npz from
    make_golden.py maketmp), with and without a target mask; the records then read back through colate_read_colate_in."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden
    z = load("maketmp_table.npz")
    d = str(tmp_path)
    sites = make_golden.maketmp_inputs(d)
    cli = os.path.join(ROOT, "colate_b200", "bin", "Colate")
    for tag, extra in (("nomask", []), ("mask", ["--target_mask", d + "/tm"])):
        r = subprocess.run([cli, "--mode", "make_tmp", "--mut", d + "/syn", "--chr", d + "/chr.txt", "--target_table", d + "/table.txt",
                            "--ref_genome", d + "/refg", "-o", d + "/" + tag] + extra, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        got = open(d + "/" + tag + ".colate.in", "rb").read()
        assert got == z[tag].tobytes(), tag
        rc, bp, aaf, daf, al = api.read_colate_in(d + "/" + tag + ".colate.in", sites.chr_names)
        assert len(bp) > 100 and set(np.unique(aaf + daf)) == {1} and 1 not in set(rc)     # haploid; nothing for the second chromosome
    if po.ref_cli():
        r = subprocess.run([po.ref_cli(), "--mode", "make_tmp", "--mut", d + "/syn", "--chr", d + "/chr.txt", "--target_table", d + "/table.txt",
                            "--ref_genome", d + "/refg", "-o", d + "/live"], capture_output=True, text=True)
        assert r.returncode == 0 and open(d + "/live.colate.in", "rb").read() == z["nomask"].tobytes()


def test_make_tmp_from_a_pileup_matches_the_reference_cli(built, tmp_path):
    """SURVEY.md 8(f) N4, the bam variant on pre-decoded arrays: colate_maketmp_pileup against the .colate.in the reference CLI
    wrote with `--mode make_tmp --target_bam` (maketmp_bam, coal.cpp:2527-2680) from synthetic reads; the pileup handed over is
    the one the reference's own bam_parser held at every row (fixture stage1_bambam.npz, make_golden.py bambam)."""
    from helpers import bambam_masks, sites_from
    z = load("stage1_bambam.npz")
    sites = sites_from(z)
    d = str(tmp_path)
    synth.write_dataset(d, sites, {})
    tm, _ = bambam_masks(z)
    for c, nm in enumerate(sites.chr_names):
        synth.write_mask(os.path.join(d, f"tm_chr{nm}.fa"), tm[c])
    n = len(sites.chr_names)
    arr = lambda xs: (C.c_char_p * n)(*[x.encode() for x in xs])
    names, muts = arr(sites.chr_names), arr([os.path.join(d, f"syn_chr{c}.mut") for c in sites.chr_names])
    cnt = np.ascontiguousarray(z["t_counts"], dtype=np.int32)
    for tag, masks in (("plain", None), ("masked", arr([os.path.join(d, f"tm_chr{c}.fa") for c in sites.chr_names]))):
        out = os.path.join(d, tag + ".colate.in")
        nrec = api.lib().colate_maketmp_pileup(n, names, muts, api.ptr(cnt), cnt.shape[0], masks, out.encode())
        assert nrec > 200, api.lib().colate_last_error()
        assert open(out, "rb").read() == z["maketmp_bam_" + tag].tobytes(), tag
    assert api.lib().colate_maketmp_pileup(n, names, muts, api.ptr(cnt), cnt.shape[0] - 1, None, (d + "/x").encode()) < 0   # row count must match


def test_make_tmp_from_genotype_records_matches_the_reference_cli(built, tmp_path):
    """SURVEY.md 8(f) N4, the vcf variant on pre-decoded arrays: colate_maketmp_records against the .colate.in the reference CLI
    wrote with `--mode make_tmp --target_bcf` (maketmp_vcf, coal.cpp:2325-2525) from synthetic genotype records (fixture
    maketmp_vcf.npz, make_golden.py maketmp_vcf): alleles as the row has them, flipped, a third allele in the record or in a
    genotype, multi-letter alleles, "one allele alone" records, records between rows, a file that ends before the rows do,
    rows read off the reference genome; without and with a target mask."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden
    z = load("maketmp_vcf.npz")
    d = str(tmp_path)
    sites, recs, n_hap = make_golden.maketmp_vcf_inputs(d)
    n = len(sites.chr_names)
    arr = lambda xs: (C.c_char_p * n)(*[x.encode() for x in xs])
    per = lambda stem, ext: arr([os.path.join(d, f"{stem}_chr{c}.{ext}") for c in sites.chr_names])
    code = lambda a: 0 if len(a) == 0 else a[0] if len(a) == 1 else 0xff
    # what a BCF decoder hands over: the arrays of include/colate_b200.h: colate_maketmp_records
    off = np.zeros(n + 1, np.int64)
    off[1:] = np.cumsum([len(r) for r in recs])
    flat = [r for rs in recs for r in rs]
    pos = np.array([p0 + 1 for p0, _, _ in flat], np.int32)
    a0 = np.array([code(al[0]) for _, al, _ in flat], np.uint8)
    a1 = np.array([code(al[1]) for _, al, _ in flat], np.uint8)
    alt = np.array([sum(gt) for _, _, gt in flat], np.int32)
    bi = np.array([max(gt) <= 1 for _, _, gt in flat], np.uint8)
    nh = np.full(n, n_hap, np.int32)
    L = api.lib()
    for tag, masks in (("plain", None), ("masked", per("tm", "fa"))):
        out = os.path.join(d, tag + ".colate.in")
        nrec = L.colate_maketmp_records(n, arr(sites.chr_names), per("syn", "mut"), api.ptr(off), api.ptr(pos), api.ptr(a0), api.ptr(a1),
                                        api.ptr(alt), api.ptr(bi), api.ptr(nh), per("g", "fa"), masks, out.encode())
        assert nrec > 500, L.colate_last_error()
        assert open(out, "rb").read() == z[tag].tobytes(), tag
        rc, bp, aaf, daf, al = api.read_colate_in(out, sites.chr_names)
        assert len(bp) == nrec and set(np.unique(aaf + daf)) == {n_hap} and 0 < (daf > 0).mean() < 1
    # a chromosome without records: the reference reads its first record unconditionally
    off2 = off.copy(); off2[1] = off2[0]
    assert L.colate_maketmp_records(n, arr(sites.chr_names), per("syn", "mut"), api.ptr(off2), api.ptr(pos), api.ptr(a0), api.ptr(a1),
                                    api.ptr(alt), api.ptr(bi), api.ptr(nh), None, None, (d + "/x").encode()) < 0
    if po.ref_cli():
        r = subprocess.run([po.ref_cli(), "--mode", "make_tmp", "--mut", d + "/syn", "--chr", d + "/chr.txt", "--target_bcf", d + "/t",
                            "--ref_genome", d + "/g", "-o", d + "/live"], capture_output=True, text=True)
        assert r.returncode == 0 and open(d + "/live.colate.in", "rb").read() == z["plain"].tobytes()


@pytest.mark.parametrize("seed,ns,pl,p_rec,rows", [(41, 1, 2, 0.5, (700, 300)), (42, 5, 2, 0.95, (400, 900)), (43, 2, 1, 0.2, (1000, 50)),
                                                  (44, 4, 2, 0.7, (60, 60))])
def test_make_tmp_from_genotype_records_fuzz_against_the_live_reference(built, tmp_path, seed, ns, pl, p_rec, rows):
    """colate_maketmp_records against the reference CLI run HERE (oracle/_ref, `--mode make_tmp --target_bcf` through the fake BCF
    reader) on further random datasets: haploid and diploid samples, sparse and dense record files, with and without a mask."""
    if not po.ref_cli():
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden
    d = str(tmp_path)
    sites, recs, n_hap = make_golden.maketmp_vcf_inputs(d, seed=seed, ns=ns, pl=pl, p_rec=p_rec, rows=rows)
    n = len(sites.chr_names)
    arr = lambda xs: (C.c_char_p * n)(*[x.encode() for x in xs])
    per = lambda stem, ext: arr([os.path.join(d, f"{stem}_chr{c}.{ext}") for c in sites.chr_names])
    code = lambda a: 0 if len(a) == 0 else a[0] if len(a) == 1 else 0xff
    off = np.zeros(n + 1, np.int64)
    off[1:] = np.cumsum([len(r) for r in recs])
    flat = [r for rs in recs for r in rs]
    pos = np.array([p0 + 1 for p0, _, _ in flat], np.int32)
    a0 = np.array([code(al[0]) for _, al, _ in flat], np.uint8)
    a1 = np.array([code(al[1]) for _, al, _ in flat], np.uint8)
    alt = np.array([sum(gt) for _, _, gt in flat], np.int32)
    bi = np.array([max(gt) <= 1 for _, _, gt in flat], np.uint8)
    nh = np.full(n, n_hap, np.int32)
    for tag, masks, extra in (("plain", None, []), ("masked", per("tm", "fa"), ["--target_mask", d + "/tm"])):
        r = subprocess.run([po.ref_cli(), "--mode", "make_tmp", "--mut", d + "/syn", "--chr", d + "/chr.txt", "--target_bcf", d + "/t",
                            "--ref_genome", d + "/g", "-o", d + "/ref_" + tag] + extra, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-1000:]
        out = os.path.join(d, tag + ".colate.in")
        nrec = api.lib().colate_maketmp_records(n, arr(sites.chr_names), per("syn", "mut"), api.ptr(off), api.ptr(pos), api.ptr(a0), api.ptr(a1),
                                                api.ptr(alt), api.ptr(bi), api.ptr(nh), per("g", "fa"), masks, out.encode())
        assert nrec >= 0, api.lib().colate_last_error()
        assert open(out, "rb").read() == open(d + "/ref_" + tag + ".colate.in", "rb").read(), (seed, tag)


def test_shipped_objects_use_the_copy_engines(built):
    """The sm_100a objects the library is linked from carry what DESIGN.md says the kernels are built on (cuobjdump -sass, no
    GPU needed): 1-D bulk copies + mbarriers in k_sample and k_replay, a TMA tensor store in k_gen_tma, st.async into peer CTAs
    and cluster barriers in k_em_split, and no tensor-core instruction anywhere (the path has no contraction)."""
    import re
    import shutil
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    csrc = os.path.join(ROOT, "colate_b200", "csrc")

    def kernels(obj):
        out = subprocess.run(["cuobjdump", "-sass", os.path.join(csrc, obj)], capture_output=True, text=True).stdout
        assert "sm_100a" in out, obj
        return {f.split("\n", 1)[0].strip(): f for f in re.split(r"\n\s*Function : ", out)[1:]}

    def body(ks, name):
        hits = [f for m, f in ks.items() if name in m]
        assert hits, name
        return "\n".join(hits)

    sites, mt, em = kernels("kernels_sites.o"), kernels("kernels_mt.o"), kernels("kernels_em.o")
    for k in ("k_sample", "k_replay"):
        assert "UBLKCP" in body(sites, k) and "SYNCS" in body(sites, k), k
    assert "UTMASTG" in body(mt, "k_gen_tma")
    assert "LDS.128" in body(mt, "k_jump")
    assert "STAS" in body(em, "k_em_split") and "UCGABAR" in body(em, "k_em_split")
    for ks in (sites, mt, em):
        for m, f in ks.items():
            assert not re.search(r"\b(HMMA|IMMA|DMMA|UTCHMMA|UTCIMMA|QGMMA)", f), m


def test_mask_bits_from_fasta(built):
    sites = synth.make_sites(3, [400, 300], [3e5, 2e5])
    masks = [synth.make_mask(1, 300000, 0.4, 50, 500), synth.make_mask(2, 100000, 0.3, 50, 500, lower=True)]   # 2nd: short + lower case
    want = api.mask_bits_from_seq([m.upper() for m in masks], sites.site_off, sites.pos)
    bits = np.zeros((sites.n + 31) // 32, np.uint32)
    with tempfile.TemporaryDirectory() as d:
        for c, m in enumerate(masks):
            p = os.path.join(d, f"m_chr{c + 1}.fa")
            synth.write_mask(p, m, width=61)
            lo, hi = int(sites.site_off[c]), int(sites.site_off[c + 1])
            api.mask_bits_from_fasta(p, sites.pos[lo:hi], lo, bits)
    assert same(bits, want)
    beyond = sites.pos[int(sites.site_off[1]):] >= 100000
    idx = np.arange(int(sites.site_off[1]), sites.n)[beyond]
    assert ((bits[idx >> 5] >> (idx & 31)) & 1).all()                  # rows beyond the mask end pass (coal.cpp:2169)


def test_cli_surface(built):
    cli = os.path.join(ROOT, "colate_b200", "bin", "Colate")
    assert os.path.exists(cli)
    r = subprocess.run([cli, "--mode", "mut", "--num_bootstrapz", "3"], capture_output=True, text=True)
    assert r.returncode != 0 and "does not exist" in r.stderr            # cxxopts rejects unknown options
    r = subprocess.run([cli, "--mode", "mut"], capture_output=True, text=True)
    assert r.returncode == 0 and "Not enough arguments supplied." in r.stdout
    r = subprocess.run([cli, "--mode", "nonsense"], capture_output=True, text=True)
    assert "Invalid or missing mode." in r.stdout
    import torch
    if not torch.cuda.is_available():
        with tempfile.TemporaryDirectory() as d:
            r = subprocess.run([cli, "--mode", "mut", "--mut", d + "/x", "--target_tmp", "a", "--reference_tmp", "b", "--bins", "3,7,0.2",
                                "--num_bootstrap", "2", "-o", d + "/o"], capture_output=True, text=True)
            assert r.returncode != 0 and "no CPU fallback" in r.stderr


def test_generator_stream_tile_order(built):
    """internal.h: stream_phys -- where k_gen puts word o of the generator stream so that k_sample can fetch tile t /
    chunk c of 32 used rows as one contiguous block: a bijection onto whole tiles, linear in front of the tiling origin,
    rows of a block 20 words apart, blocks of a tile back to back."""
    from colate_b200 import _lib
    phys = _lib.lib().colate_test_stream_phys
    assert [phys(o, -1) for o in (0, 5, 12345)] == [0, 5, 12345]              # linear mode (test hook / plain stream)
    for off in (0, 200 * 7, 200 * 1000):
        assert all(phys(o, off) == o for o in range(max(0, off - 300), off))  # words before the origin stay in place
        n_rows = 32 * 3                                                         # three whole tiles
        got = np.array([phys(off + q, off) for q in range(200 * n_rows)], dtype=np.int64) - off
        assert np.array_equal(np.sort(got), np.arange(200 * n_rows))           # a permutation of the tiles' words
        q = np.arange(200 * n_rows)
        row, w = q // 200, q % 200
        want = ((row // 32) * 10 + w // 20) * 640 + (row % 32) * 20 + w % 20    # [tile][chunk][row][20 words]
        assert np.array_equal(got, want)
