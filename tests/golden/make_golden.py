#!/usr/bin/env python3
"""Generates tests/golden/*.npz|*.coal from the UNMODIFIED reference compiled by oracle/Makefile
(oracle/_ref/libcolate_ref.so = the reference's own parse_tmptmp / coal_EM behind oracle/ref_probe.cpp,
oracle/_ref/Colate = its CLI).  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

The reference ships no golden vectors for this path; these fixtures are its outputs
on small seeded inputs and pin both the oracle (CPU tests) and the CUDA path (GPU tests)."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from colate_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def dataset(seed, rows, lens, weird, pt=0.7, pr=0.7):
    sites = synth.make_sites(seed, rows, lens, weird=weird)
    gt = synth.make_genome(seed + 100, sites, pt, weird=weird)
    gr = synth.make_genome(seed + 200, sites, pr, weird=weird)
    return sites, gt, gr


def pack(sites, gt, gr):
    d = dict(chr_names=np.array(sites.chr_names), site_off=sites.site_off, pos=sites.pos, age_begin=sites.age_begin,
             age_end=sites.age_end, flipped=sites.flipped, n_branch=sites.n_branch, anc=sites.anc, der=sites.der, odd=sites.odd,
             chrom_len=np.array(sites.chrom_len, dtype=np.int64))
    for nm, g in (("t", gt), ("r", gr)):
        for k in ("chrom", "bp", "anc", "der", "aaf", "daf"):
            d[f"{nm}_{k}"] = getattr(g, k)
    return d


def reject_fixture():
    """parse_tmptmp on rows whose age interval reaches past the age grid (rejection sampling, coal.cpp:2279-2294)."""
    seed = 17
    sites = synth.make_sites(seed, [1400, 1000], [2.4e8, 1.3e8], weird=0.05)
    deep = synth.add_deep_rows(sites, seed + 1, 0.03)
    gt = synth.make_genome(seed + 100, sites, 0.8)
    gr = synth.make_genome(seed + 200, sites, 0.8)
    d = tempfile.mkdtemp()
    synth.write_dataset(d, sites, {"t": gt, "r": gr})
    r = po.ref_parse_tmptmp(d, sites.chr_names, "syn", "t", "r", seed=seed)
    out = pack(sites, gt, gr)
    for k in ("num_blocks", "shared", "notshared", "shared_emp", "notshared_emp", "mt", "next_words"):
        out[f"ref_{k}"] = np.asarray(r[k])
    out["seed"] = seed
    out["deep_rows"] = deep
    np.savez_compressed(os.path.join(OUT, "stage1_reject.npz"), **out)
    print("stage1_reject.npz:", len(deep), "deep rows, blocks", r["num_blocks"])


HEADER = b"snp;pos_of_snp;dist;rs-id;tree_index;branch_indices;is_not_mapping;is_flipped;age_begin;age_end;ancestral_allele/alternative_allele;upstream_allele;downstream_allele;\n"


def mut_reader_cases():
    """(valid text, [invalid single data lines]) for the .mut reader: number spellings std::stoi / std::stof accept in every
    form the device parser has to get right, and lines the reference does not survive."""
    rng = np.random.default_rng(7)
    ages = ["0", "-0", "0.0", "+5", ".5", "5.", "1e3", "1E-3", "1.5e+2", "12345.678", "0.000123", "16777217", "16777219",
            "33554434.000000001", "8388608.5", "8388609.5", "0.1", "0.30000001192092896", "3.4028235e38", "inf", "-inf", "nan",
            "0x1p3", "1e", "1e+", " 12", "12 ", "1.2.3", "123456789012345678", "1234567890123456789012",
            "0.00000000000000000000000000001", "1000000000000000000000000", "9007199254740993", "1.17549435e-38", "  7.5", "\t3",
            "12abc", "1e5x", "4.5e+", "+.5e1", "-1e-3", "00012.5000", "1e0000000002"]
    for _ in range(200):                      # exact float midpoints printed with all their digits, and their neighbours
        f = np.float32(np.exp(rng.uniform(-20, 20)))
        mid = (np.float64(f) + np.float64(np.nextafter(f, np.float32(np.inf)))) / 2
        ages += [format(mid, ".30g"), format(np.nextafter(mid, 0), ".30g"), format(np.nextafter(mid, np.inf), ".25g")]
    for _ in range(3000):
        x = np.exp(rng.uniform(-12, 18))
        ages.append(format(x, rng.choice([".3f", ".6g", ".9g", ".12g", ".17g", "e", ".1f"])))
    lines = [HEADER]
    for i, a in enumerate(ages):
        b = ages[(i * 7 + 3) % len(ages)]
        pos = ["17", " 42", "+9", "-3", "007", "2147483647", "12x"][i % 7]
        flip = ["0", "1", "00", "2", " 0"][i % 5]
        br = ["17", "17 23", " 5", "", "1 2 3"][i % 5]
        typ = ["A/C", "G/T", "AT/C", "A/", "/C", "0/1", "N/A", "A/C extra", "ACGTACGTACGTACGTACGT/A", "A/C"][i % 10]
        tail = ["A;C;\n", "\n", ";;\n", "A;C;10 20 30\n", "A;C;5;6;7;\n"][i % 5]
        lines.append(f"{i};{pos};10;.;5;{br};0;{flip};{a};{b};{typ};".encode() + tail.encode())
    good = b"".join(lines)
    row = lambda **kw: "{snp};{pos};{dist};rs;{tree};{br};0;{fl};{ab};{ae};A/C;A;C;".format(**{**dict(snp=1, pos=5, dist=1, tree=1, br=7, fl=0, ab="1.5", ae="2.5"), **kw})
    bad = [row(ab="1e39"), row(ae="1e-50"), row(ab=""), row(ae="e5"), row(ab="4.9406564584124654e-324"), row(ae="1e400"), row(ab="7e-46"),
           row(pos=""), row(pos="99999999999"), row(pos="x1"), row(snp="a"), row(dist=""), row(tree="t"), row(br="1 x"), row(fl=""),
           row(fl="f"), "1;2;3;.;5;6;0;0;1.0", "9;5;1;.;1;7;0;0;1.5;2.5", row() + "7;;", row() + "x;"]
    return good, [b.encode() + b"\n" for b in bad]


def mut_reader_fixture():
    """Mutations::Read (include/src/mutations.cpp:56-283) itself, through oracle/ref_probe.cpp, on the reader test cases."""
    good, bad = mut_reader_cases()
    d = tempfile.mkdtemp()
    open(d + "/good.mut", "wb").write(good)
    rows = po.ref_read_mut(d + "/good.mut")
    assert rows is not None, "the reference did not survive the valid file"
    died = []
    for i, b in enumerate(bad):
        open(d + f"/bad{i}.mut", "wb").write(HEADER + b)
        died.append(po.ref_read_mut(d + f"/bad{i}.mut") is None)
    print("reference reader:", int(rows["n"]), "rows; invalid cases it died on:", sum(died), "of", len(bad))
    np.savez_compressed(os.path.join(OUT, "mut_reader.npz"), text=np.frombuffer(good, np.uint8), pos=rows["pos"], age_begin=rows["age_begin"],
                        age_end=rows["age_end"], meta=po.ref_meta(rows), bad=np.array(bad), ref_died=np.array(died))


def maketmp_inputs(d):
    """Dataset + haploid table for --mode make_tmp --target_table: entries at most row positions (ancestral / derived / a third
    allele), some at positions that are not rows, a chromosome without entries, the table ending before the last chromosome."""
    z = np.load(os.path.join(OUT, "stage1_small.npz"), allow_pickle=False)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import dataset_from
    sites, gt, gr = dataset_from(z)
    rng = np.random.default_rng(5)
    synth.write_dataset(d, sites, {"t": gt, "r": gr})
    masks = [synth.make_mask(90 + c, int(L) if c != 0 else int(L) // 2, 0.25, lower=(c == 1)) for c, L in enumerate(sites.chrom_len)]
    for c, nm in enumerate(sites.chr_names):
        synth.write_mask(os.path.join(d, f"tm_chr{nm}.fa"), masks[c])
        open(os.path.join(d, f"refg_chr{nm}.fa"), "w").write(">ref\nACGT\n")
    lines = []
    for c, nm in enumerate(sites.chr_names):
        if c == 1:
            continue                                   # no entry at all for the second chromosome
        lo, hi = int(sites.site_off[c]), int(sites.site_off[c + 1])
        if c == 2:
            hi = lo + (hi - lo) * 2 // 3               # the table runs dry inside the last chromosome
        for m in range(lo, hi):
            u = rng.random()
            if u < 0.25:
                continue
            al = chr(sites.anc[m]) if u < 0.55 else chr(sites.der[m]) if u < 0.9 else "ACGT"[int(rng.integers(0, 4))]
            lines.append(f"{nm} {int(sites.pos[m])} {al}")
            if u > 0.97:
                lines.append(f"{nm} {int(sites.pos[m]) + 1} A")      # a position that is not a row (ascending order kept)
    open(os.path.join(d, "table.txt"), "w").write("\n".join(lines) + "\n")
    return sites


def maketmp_fixture():
    """`Colate --mode make_tmp --target_table` (maketmp_table, coal.cpp:2682-2808) by the reference CLI."""
    d = tempfile.mkdtemp()
    maketmp_inputs(d)
    out = {}
    for tag, extra in (("nomask", []), ("mask", ["--target_mask", d + "/tm"])):
        pr = subprocess.run([po.ref_cli(), "--mode", "make_tmp", "--mut", d + "/syn", "--chr", d + "/chr.txt", "--target_table", d + "/table.txt",
                             "--ref_genome", d + "/refg", "-o", d + "/" + tag] + extra, capture_output=True, text=True)
        assert pr.returncode == 0, pr.stderr
        out[tag] = np.frombuffer(open(d + "/" + tag + ".colate.in", "rb").read(), np.uint8)
        print("make_tmp", tag, out[tag].shape[0], "bytes")
    np.savez_compressed(os.path.join(OUT, "maketmp_table.npz"), **out)


def n2_fixture():
    """the reference CLI (a) started from a <out>.colate_mat cache (coal.cpp:3169-3170, 3471-3499: parsing is
    skipped, the counts are read back from 6-digit text) and (b) warm-started from a .coal file (--coal, coal.cpp:3508-3549,
    3638-3646) on the cli_small dataset."""
    z = np.load(os.path.join(OUT, "cli_small.npz"), allow_pickle=False)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import dataset_from
    sites, gt, gr = dataset_from(z)
    seed = int(z["seed"])
    d = tempfile.mkdtemp()
    synth.write_dataset(d, sites, {"t": gt, "r": gr})
    # (a) a cache in the reference's own format (coal.cpp:3336-3343, 3453-3465): age grid line, then per replicate the shared
    # and the not-shared line, operator<< with default precision (== %g)
    o = po.stage1(sites, gt, gr, seed=seed)
    w = po.draw_block_weights(o["rng"], 3, o["num_blocks"])
    counts = po.stage2(w, o, 0.0)
    lines = [" ".join("%g" % v for v in po.age_bins()) + " "]
    for r in range(3):
        lines += [" ".join("%g" % v for v in counts[r, 0]) + " ", " ".join("%g" % v for v in counts[r, 1]) + " "]
    mat = "\n".join(lines) + "\n"
    open(os.path.join(OUT, "n2_cache.colate_mat"), "w").write(mat)
    open(d + "/cached.colate_mat", "w").write(mat)
    base = [po.ref_cli(), "--mode", "mut", "--mut", d + "/syn", "--chr", d + "/chr.txt", "--target_tmp", d + "/t.colate.in",
            "--reference_tmp", d + "/r.colate.in", "--seed", str(seed)]
    pr = subprocess.run(base + ["--bins", "3,7,0.2", "--num_bootstraps", "3", "-o", d + "/cached"], capture_output=True, text=True)
    assert pr.returncode == 0 and "Loading precomputed file" in pr.stderr, pr.stderr
    open(os.path.join(OUT, "n2_cache.coal"), "w").write(open(d + "/cached.coal").read())
    # (b) warm start from the golden .coal of the ancient run (a non-ancient .coal starts "0 0 ..." and trips the
    # reference's own assert(epochs[e] > epochs[e-1]) at coal.cpp:3548)
    pr = subprocess.run(base + ["--coal", os.path.join(OUT, "cli_ancient.coal"), "--target_age", "7000", "--reference_age", "0",
                                "--years_per_gen", "28", "-o", d + "/warm"], capture_output=True, text=True)
    assert pr.returncode == 0, pr.stderr
    open(os.path.join(OUT, "n2_warm.coal"), "w").write(open(d + "/warm.coal").read())
    print("n2 fixtures written")


def bambam_reads(rng, sites, genome, p_cov, mean_extra):
    """Synthetic aligned reads around the .mut rows of every chromosome, sorted by (contig, start): lengths 24..80 (some under the
    length filter), mapping qualities some under the filter, base qualities some under 30, up to a dozen mismatches against the
    reference genome (some reads over the mismatch filter), rows at any offset of a read (the first / last three bases of a read
    never count, htslib.cpp:66), the row's allele ancestral / derived / a third base."""
    reads = []
    for c in range(len(sites.chr_names)):
        g = genome[c]
        L = g.shape[0]
        lo, hi = int(sites.site_off[c]), int(sites.site_off[c + 1])
        per = []
        for m in range(lo, hi):
            if rng.random() >= p_cov:
                continue
            p0 = int(sites.pos[m]) - 1
            anc, der = int(sites.anc[m]), int(sites.der[m])
            for _ in range(1 + rng.poisson(mean_extra)):
                ln = int(rng.integers(24, 81))
                off = int(rng.integers(0, ln))
                st = p0 - off
                if st < 0 or st + ln > L:
                    continue
                seq = g[st:st + ln].copy()
                u = rng.random()
                alt = [b for b in b"ACGT" if b not in (anc, der)]
                seq[off] = der if u < 0.35 else (anc if u < 0.9 and anc in b"ACGT" else alt[int(rng.integers(0, len(alt)))])
                nmm = int(rng.choice([0, 0, 0, 1, 2, 4, 9, 12]))
                for k in rng.integers(0, ln, size=nmm):
                    if k != off:
                        seq[k] = b"ACGT"[(b"ACGT".index(bytes([seq[k]])) + 1) % 4] if bytes([seq[k]]) in b"ACGT" else seq[k]
                qual = np.where(rng.random(ln) < 0.1, 12, 37).astype(np.uint8)
                mapq = int(rng.choice([0, 10, 19, 20, 37, 60, 60, 60]))
                per.append((c, st, mapq, bool(rng.random() < 0.5), seq.tobytes(), qual))
        per.sort(key=lambda r: r[1])
        reads += per
    return reads


def bambam_inputs(seed=23):
    """The synthetic bam/bam dataset of the fixture, deterministically from its seed: rows, chromosome lengths, the reference
    genome per chromosome (uint8 ACGT) and the aligned reads of the target and the reference sample (tests regenerate them to
    drive the device pileup with the very reads the reference's bam_parser saw)."""
    rng = np.random.default_rng(seed)
    lens = [65_000_000, 31_000_000]
    sites = synth.make_sites(seed, [1300, 700], lens, weird=0.1)
    genome = []
    for c, L in enumerate(lens):
        g = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, L)].copy()
        lo, hi = int(sites.site_off[c]), int(sites.site_off[c + 1])
        for m in range(lo, hi):                      # the reference genome mostly carries the ancestral allele at a row
            a = int(sites.anc[m])
            if a in b"ACGT" and rng.random() < 0.85:
                g[int(sites.pos[m]) - 1] = a
        genome.append(g)
    reads_t = bambam_reads(rng, sites, genome, 0.85, 1.2)
    reads_r = bambam_reads(rng, sites, genome, 0.9, 2.0)
    return sites, lens, genome, reads_t, reads_r


def bambam_fixture():
    """the bam/bam front-end.  The reference's own parse_onebambam (coal.cpp:1799-2069) and bam_parser
    (include/vcf/htslib.cpp) run on synthetic reads served by oracle/hts_stubs.c; the fixture keeps the rows, the pileup the
    reference's bam_parser holds at every row (the pre-decoded arrays colate_set_pileup takes) and parse_onebambam's outputs,
    with and without masks."""
    seed = 23
    sites, lens, genome, reads_t, reads_r = bambam_inputs(seed)
    d = tempfile.mkdtemp()
    synth.write_dataset(d, sites, {})
    for c, L in enumerate(lens):
        g = genome[c]
        with open(os.path.join(d, f"g_chr{sites.chr_names[c]}.fa"), "wb") as f:
            f.write(b">ref\n")
            for i in range(0, L, 1 << 20):
                f.write(g[i:i + (1 << 20)].tobytes() + b"\n")
    bam_names = ["chr" + sites.chr_names[0], sites.chr_names[1]]      # bam_parser accepts "<name>" and "chr<name>" (htslib.cpp:388)
    po.write_fake_bam(os.path.join(d, "t.bam"), bam_names, reads_t)
    po.write_fake_bam(os.path.join(d, "r.bam"), bam_names, reads_r)
    masks = {"tm": [synth.make_mask(seed * 10 + c, int(L) if c else int(L) // 2, 0.25) for c, L in enumerate(lens)],
             "rm": [synth.make_mask(seed * 20 + c, int(L), 0.15) for c, L in enumerate(lens)]}
    for mname, per_chr in masks.items():
        for c, nm in enumerate(sites.chr_names):
            synth.write_mask(os.path.join(d, f"{mname}_chr{nm}.fa"), per_chr[c])
    out = dict(chr_names=np.array(sites.chr_names), site_off=sites.site_off, pos=sites.pos, age_begin=sites.age_begin,
               age_end=sites.age_end, flipped=sites.flipped, n_branch=sites.n_branch, anc=sites.anc, der=sites.der, odd=sites.odd,
               chrom_len=np.array(lens, dtype=np.int64), seed=seed)
    for tag, bam in (("t", "t.bam"), ("r", "r.bam")):
        cnt = np.zeros((sites.n, 4), np.int32)
        for c, nm in enumerate(sites.chr_names):
            lo, hi = int(sites.site_off[c]), int(sites.site_off[c + 1])
            cnt[lo:hi], _ = po.ref_bam_pileup(os.path.join(d, bam), nm, os.path.join(d, f"g_chr{nm}.fa"), sites.pos[lo:hi])
        out[f"{tag}_counts"] = cnt
    for tag, tm, rm in (("plain", None, None), ("masked", "tm", "rm")):
        r = po.ref_parse_onebambam(d, sites.chr_names, "syn", "t.bam", "r.bam", "g", seed=seed, tmask=tm, rmask=rm)
        assert r["emp_rest"].sum() == 0
        for k in ("num_blocks", "shared", "notshared", "shared_emp", "notshared_emp", "mt"):
            out[f"ref_{tag}_{k}"] = np.asarray(r[k])
        print("stage1_bambam.npz", tag, "blocks", r["num_blocks"], "sum shared", r["shared"].sum(), "notshared", r["notshared"].sum())
    # (the masks are regenerated by the tests: synth.make_mask(seed * 10 + c, len or len // 2 for c == 0, 0.25) / (seed * 20 + c, len, 0.15))
    # the whole CLI on the same inputs: <out>.colate_mat (coal.cpp:3336-3343, 3453-3465: counts / 1e3, six digits) and <out>.coal
    for tag, extra in (("bambam_R1", []), ("bambam_R3", ["--num_bootstraps", "3"])):
        o = os.path.join(d, tag)
        cmd = [po.ref_cli(), "--mode", "mut", "--mut", os.path.join(d, "syn"), "--chr", os.path.join(d, "chr.txt"), "--target_bam", os.path.join(d, "t.bam"),
               "--reference_bam", os.path.join(d, "r.bam"), "--ref_genome", os.path.join(d, "g"), "--bins", "3,7,0.2", "--seed", str(seed), "-o", o] + extra
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        for ext in (".coal", ".colate_mat"):
            with open(o + ext) as f, open(os.path.join(OUT, tag + ext), "w") as g:
                g.write(f.read())
        print(tag, "written:", [ln for ln in r.stderr.splitlines() if "blocks" in ln or "iterations" in ln][:4])
    # --mode make_tmp --target_bam (maketmp_bam, coal.cpp:2527-2680) on the target's reads: the .colate.in bytes
    for tag, extra in (("plain", []), ("masked", ["--target_mask", os.path.join(d, "tm")])):
        o = os.path.join(d, "mk_" + tag)
        r = subprocess.run([po.ref_cli(), "--mode", "make_tmp", "--mut", os.path.join(d, "syn"), "--chr", os.path.join(d, "chr.txt"),
                            "--target_bam", os.path.join(d, "t.bam"), "--ref_genome", os.path.join(d, "g"), "-o", o] + extra, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        out["maketmp_bam_" + tag] = np.frombuffer(open(o + ".colate.in", "rb").read(), np.uint8)
        print("make_tmp --target_bam", tag, out["maketmp_bam_" + tag].shape[0], "bytes")
    np.savez_compressed(os.path.join(OUT, "stage1_bambam.npz"), **out)
    print("covered rows: target", int((out["t_counts"].sum(1) > 0).sum()), "reference", int((out["r_counts"].sum(1) > 0).sum()), "of", sites.n)
    import shutil
    shutil.rmtree(d, ignore_errors=True)


def vcf_records(rng, sites, c, n_hap, p_rec, ref_side):
    """Synthetic genotype records of one sample set on chromosome c: at most row positions (alleles as the row has them, flipped,
    a third allele in the record or in a genotype, unrelated alleles, multi-letter alleles; for the target also records that
    report a single allele) and at positions that are not rows."""
    lo, hi = int(sites.site_off[c]), int(sites.site_off[c + 1])
    recs = {}
    for m in range(lo, hi):
        if rng.random() >= p_rec:
            continue
        anc, der = bytes([int(sites.anc[m])]), bytes([int(sites.der[m])])
        third = bytes([[b for b in b"ACGT" if bytes([b]) not in (anc, der)][0]])
        u = rng.random()
        if u < 0.55:
            al = [anc, der]
        elif u < 0.75:
            al = [der, anc]
        elif u < 0.80:
            al = [anc, third]
        elif u < 0.85:
            al = [anc, der, third]
        elif u < 0.90:
            al = [anc + b"T", der]
        elif u < 0.95 and not ref_side:
            al = [anc] if rng.random() < 0.5 else [der]
        else:
            al = [third, anc]
        if len(al) == 1:
            gt = [0] * n_hap if rng.random() < 0.8 else [1] + [0] * (n_hap - 1)      # "alt not reported": only sensible if all-reference
        else:
            gt = list((rng.random(n_hap) < 0.4).astype(int))
            if len(al) == 3 and rng.random() < 0.6:
                gt[int(rng.integers(0, n_hap))] = 2
        recs[int(sites.pos[m]) - 1] = (al, gt)
    L = int(sites.chrom_len[c])
    for p in rng.integers(0, L, size=max(4, (hi - lo) // 5)):            # records between the rows
        recs.setdefault(int(p), ([b"A", b"G"], list((rng.random(n_hap) < 0.5).astype(int))))
    return [(p, recs[p][0], recs[p][1]) for p in sorted(recs)]


def vcfvcf_fixture():
    """SURVEY.md 8(f) N3, the bcf/bcf front-end: the reference's own parse_vcfvcf (coal.cpp:907-1228) and vcf_parser run on
    synthetic genotype records served by oracle/hts_stubs.c, without and with --ref_genome (+ masks).  The fixture keeps the
    rows, the records, the reference genome's base at every row and parse_vcfvcf's outputs."""
    seed = 29
    rng = np.random.default_rng(seed)
    lens = [65_000_000, 31_000_000]
    sites = synth.make_sites(seed, [1500, 900], lens, weird=0.08)
    d = tempfile.mkdtemp()
    synth.write_dataset(d, sites, {})
    n_t, pl_t, n_r, pl_r = 1, 2, 3, 2
    out = dict(chr_names=np.array(sites.chr_names), site_off=sites.site_off, pos=sites.pos, age_begin=sites.age_begin,
               age_end=sites.age_end, flipped=sites.flipped, n_branch=sites.n_branch, anc=sites.anc, der=sites.der, odd=sites.odd,
               chrom_len=np.array(lens, dtype=np.int64), seed=seed, n_hap_target=n_t * pl_t, n_hap_ref=n_r * pl_r)
    refg_rows = np.zeros(sites.n, np.uint8)
    for c, nm in enumerate(sites.chr_names):
        for tag, ns, pl, p_rec in (("t", n_t, pl_t, 0.75), ("r", n_r, pl_r, 0.85)):
            recs = vcf_records(rng, sites, c, ns * pl, p_rec, tag == "r")
            po.write_fake_bcf(os.path.join(d, f"{tag}_chr{nm}.bcf"), ns, pl, recs)
            out[f"{tag}{c}_pos"] = np.array([r[0] for r in recs], np.int32)
            out[f"{tag}{c}_nal"] = np.array([len(r[1]) for r in recs], np.int32)
            out[f"{tag}{c}_al"] = np.array([[a.ljust(3, b"\0") for a in (r[1] + [b""] * 3)[:3]] for r in recs], dtype="S3")
            out[f"{tag}{c}_gt"] = np.array([r[2] for r in recs], np.int8)
        # reference genome: derived / ancestral / another base at the rows
        L = lens[c]
        g = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, L)].copy()
        lo, hi = int(sites.site_off[c]), int(sites.site_off[c + 1])
        for m in range(lo, hi):
            u = rng.random()
            if u < 0.45 and int(sites.der[m]) in b"ACGT":
                g[int(sites.pos[m]) - 1] = int(sites.der[m])
            elif u < 0.85 and int(sites.anc[m]) in b"ACGT":
                g[int(sites.pos[m]) - 1] = int(sites.anc[m])
        refg_rows[lo:hi] = g[sites.pos[lo:hi] - 1]
        with open(os.path.join(d, f"g_chr{nm}.fa"), "wb") as f:
            f.write(b">ref\n")
            for i in range(0, L, 1 << 20):
                f.write(g[i:i + (1 << 20)].tobytes() + b"\n")
    out["refg_at_rows"] = refg_rows
    masks = {"tm": [synth.make_mask(seed * 10 + c, int(L) if c else int(L) // 2, 0.25) for c, L in enumerate(lens)],
             "rm": [synth.make_mask(seed * 20 + c, int(L), 0.15) for c, L in enumerate(lens)]}
    for mname, per_chr in masks.items():
        for c, nm in enumerate(sites.chr_names):
            synth.write_mask(os.path.join(d, f"{mname}_chr{nm}.fa"), per_chr[c])
    for tag, kw in (("plain", {}), ("refg", dict(ref_genome="g")), ("refg_masked", dict(ref_genome="g", tmask="tm", rmask="rm"))):
        r = po.ref_parse_vcfvcf(d, sites.chr_names, "syn", "t", "r", seed=seed, **kw)
        assert r["emp_rest"].sum() == 0
        for k in ("num_blocks", "shared", "notshared", "shared_emp", "notshared_emp", "mt"):
            out[f"ref_{tag}_{k}"] = np.asarray(r[k])
        print("stage1_vcfvcf.npz", tag, "blocks", r["num_blocks"], "sum shared", r["shared"].sum(), "notshared", r["notshared"].sum())
    np.savez_compressed(os.path.join(OUT, "stage1_vcfvcf.npz"), **out)
    import shutil
    shutil.rmtree(d, ignore_errors=True)


MAKETMP_VCF_LENS = [3_000_000, 2_000_000]


def maketmp_vcf_inputs(d, seed=31, ns=3, pl=2, p_rec=0.8, rows=(1500, 900)):
    """Dataset, genotype records of a 3-sample diploid target (fake BCFs), reference genome and target mask for
    `--mode make_tmp --target_bcf`: records as vcf_records makes them, some rewritten to "the derived / ancestral allele alone"
    (second allele the empty string: coal.cpp:2413), the first chromosome's records ending before its last rows."""
    rng = np.random.default_rng(seed)
    lens = MAKETMP_VCF_LENS
    sites = synth.make_sites(seed, list(rows), lens, weird=0.08)
    synth.write_dataset(d, sites, {})
    recs_all = []
    for c, nm in enumerate(sites.chr_names):
        recs = vcf_records(rng, sites, c, ns * pl, p_rec, True)
        lo, hi = int(sites.site_off[c]), int(sites.site_off[c + 1])
        by_pos = {int(sites.pos[m]) - 1: m for m in range(lo, hi)}
        out = []
        for p0, al, gt in recs:
            if c == 0 and p0 > int(sites.pos[hi - 40]):
                break                                          # the reader runs off the end before the rows do
            m = by_pos.get(p0)
            if m is not None and rng.random() < 0.12:
                first = bytes([int(sites.der[m])]) if rng.random() < 0.7 else bytes([int(sites.anc[m])])
                al = [first, b""]
                gt = [0] * (ns * pl) if rng.random() < 0.75 else [1] + [0] * (ns * pl - 1)
            out.append((p0, al, gt))
        po.write_fake_bcf(os.path.join(d, f"t_chr{nm}.bcf"), ns, pl, out)
        recs_all.append(out)
        L = lens[c]
        g = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, L)].copy()
        for m in range(lo, hi):
            u = rng.random()
            if u < 0.45 and int(sites.der[m]) in b"ACGT":
                g[int(sites.pos[m]) - 1] = int(sites.der[m])
            elif u < 0.85 and int(sites.anc[m]) in b"ACGT":
                g[int(sites.pos[m]) - 1] = int(sites.anc[m])
        with open(os.path.join(d, f"g_chr{nm}.fa"), "wb") as f:
            f.write(b">ref\n")
            for i in range(0, L, 60000):
                f.write(g[i:i + 60000].tobytes() + b"\n")
        synth.write_mask(os.path.join(d, f"tm_chr{nm}.fa"), synth.make_mask(seed * 10 + c, L if c else L // 2, 0.25, run_lo=200, run_hi=20000))
    return sites, recs_all, ns * pl


def maketmp_vcf_fixture():
    """SURVEY.md 8(f) N4, the vcf variant: `Colate --mode make_tmp --target_bcf` (maketmp_vcf, coal.cpp:2325-2525) by the reference
    CLI on synthetic genotype records served by oracle/hts_stubs.c, without and with a target mask."""
    d = tempfile.mkdtemp()
    maketmp_vcf_inputs(d)
    out = {}
    for tag, extra in (("plain", []), ("masked", ["--target_mask", d + "/tm"])):
        pr = subprocess.run([po.ref_cli(), "--mode", "make_tmp", "--mut", d + "/syn", "--chr", d + "/chr.txt", "--target_bcf", d + "/t",
                             "--ref_genome", d + "/g", "-o", d + "/" + tag] + extra, capture_output=True, text=True)
        assert pr.returncode == 0, pr.stderr[-2000:]
        out[tag] = np.frombuffer(open(d + "/" + tag + ".colate.in", "rb").read(), np.uint8)
        print("make_tmp --target_bcf", tag, out[tag].shape[0], "bytes")
    np.savez_compressed(os.path.join(OUT, "maketmp_vcf.npz"), **out)
    import shutil
    shutil.rmtree(d, ignore_errors=True)


def main():
    assert po.ref_available() and po.ref_cli(), "build oracle/_ref first (make -C oracle ref)"
    if len(sys.argv) > 1 and sys.argv[1] == "bambam":
        return bambam_fixture()
    if len(sys.argv) > 1 and sys.argv[1] == "vcfvcf":
        return vcfvcf_fixture()
    if len(sys.argv) > 1 and sys.argv[1] == "reject":
        return reject_fixture()
    if len(sys.argv) > 1 and sys.argv[1] == "n2":
        return n2_fixture()
    if len(sys.argv) > 1 and sys.argv[1] == "mut_reader":
        return mut_reader_fixture()
    if len(sys.argv) > 1 and sys.argv[1] == "maketmp":
        return maketmp_fixture()
    if len(sys.argv) > 1 and sys.argv[1] == "maketmp_vcf":
        return maketmp_vcf_fixture()
    bambam_fixture()
    vcfvcf_fixture()
    maketmp_fixture()
    maketmp_vcf_fixture()
    mut_reader_fixture()
    n2_fixture()
    reject_fixture()
    # ---- stage i: parse_tmptmp on weird rows, with and without masks
    seed = 11
    sites, gt, gr = dataset(seed, [1500, 900, 1200], [2.4e8, 6.1e7, 1.3e8], weird=0.12)
    masks = {"tm": [synth.make_mask(seed * 10 + c, int(L) if c != 1 else int(L) // 2, 0.3, lower=(c == 2)) for c, L in enumerate(sites.chrom_len)],
             "rm": [synth.make_mask(seed * 20 + c, int(L), 0.2) for c, L in enumerate(sites.chrom_len)]}
    d = tempfile.mkdtemp()
    synth.write_dataset(d, sites, {"t": gt, "r": gr}, masks)
    out = pack(sites, gt, gr)
    # masks are stored as pass bits at the site positions (the full sequences are ~400 MB)
    from colate_b200 import api
    up = lambda ms: [bytes(m).upper() for m in ms]
    out["tmask_bits"] = api.mask_bits_from_seq(up(masks["tm"]), sites.site_off, sites.pos)
    out["rmask_bits"] = api.mask_bits_from_seq(up(masks["rm"]), sites.site_off, sites.pos)
    for tag, tm, rm in (("nomask", None, None), ("mask", "tm", "rm")):
        r = po.ref_parse_tmptmp(d, sites.chr_names, "syn", "t", "r", seed=seed, tmask=tm, rmask=rm)
        assert r["emp_rest"].sum() == 0
        for k in ("num_blocks", "shared", "notshared", "shared_emp", "notshared_emp", "mt", "next_words"):
            out[f"ref_{tag}_{k}"] = np.asarray(r[k])
    out["seed"] = seed
    np.savez_compressed(os.path.join(OUT, "stage1_small.npz"), **out)

    # ---- whole path: reference CLI .coal for three flag sets + E-step vectors
    seed = 1
    sites, gt, gr = dataset(seed, [30000, 20000], [2.5e8, 1.2e8], weird=0.02)
    d = tempfile.mkdtemp()
    synth.write_dataset(d, sites, {"t": gt, "r": gr})
    cli = pack(sites, gt, gr)
    cases = {"bins02_R1": ["--bins", "3,7,0.2"], "bins02_R3": ["--bins", "3,7,0.2", "--num_bootstraps", "3"],
             "ancient": ["--bins", "3,7,0.2", "--target_age", "7000", "--reference_age", "0", "--years_per_gen", "28"],
             "bins01_R1": ["--bins", "3,7,0.1"]}
    for name, extra in cases.items():
        cmd = [po.ref_cli(), "--mode", "mut", "--mut", d + "/syn", "--chr", d + "/chr.txt", "--target_tmp", d + "/t.colate.in",
               "--reference_tmp", d + "/r.colate.in", "--seed", str(seed), "-o", d + "/" + name] + extra
        pr = subprocess.run(cmd, capture_output=True, text=True)
        assert pr.returncode == 0, pr.stderr
        txt = open(d + "/" + name + ".coal").read()
        open(os.path.join(OUT, f"cli_{name}.coal"), "w").write(txt)
        nb = [ln for ln in pr.stderr.split("\n") if ln.startswith("Number of blocks")]
        cli[f"{name}_num_blocks"] = int(nb[0].split(":")[1])
        cli[f"{name}_args"] = np.array(extra)
    cli["seed"] = seed
    np.savez_compressed(os.path.join(OUT, "cli_small.npz"), **cli)

    # ---- E-step known answers from the reference's coal_EM (the object its unit test exercises)
    ab = po.age_bins()
    rng = np.random.default_rng(0)
    est = {}
    for tag, bins, age in (("b02", "3,7,0.2", 0.0), ("b01", "3,7,0.1", 0.0), ("anc", "3,7,0.2", 250.0)):
        ep, null = po.epochs_from_bins(bins, age, 28.0)
        E = len(ep)
        rates = np.stack([np.full(E, 1 / 20000.), np.exp(rng.uniform(np.log(1e-7), np.log(1e-2), E)), np.full(E, 1e-7), np.full(E, 1e-1)])
        rates[1, 2] = 0.0
        num = np.zeros((4, 2, 185, E)); den = np.zeros_like(num); ll = np.zeros((4, 2, 185))
        for i in range(4):
            for s in (0, 1):
                for b in range(185):
                    ll[i, s, b], num[i, s, b], den[i, s, b] = po.ref_estep(s == 0, ep, rates[i], ab[b])
        est.update({f"{tag}_epochs": ep, f"{tag}_ep_null": null, f"{tag}_rates": rates, f"{tag}_num": num, f"{tag}_den": den, f"{tag}_ll": ll})
    np.savez_compressed(os.path.join(OUT, "estep_ref.npz"), **est)

    # ---- <random>: libstdc++ outputs
    w = np.zeros(4000, np.uint32); po.ref().ref_mt_words(1, 0, 4000, w)
    w2 = np.zeros(1000, np.uint32); po.ref().ref_mt_words(123456789, 10**6, 1000, w2)
    u = np.zeros(2000); po.ref().ref_uniform_real(7, 3, 2000, u)
    ints = {}
    for nb in (1, 9, 105, 500):
        a = np.zeros(3000, np.int32); po.ref().ref_uniform_int(3, 11, nb, 3000, a); ints[f"int_{nb}"] = a
    np.savez_compressed(os.path.join(OUT, "random_ref.npz"), words_seed1=w, words_seed123456789_skip1e6=w2, real_seed7_skip3=u, **ints)
    print("golden fixtures written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print("  ", f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
