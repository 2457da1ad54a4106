"""Shared test helpers: golden fixtures -> synth objects."""
import os

import numpy as np

from colate_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def dataset_from(z):
    sites = synth.Sites([str(x) for x in z["chr_names"]], z["site_off"], z["pos"], z["age_begin"], z["age_end"], z["flipped"],
                        z["n_branch"], z["anc"], z["der"], z["odd"], [int(x) for x in z["chrom_len"]])
    gs = []
    for nm in ("t", "r"):
        gs.append(synth.Genome(*[z[f"{nm}_{k}"] for k in ("chrom", "bp", "anc", "der", "aaf", "daf")]))
    return sites, gs[0], gs[1]


def same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def sites_from(z):
    """Fixtures without .colate.in genomes (the pileup front-end, stage1_bambam.npz)."""
    return synth.Sites([str(x) for x in z["chr_names"]], z["site_off"], z["pos"], z["age_begin"], z["age_end"], z["flipped"],
                       z["n_branch"], z["anc"], z["der"], z["odd"], [int(x) for x in z["chrom_len"]])


def bambam_masks(z):
    """The masks tests/golden/make_golden.py::bambam_fixture used (regenerated from their seeds)."""
    seed = int(z["seed"])
    lens = [int(x) for x in z["chrom_len"]]
    tm = [synth.make_mask(seed * 10 + c, L if c else L // 2, 0.25) for c, L in enumerate(lens)]
    rm = [synth.make_mask(seed * 20 + c, L, 0.15) for c, L in enumerate(lens)]
    return tm, rm


def vcfvcf_counts(z, with_refg):
    """Per-row (AAF, DAF) of the vcfvcf fixture for both genomes: the decoder half of parse_vcfvcf (oracle/pyoracle.py:
    decode_vcfvcf) on the fixture's genotype records."""
    from oracle import pyoracle as po
    sites = sites_from(z)
    meta = sites.meta()
    tc = np.zeros((sites.n, 2), np.int32)
    rc = np.zeros((sites.n, 2), np.int32)
    for c in range(len(sites.chr_names)):
        lo, hi = int(sites.site_off[c]), int(sites.site_off[c + 1])
        recs = {}
        for tag in ("t", "r"):
            al = z[f"{tag}{c}_al"]
            recs[tag] = [(int(p), [bytes(a) for a in al[i][:int(n)]], [int(g) for g in gt])
                         for i, (p, n, gt) in enumerate(zip(z[f"{tag}{c}_pos"], z[f"{tag}{c}_nal"], z[f"{tag}{c}_gt"]))]
        usable = ((meta[lo:hi] & 1) != 0) & (sites.anc[lo:hi] != sites.der[lo:hi])          # coal.cpp:966-995
        tc[lo:hi], rc[lo:hi] = po.decode_vcfvcf(sites.pos[lo:hi], sites.anc[lo:hi], sites.der[lo:hi], usable, recs["t"], recs["r"],
                                                 int(z["n_hap_target"]), int(z["n_hap_ref"]), z["refg_at_rows"][lo:hi] if with_refg else None)
    return sites, tc, rc
