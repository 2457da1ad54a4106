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
