"""GPU: the `Colate` CLI host end to end (files -> readers -> C-ABI -> .coal/.bin) against the
.coal files the UNMODIFIED reference CLI wrote for the same inputs (tests/golden/cli_*.coal)."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from colate_b200 import api, synth
from oracle import pyoracle as po
from helpers import GOLDEN, dataset_from, load

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "colate_b200", "bin", "Colate")


@pytest.fixture(scope="module")
def dataset_dir():
    z = load("cli_small.npz")
    sites, gt, gr = dataset_from(z)
    d = tempfile.mkdtemp(prefix="colate_cli_")
    synth.write_dataset(d, sites, {"t": gt, "r": gr})
    return d, z, sites, gt, gr


@pytest.mark.parametrize("name", ["bins02_R1", "bins02_R3", "ancient", "bins01_R1"])
def test_cli_coal_identical_to_reference(dataset_dir, name):
    d, z, sites, gt, gr = dataset_dir
    extra = [str(x) for x in z[f"{name}_args"]]
    if name == "bins02_R3":
        extra = [("--num_bootstrap" if x == "--num_bootstraps" else x) for x in extra]   # the README's spelling is accepted too
    out = os.path.join(d, "gpu_" + name)
    cmd = [CLI, "--mode", "mut", "--mut", d + "/syn", "--chr", d + "/chr.txt", "--target_tmp", d + "/t.colate.in",
           "--reference_tmp", d + "/r.colate.in", "--seed", str(int(z["seed"])), "-o", out] + extra
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert f"Number of blocks: {int(z[f'{name}_num_blocks'])}" in r.stderr
    assert open(out + ".coal").read() == open(os.path.join(GOLDEN, f"cli_{name}.coal")).read()
    # the fp64 side output equals the oracle bit for bit (the .coal text only carries 6 digits)
    raw = open(out + ".bin", "rb").read()
    R, E = np.frombuffer(raw[8:16], np.int32)
    rates = np.frombuffer(raw[16 + 8 * E:16 + 8 * E + 8 * R * E], np.float64).reshape(R, E)
    iters = np.frombuffer(raw[16 + 8 * E + 8 * R * E:], np.int32)
    opt = dict(zip(extra[::2], extra[1::2]))
    o = po.stage1(sites, gt, gr, seed=int(z["seed"]))
    w = po.draw_block_weights(o["rng"], R, o["num_blocks"])
    age, ypg = po.ages(opt.get("--target_age"), opt.get("--reference_age"), float(opt["--years_per_gen"]) if "--years_per_gen" in opt else None)
    counts = po.stage2(w, o, age)
    ep, null = po.epochs_from_bins(opt["--bins"], age, ypg)
    for i in range(R):
        ro, it, _ = po.em_run(ep, np.full(len(ep), 1 / 20000.), counts[i])
        if age > 0:
            ro = ro.copy(); ro[:null + 1] = 0          # zeroed at output time, coal.cpp:3832-3834
        assert it == iters[i]
        assert np.array_equal(rates[i], ro)


def test_cli_generator_fallback_without_tma_is_byte_identical(dataset_dir):
    """The generator's scatter through the TMA unit (k_gen_tma, kernels_mt.cu) and its fallback that computes every word's place
    itself (k_gen<true>: COLATE_GEN_NO_TMA=1, also taken when cuTensorMapEncodeTiled is unavailable) write the same stream: the
    reference CLI's golden .coal either way, and the fp64 side output equal byte for byte."""
    d, z, sites, gt, gr = dataset_dir
    name = "bins02_R3"
    extra = [str(x) for x in z[f"{name}_args"]]
    outs = []
    for tag, env in (("tma", {}), ("notma", {"COLATE_GEN_NO_TMA": "1"})):
        out = os.path.join(d, "gen_" + tag)
        cmd = [CLI, "--mode", "mut", "--mut", d + "/syn", "--chr", d + "/chr.txt", "--target_tmp", d + "/t.colate.in",
               "--reference_tmp", d + "/r.colate.in", "--seed", str(int(z["seed"])), "-o", out] + extra
        r = subprocess.run(cmd, capture_output=True, text=True, env={**os.environ, **env})
        assert r.returncode == 0, r.stderr
        assert open(out + ".coal").read() == open(os.path.join(GOLDEN, f"cli_{name}.coal")).read(), tag
        outs.append(open(out + ".bin", "rb").read())
    assert outs[0] == outs[1]


def test_cli_masks_and_cache(dataset_dir):
    d, z, sites, gt, gr = dataset_dir
    # masks: per-chromosome fasta files; compare stage i through the API with the same masks
    masks = {"tm": [synth.make_mask(70 + c, int(L), 0.3, lower=(c == 1)) for c, L in enumerate(sites.chrom_len)],
             "rm": [synth.make_mask(80 + c, int(L) // 2, 0.2) for c, L in enumerate(sites.chrom_len)]}
    for mname, per_chr in masks.items():
        for c, nm in enumerate(sites.chr_names):
            synth.write_mask(os.path.join(d, f"{mname}_chr{nm}.fa"), per_chr[c])
    out = os.path.join(d, "gpu_masked")
    cmd = [CLI, "--mode", "mut", "--mut", d + "/syn", "--chr", d + "/chr.txt", "--target_tmp", d + "/t.colate.in",
           "--reference_tmp", d + "/r.colate.in", "--target_mask", d + "/tm", "--reference_mask", d + "/rm", "--bins", "3,7,0.2",
           "--seed", "1", "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    up = lambda ms: [bytes(m).upper() for m in ms]
    o = po.stage1(sites, gt, gr, seed=1, tmask=up(masks["tm"]), rmask=up(masks["rm"]))
    counts = po.stage2(np.ones((1, o["num_blocks"]), np.int32), o, 0.0)
    ep, _ = po.epochs_from_bins("3,7,0.2")
    ro, it, _ = po.em_run(ep, np.full(len(ep), 1 / 20000.), counts[0])
    raw = open(out + ".bin", "rb").read()
    E = int(np.frombuffer(raw[12:16], np.int32)[0])
    assert np.array_equal(np.frombuffer(raw[16 + 8 * E:16 + 16 * E], np.float64), ro)
    assert f"Number of blocks: {o['num_blocks']}" in r.stderr


def _run(args):
    r = subprocess.run([CLI, "--mode", "mut"] + args, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return r


def test_cli_devices_sharding_is_byte_identical(dataset_dir):
    """--devices: chromosomes dealt to the devices for stage i (generator offsets and block bases exchanged in the process),
    replicates round-robin for stages ii-iii.  .coal and .bin do not depend on the device list -- two, three (more devices
    than chromosomes: one owns nothing) or, on a multi-GPU box, distinct GPUs."""
    import torch
    d, z, sites, gt, gr = dataset_dir
    base = ["--mut", d + "/syn", "--chr", d + "/chr.txt", "--target_tmp", d + "/t.colate.in", "--reference_tmp", d + "/r.colate.in",
            "--bins", "3,7,0.2", "--seed", "1", "--num_bootstraps", "23"]
    lists = ["0", "0,0", "0,0,0"] + (["0,1"] if torch.cuda.device_count() > 1 else [])
    outs = []
    for i, devs in enumerate(lists):
        out = os.path.join(d, f"gpu_dev{i}")
        r = _run(base + ["--devices", devs, "-o", out] + (["--host_parse"] if i == 1 else []))
        assert "Number of blocks: 13" in r.stderr
        outs.append((open(out + ".coal", "rb").read(), open(out + ".bin", "rb").read()))
    for o in outs[1:]:
        assert o == outs[0]
    # ... and equal the oracle: replicate 22 of 23
    raw = outs[0][1]
    R, E = np.frombuffer(raw[8:16], np.int32)
    rates = np.frombuffer(raw[16 + 8 * E:16 + 8 * E + 8 * R * E], np.float64).reshape(R, E)
    o = po.stage1(sites, gt, gr, seed=1)
    w = po.draw_block_weights(o["rng"], 23, o["num_blocks"])
    ep, _ = po.epochs_from_bins("3,7,0.2")
    ro, it, _ = po.em_run(ep, np.full(len(ep), 1 / 20000.), po.stage2(w, o, 0.0)[22])
    assert np.array_equal(rates[22], ro)


def test_cli_colate_mat_cache_and_coal_warm_start(dataset_dir):
    """SURVEY.md 8(f) N2.  (a) <out>.colate_mat present: parsing is skipped and the counts come from the 6-digit text
    (coal.cpp:3169-3170, 3471-3499); (b) --coal: epochs and initial rates from a .coal file (coal.cpp:3508-3549, 3638-3646).
    Both against .coal files the reference CLI wrote from the same inputs (tests/golden/make_golden.py n2), (a) also bitwise
    against the oracle's EM on the read-back counts."""
    import shutil
    d, z, sites, gt, gr = dataset_dir
    common = ["--mut", d + "/syn", "--chr", d + "/chr.txt", "--target_tmp", d + "/t.colate.in", "--reference_tmp", d + "/r.colate.in", "--seed", "1"]
    out = os.path.join(d, "gpu_cached")
    shutil.copy(os.path.join(GOLDEN, "n2_cache.colate_mat"), out + ".colate_mat")
    r = _run(common + ["--bins", "3,7,0.2", "--num_bootstraps", "3", "-o", out])
    assert "Loading precomputed file" in r.stderr and "Number of blocks" not in r.stderr
    assert open(out + ".coal").read() == open(os.path.join(GOLDEN, "n2_cache.coal")).read()
    vals = np.array(open(out + ".colate_mat").read().split(), dtype=np.float64)
    counts = vals[185:].reshape(3, 2, 185)
    raw = open(out + ".bin", "rb").read()
    R, E = np.frombuffer(raw[8:16], np.int32)
    rates = np.frombuffer(raw[16 + 8 * E:16 + 8 * E + 8 * R * E], np.float64).reshape(R, E)
    ep, _ = po.epochs_from_bins("3,7,0.2")
    for i in range(3):
        ro, it, _ = po.em_run(ep, np.full(len(ep), 1 / 20000.), counts[i], age_bin=vals[:185])   # the EM runs on the grid read back from the file
        assert np.array_equal(rates[i], ro)
    os.remove(out + ".colate_mat")
    out = os.path.join(d, "gpu_warm")
    r = _run(common + ["--coal", os.path.join(GOLDEN, "cli_ancient.coal"), "--target_age", "7000", "--reference_age", "0", "--years_per_gen", "28", "-o", out])
    assert open(out + ".coal").read() == open(os.path.join(GOLDEN, "n2_warm.coal")).read()
    # a non-ancient .coal starts "0 0 ...": the reference trips its own assert (coal.cpp:3548); this host reports it
    bad = subprocess.run([CLI, "--mode", "mut"] + common + ["--coal", os.path.join(GOLDEN, "cli_bins02_R1.coal"), "-o", out], capture_output=True, text=True)
    assert bad.returncode != 0 and "epochs must increase" in bad.stderr
