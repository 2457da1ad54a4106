"""GPU parity: device-side ingest of Relate .mut text vs the host reader
(colate_read_mut, which follows Mutations::Read, mutations.cpp:56-283)."""
import os
import tempfile

import numpy as np
import pytest

from colate_b200 import api, synth
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

HEADER = b"snp;pos_of_snp;dist;rs-id;tree_index;branch_indices;is_not_mapping;is_flipped;age_begin;age_end;ancestral_allele/alternative_allele;upstream_allele;downstream_allele;\n"


def _same_rows(got, want):
    for g, w, name in zip(got, want, ("pos", "age_begin", "age_end", "meta")):
        assert g.shape == w.shape, name
        # floats compared as bit patterns: -0.0 / NaN payloads included
        assert np.array_equal(g.view(np.uint32) if g.dtype == np.float32 else g, w.view(np.uint32) if w.dtype == np.float32 else w), name


def _host(text):
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "x.mut")
        open(p, "wb").write(text)
        return api.read_mut(p)


def test_ingest_matches_host_reader_on_generated_files(handle):
    sites = synth.make_sites(21, [30000, 0, 12000], [2.4e8, 1e8, 9e7], weird=0.15)
    with tempfile.TemporaryDirectory() as d:
        texts, want = [], []
        for c in range(3):
            p = os.path.join(d, f"s_chr{c}.mut")
            synth.write_mut(p, sites, c)
            texts.append(open(p, "rb").read())
            want.append(api.read_mut(p))
    rows = handle.ingest_mut(texts)
    assert rows == [len(w[0]) for w in want] == [30000, 0, 12000]
    got = handle.ingest_fetch()
    _same_rows(got, [np.concatenate([w[k] for w in want]) for k in range(4)])
    # the ingested arrays are the handle's sites: stage i on them == stage i on the parsed arrays
    gt = synth.make_genome(31, sites, 0.7)
    gr = synth.make_genome(32, sites, 0.7)
    al = lambda g: g.anc.astype(np.uint16) | (g.der.astype(np.uint16) << 8)
    for slot, g in ((0, gt), (1, gr)):
        handle.set_genome(slot, g.chrom, g.bp, g.aaf, g.daf, al(g))
        handle.set_mask(slot, None)
    s1 = handle.stage1(api.mt_seed(4))
    o = po.stage1(sites, gt, gr, seed=4)
    assert s1.num_blocks == o["num_blocks"] and s1.n_used == o["n_used_total"]
    assert np.array_equal(s1.block_stats[:, 0], o["shared"]) and np.array_equal(s1.block_stats[:, 1], o["notshared"])


def test_ingest_pinned_to_the_reference_reader(handle):
    """The device parser against Mutations::Read itself (fixture tests/golden/mut_reader.npz, written from the compiled
    reference by make_golden.py mut_reader): decimal -> float rounds exactly like std::stof on halfway cases, long mantissas,
    exponents, signs, leading / trailing zeros, inf / nan / hex; integers like std::stoi.  Lines the reference dies on
    (std::stoi / std::stof throw: no digits, out of range, overflow, underflow; missing fields) are refused."""
    from helpers import load
    z = load("mut_reader.npz")
    text = z["text"].tobytes()
    rows = handle.ingest_mut([text])
    assert rows == [len(z["pos"])]
    got = handle.ingest_fetch()
    _same_rows(got, (z["pos"], z["age_begin"], z["age_end"], z["meta"]))
    st = handle.ingest_stats()
    assert 0 < st["host_fallback_rows"] < len(z["pos"]) * 3 // 5   # the exotic spellings (10-digit positions, 17-digit mantissas, midpoints ...) went to the host, plain rows did not
    for line in z["bad"]:
        with pytest.raises(api._lib.ColateError) as e:
            handle.ingest_mut([HEADER + bytes(line)])
        assert e.value.code == -5, line


def test_ingest_errors(handle):
    for bad in (HEADER + b"1;2;3;.;5;6;0;0;1.0\n", HEADER + b"1;2;3;.;5;6;0;0;1;2;A/C;\n\n2;2;3;.;5;6;0;0;1;2;A/C;\n"):
        with pytest.raises(api._lib.ColateError) as e:
            handle.ingest_mut([bad])
        assert e.value.code == -5
    assert handle.ingest_mut([HEADER]) == [0] and handle.ingest_mut([b""]) == [0]


def _pin(a):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t, t.numpy()


def test_ingest_mut_bytes_pipelined_from_pinned_memory(handle):
    """colate_ingest_mut_texts: all chromosomes in one call, copies on the copy stream under the parse kernels."""
    sites = synth.make_sites(23, [25000, 1, 0, 9000, 14000], [2.4e8, 1e8, 9e7, 8e7, 7e7], weird=0.1)
    texts = [synth.mut_text_fast(sites, c) for c in range(5)]
    texts[2] = np.zeros(0, np.uint8)                              # an empty file
    keep = [_pin(t) if t.shape[0] else (None, t) for t in texts]
    rows = handle.ingest_mut_bytes([k[1] for k in keep])
    assert list(rows) == [25000, 1, 0, 9000, 14000]
    want = []
    with tempfile.TemporaryDirectory() as d:
        for c in (0, 1, 3, 4):
            p = os.path.join(d, f"s{c}.mut")
            synth.write_mut(p, sites, c)
            assert open(p, "rb").read() == texts[c].tobytes()     # the C writer == the Python writer
            want.append(api.read_mut(p))
    _same_rows(handle.ingest_fetch(), [np.concatenate([w[k] for w in want]) for k in range(4)])


@pytest.mark.parametrize("case", ["clean", "weird", "interleaved", "missing_chr", "truncated"])
def test_colate_in_ingest_on_the_device(handle, case):
    """colate_ingest_colate_in (runs by galloping on the host, verification + decode on the device) == the sequential
    reader (colate_read_colate_in + colate_chr_ranges + colate_set_genome) == the oracle, on clean files, files with
    records of chromosomes outside --chr, an island of another chromosome the run finder misses (host fallback), a chromosome without records (the
    reader runs to EOF and silences the later ones) and a truncated last record."""
    sites = synth.make_sites(5, [6000, 4000, 5000], [2.4e8, 1.2e8, 9e7], weird=0.05)
    gt = synth.make_genome(51, sites, 0.7, weird=0.1 if case == "weird" else 0.0)
    gr = synth.make_genome(52, sites, 0.6, weird=0.1 if case == "weird" else 0.0)
    names = list(sites.chr_names)
    with tempfile.TemporaryDirectory() as d:
        imgs = []
        for tag, g in (("t", gt), ("r", gr)):
            if case == "interleaved" and tag == "r":       # three records of chromosome 2 inside chromosome 1's run, where the
                k = int(np.searchsorted(g.chrom, 1))        # galloping probes (1, 2, 4, ... then bisection above 2048) never look:
                assert k > 2100                             # the device's header check must catch it -> sequential host decode
                order = np.concatenate([np.arange(0, 300), np.arange(k, k + 3), np.arange(300, k), np.arange(k + 3, g.n)])
                g = synth.Genome(*[x[order] for x in (g.chrom, g.bp, g.anc, g.der, g.aaf, g.daf)])
            if case == "missing_chr" and tag == "t":       # no record of the second chromosome
                sel = g.chrom != 1
                g = synth.Genome(*[x[sel] for x in (g.chrom, g.bp, g.anc, g.der, g.aaf, g.daf)])
            p = os.path.join(d, tag + ".colate.in")
            synth.write_colate_in(p, g, names, extra_names=["X", "chrUn"] if case == "weird" else None)
            img = open(p, "rb").read()
            if case == "truncated":
                img = img[:-7]
                open(p, "wb").write(img)
            imgs.append((p, img))
        handle.set_sites(sites.site_off, sites.pos, sites.age_begin, sites.age_end, sites.meta())
        want_n = []
        for slot, (p, img) in enumerate(imgs):
            rc, bp, aaf, daf, al = api.read_colate_in(p, names)
            handle.set_genome(slot, rc, bp, aaf, daf, al)
            want_n.append(len(bp))
        handle.set_mask(0, None); handle.set_mask(1, None)
        a = handle.stage1(api.mt_seed(9))
        handle.set_sites(sites.site_off, sites.pos, sites.age_begin, sites.age_end, sites.meta())
        for slot, (p, img) in enumerate(imgs):
            keep, pinned = _pin(np.frombuffer(img, dtype=np.uint8))
            assert handle.ingest_colate_in(slot, pinned if slot else img, names) == want_n[slot]
        b = handle.stage1(api.mt_seed(9))
    assert a.n_used == b.n_used and a.num_blocks == b.num_blocks and a.n_used > 0
    assert np.array_equal(a.block_stats, b.block_stats) and np.array_equal(a.block_tallies, b.block_tallies)
    assert np.array_equal(a.mt_state, b.mt_state)
    if case in ("clean", "weird"):
        o = po.stage1(sites, gt, gr, seed=9)
        assert b.n_used == o["n_used_total"] and np.array_equal(b.block_stats[:, 1], o["notshared"])


def test_unsorted_input_is_refused(handle):
    """COLATE_ERR_ORDER: the sequential reader of the reference needs ascending positions (coal.cpp:2184-2217); the
    device's binary-search join would silently differ on anything else."""
    sites = synth.make_sites(6, [3000, 2000], [2.4e8, 1.2e8])
    gt = synth.make_genome(61, sites, 0.7)
    gr = synth.make_genome(62, sites, 0.7)
    handle.load(sites, gt, gr)
    handle.stage1(api.mt_seed(1))                                   # sorted: fine (positions restart at a chromosome start)
    pos = sites.pos.copy()
    pos[[1500, 1501]] = pos[[1501, 1500]]
    handle.set_sites(sites.site_off, pos, sites.age_begin, sites.age_end, sites.meta())
    al = lambda g: g.anc.astype(np.uint16) | (g.der.astype(np.uint16) << 8)
    for slot, g in ((0, gt), (1, gr)):
        handle.set_genome(slot, g.chrom, g.bp, g.aaf, g.daf, al(g))
    with pytest.raises(api._lib.ColateError) as e:
        handle.stage1(api.mt_seed(1))
    assert e.value.code == -6 and ".mut" in str(e.value)
    handle.set_sites(sites.site_off, sites.pos, sites.age_begin, sites.age_end, sites.meta())
    with pytest.raises(api._lib.ColateError) as e:                  # new sites: every genome slot must be set again
        handle.stage1(api.mt_seed(1))
    assert e.value.code == -7
    bp = gr.bp.copy()
    bp[[700, 701]] = bp[[701, 700]]
    handle.set_genome(0, gt.chrom, gt.bp, gt.aaf, gt.daf, al(gt))
    handle.set_genome(1, gr.chrom, bp, gr.aaf, gr.daf, al(gr))
    with pytest.raises(api._lib.ColateError) as e:
        handle.stage1(api.mt_seed(1))
    assert e.value.code == -6 and ".colate.in" in str(e.value)
    handle.set_genome(1, gr.chrom, gr.bp, gr.aaf, gr.daf, al(gr))
    assert handle.stage1(api.mt_seed(1)).n_used > 0


def test_a_rank_without_chromosomes(handle):
    """n_chr == 0 is legal (world > n_chr in a chromosome-sharded job): zero used rows, zero blocks, and the generator
    state after the words of the ranks before it."""
    handle.set_sites(np.zeros(1, np.int64), np.zeros(0, np.int32), np.zeros(0, np.float32), np.zeros(0, np.float32), np.zeros(0, np.uint32))
    z = np.zeros(0, np.int32)
    for slot in (0, 1):
        handle.set_genome(slot, z, z, z, z, np.zeros(0, np.uint16))
    used, blocks = handle.stage1_flags()
    assert used.shape == (0,) and blocks.shape == (0,)
    stats, tallies, after = handle.stage1_sample(api.mt_seed(3), used_rank_base=777, block_base=5, n_blocks=0)
    want = api.mt_seed(3)
    api.lib().colate_mt_generate(want, 200 * 777, np.zeros(200 * 777, np.uint32))
    a, b = np.zeros(32, np.uint32), np.zeros(32, np.uint32)
    api.lib().colate_mt_generate(after, 32, a); api.lib().colate_mt_generate(want, 32, b)
    assert np.array_equal(a, b)
