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


def test_ingest_number_formats_are_strtof_exact(handle):
    """Decimal -> float must round exactly like strtof: halfway cases, long mantissas, exponents, signs,
    leading / trailing zeros, no digits, inf / nan / hex (host fallback), overflow and underflow."""
    rng = np.random.default_rng(7)
    ages = ["0", "-0", "0.0", "+5", ".5", "5.", "1e3", "1E-3", "1.5e+2", "12345.678", "0.000123", "16777217", "16777219",
            "33554434.000000001", "8388608.5", "8388609.5", "0.1", "0.30000001192092896", "3.4028235e38", "3.5e38", "1e39",
            "1e-39", "1.4e-45", "7e-46", "1e-50", "inf", "-inf", "nan", "0x1p3", "1e", "1e+", "e5", "", " 12", "12 ", "1.2.3",
            "123456789012345678", "1234567890123456789012", "0.00000000000000000000000000001", "1e400", "4.9406564584124654e-324",
            "1000000000000000000000000", "9007199254740993", "1.17549435e-38", "1.17549421e-38"]
    # exact float midpoints printed with all their digits, and their neighbours
    for _ in range(200):
        f = np.float32(np.exp(rng.uniform(-20, 20)))
        mid = (np.float64(f) + np.float64(np.nextafter(f, np.float32(np.inf)))) / 2
        ages += [format(mid, ".30g"), format(np.nextafter(mid, 0), ".30g"), format(np.nextafter(mid, np.inf), ".25g")]
    for _ in range(3000):
        x = np.exp(rng.uniform(-12, 18))
        ages.append(format(x, rng.choice([".3f", ".6g", ".9g", ".12g", ".17g", "e", ".1f"])))
    lines = [HEADER]
    for i, a in enumerate(ages):
        b = ages[(i * 7 + 3) % len(ages)]
        pos = ["17", " 42", "+9", "-3", "007", "2147483647", "99999999999"][i % 7]
        flip = ["0", "1", "00", "2", ""][i % 5]
        br = ["17", "17 23", " 5", "", "1 2 3"][i % 5]
        typ = ["A/C", "G/T", "AT/C", "A/", "/C", "0/1", "N/A", "A/C extra", "ACGTACGTACGTACGTACGT/A", "A/C"][i % 10]
        tail = ["A;C;\n", "\n", ";\n", "A;C;10 20 30\n"][i % 4]
        lines.append(f"{i};{pos};10;.;5;{br};0;{flip};{a};{b};{typ};".encode() + tail.encode())
    lines.append(b"9;5;1;.;1;7;0;0;1.5;2.5")              # last line: no type field, no newline
    text = b"".join(lines)
    want = _host(text)
    rows = handle.ingest_mut([text])
    assert rows == [len(want[0])]
    _same_rows(handle.ingest_fetch(), want)
    st = handle.ingest_stats()
    assert 0 < st["host_fallback_rows"] < len(ages)       # the exotic spellings went to the host, the bulk did not


def test_ingest_errors(handle):
    for bad in (HEADER + b"1;2;3;.;5;6;0;0;1.0\n", HEADER + b"1;2;3;.;5;6;0;0;1;2;A/C;\n\n2;2;3;.;5;6;0;0;1;2;A/C;\n"):
        with pytest.raises(api._lib.ColateError) as e:
            handle.ingest_mut([bad])
        assert e.value.code == -5
    assert handle.ingest_mut([HEADER]) == [0] and handle.ingest_mut([b""]) == [0]
