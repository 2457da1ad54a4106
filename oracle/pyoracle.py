"""TEST INFRASTRUCTURE ONLY: ctypes bindings for oracle/liboracle.so (the C restatement)
and oracle/_ref/libcolate_ref.so (the unmodified reference behind oracle/ref_probe.cpp).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NBINS, MAX_BLOCKS = 185, 500

_p = np.ctypeslib.ndpointer
_f64 = _p(dtype=np.float64, flags="C_CONTIGUOUS")
_i64 = _p(dtype=np.int64, flags="C_CONTIGUOUS")
_i32 = _p(dtype=np.int32, flags="C_CONTIGUOUS")
_u32 = _p(dtype=np.uint32, flags="C_CONTIGUOUS")
_f32 = _p(dtype=np.float32, flags="C_CONTIGUOUS")
_u8 = _p(dtype=np.uint8, flags="C_CONTIGUOUS")


def build(ref: bool = True):
    """make -C oracle (liboracle.so; plus _ref/ when /root/reference is present)."""
    subprocess.run(["make", "-s", "-C", HERE, "liboracle.so"] + (["ref"] if ref else []), check=True)


class MT(C.Structure):
    _fields_ = [("mt", C.c_uint32 * 624), ("pos", C.c_uint32)]

    def words(self):
        return np.frombuffer(self, dtype=np.uint32, count=625).copy()


class GenomeC(C.Structure):
    _fields_ = [("n", C.c_int64), ("chrom", C.c_void_p), ("bp", C.c_void_p), ("anc", C.c_void_p),
                ("der", C.c_void_p), ("aaf", C.c_void_p), ("daf", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = C.CDLL(path)
        L.oracle_mt_seed.argtypes = [C.POINTER(MT), C.c_uint32]
        L.oracle_mt_next.argtypes = [C.POINTER(MT)]
        L.oracle_mt_next.restype = C.c_uint32
        L.oracle_mt_words.argtypes = [C.c_uint32, C.c_long, C.c_int, _u32]
        L.oracle_uniform_real_n.argtypes = [C.c_uint32, C.c_long, C.c_int, _f64]
        L.oracle_uniform_int_n.argtypes = [C.c_uint32, C.c_long, C.c_int, C.c_int, _i32]
        L.oracle_age_bins.argtypes = [_f64]
        L.oracle_bin_of_double_age.argtypes = [C.c_double]
        L.oracle_bin_of_float_age.argtypes = [C.c_float]
        L.oracle_site_meta.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_char_p]
        L.oracle_site_meta.restype = C.c_uint32
        L.oracle_stage1.argtypes = [C.c_int, _i64, _i32, _f32, _f32, _u32,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.POINTER(GenomeC), C.POINTER(GenomeC), C.POINTER(MT),
                                    _f64, _f64, _f64, _f64, _i64, _i64, _i64, _i64, C.POINTER(C.c_int64)]
        L.oracle_stage1_pileup.argtypes = [C.c_int, _i64, _i32, _f32, _f32, _u32,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _i32, _i32, C.POINTER(MT),
                                           _f64, _f64, _f64, _f64, _i64, _i64, _i64, _i64, C.POINTER(C.c_int64)]
        L.oracle_stage1_counts.argtypes = [C.c_int] + list(L.oracle_stage1_pileup.argtypes)
        L.oracle_stage2_norm.argtypes = [C.c_int, _f64]
        L.oracle_draw_block_weights.argtypes = [C.POINTER(MT), C.c_int, C.c_int, _i32]
        L.oracle_stage2.argtypes = [C.c_int, C.c_int, _i32, _f64, _f64, _f64, _f64, C.c_double, _f64, _f64]
        L.oracle_age_generations.argtypes = [C.c_char_p, C.c_char_p, C.c_float, C.c_int, C.POINTER(C.c_double)]
        L.oracle_age_generations.restype = C.c_double
        L.oracle_epochs_from_bins.argtypes = [C.c_char_p, C.c_double, C.c_double, _f64, C.c_int, C.POINTER(C.c_int)]
        L.oracle_epochs_from_coal_line.argtypes = [C.c_char_p, C.c_double, _f64, C.c_int]
        L.oracle_estep.argtypes = [C.c_int, C.c_int, _f64, _f64, C.c_double, _f64, _f64]
        L.oracle_estep.restype = C.c_double
        L.oracle_em_run.argtypes = [C.c_int, _f64, _f64, _f64, _f64, C.c_int, _f64, C.POINTER(C.c_double)]
        L.oracle_libm.argtypes = [C.c_int, C.c_int, _f64, _f64]
        L.oracle_write_coal.argtypes = [C.c_char_p, C.c_int, C.c_int, _f64, _f64, C.c_int, C.c_int]
        _lib = L
    return _lib


def _genome_c(g):
    arrs = [np.ascontiguousarray(x) for x in (g.chrom, g.bp, g.anc, g.der, g.aaf, g.daf)]
    gc = GenomeC(g.n, *[a.ctypes.data for a in arrs])
    gc._keep = arrs
    return gc


def age_bins():
    out = np.zeros(NBINS)
    lib().oracle_age_bins(out)
    return out


def mt_seed(seed: int) -> MT:
    g = MT()
    lib().oracle_mt_seed(C.byref(g), seed & 0xFFFFFFFF)
    return g


def stage1(sites, target, reference, seed=1, tmask=None, rmask=None, rng: MT | None = None):
    """Oracle stage i on parsed arrays.  tmask/rmask: list of bytes (upper-cased mask sequence
    per chromosome) or None.  Returns dict with num_blocks, the four [nb,185] fp64 block
    vectors, the integer tallies, n_used and the generator state after the stage."""
    L = lib()
    rng = rng or mt_seed(seed)
    n_chr = len(sites.chr_names)
    z = lambda dt=np.float64: np.zeros((MAX_BLOCKS, NBINS), dtype=dt)
    S, N, SE, NE = z(), z(), z(), z()
    nS, nN, nE = z(np.int64), z(np.int64), z(np.int64)
    nU = np.zeros(MAX_BLOCKS, dtype=np.int64)
    tot = C.c_int64(0)

    def mk(mask):
        if mask is None:
            return None, None, None
        bufs = [C.create_string_buffer(m, len(m)) for m in mask]
        ptrs = (C.c_char_p * n_chr)(*[C.cast(b, C.c_char_p) for b in bufs])
        lens = (C.c_int64 * n_chr)(*[len(m) for m in mask])
        return bufs, ptrs, lens

    tb, tp, tl = mk(tmask)
    rb, rp, rl = mk(rmask)
    gt, gr = _genome_c(target), _genome_c(reference)
    nb = L.oracle_stage1(n_chr, np.ascontiguousarray(sites.site_off, dtype=np.int64),
                         np.ascontiguousarray(sites.pos), np.ascontiguousarray(sites.age_begin),
                         np.ascontiguousarray(sites.age_end), np.ascontiguousarray(sites.meta()),
                         C.cast(tp, C.c_void_p) if tp else None, C.cast(tl, C.c_void_p) if tl else None,
                         C.cast(rp, C.c_void_p) if rp else None, C.cast(rl, C.c_void_p) if rl else None,
                         C.byref(gt), C.byref(gr), C.byref(rng), S, N, SE, NE, nS, nN, nE, nU, C.byref(tot))
    if nb < 0:
        return {"num_blocks": nb}
    return {"num_blocks": nb, "shared": S[:nb].copy(), "notshared": N[:nb].copy(), "shared_emp": SE[:nb].copy(),
            "notshared_emp": NE[:nb].copy(), "n_shared": nS[:nb].copy(), "n_notshared": nN[:nb].copy(),
            "n_emp": nE[:nb].copy(), "n_used": nU[:nb].copy(), "n_used_total": tot.value, "rng": rng}


def stage1_pileup(sites, t_counts, r_counts, seed=1, tmask=None, rmask=None, rng: MT | None = None):
    """Oracle stage i of the bam/bam front-end (parse_onebambam, coal.cpp:1799-2069) on pre-decoded pileups:
    t_counts / r_counts [n_site][4] = reads showing A, C, G, T at each row.  Returns as stage1()."""
    L = lib()
    rng = rng or mt_seed(seed)
    n_chr = len(sites.chr_names)
    z = lambda dt=np.float64: np.zeros((MAX_BLOCKS, NBINS), dtype=dt)
    S, N, SE, NE = z(), z(), z(), z()
    nS, nN, nE = z(np.int64), z(np.int64), z(np.int64)
    nU = np.zeros(MAX_BLOCKS, dtype=np.int64)
    tot = C.c_int64(0)

    def mk(mask):
        if mask is None:
            return None, None, None
        bufs = [C.create_string_buffer(m, len(m)) for m in mask]
        ptrs = (C.c_char_p * n_chr)(*[C.cast(b, C.c_char_p) for b in bufs])
        lens = (C.c_int64 * n_chr)(*[len(m) for m in mask])
        return bufs, ptrs, lens

    tb, tp, tl = mk(tmask)
    rb, rp, rl = mk(rmask)
    width = np.asarray(t_counts).shape[1]         # 4: A, C, G, T pileups; 2: (AAF, DAF) per row from a bcf decoder
    nb = L.oracle_stage1_counts(width, n_chr, np.ascontiguousarray(sites.site_off, dtype=np.int64), np.ascontiguousarray(sites.pos),
                                np.ascontiguousarray(sites.age_begin), np.ascontiguousarray(sites.age_end), np.ascontiguousarray(sites.meta()),
                                C.cast(tp, C.c_void_p) if tp else None, C.cast(tl, C.c_void_p) if tl else None,
                                C.cast(rp, C.c_void_p) if rp else None, C.cast(rl, C.c_void_p) if rl else None,
                                np.ascontiguousarray(t_counts, dtype=np.int32), np.ascontiguousarray(r_counts, dtype=np.int32),
                                C.byref(rng), S, N, SE, NE, nS, nN, nE, nU, C.byref(tot))
    if nb < 0:
        return {"num_blocks": nb}
    return {"num_blocks": nb, "shared": S[:nb].copy(), "notshared": N[:nb].copy(), "shared_emp": SE[:nb].copy(),
            "notshared_emp": NE[:nb].copy(), "n_shared": nS[:nb].copy(), "n_notshared": nN[:nb].copy(),
            "n_emp": nE[:nb].copy(), "n_used": nU[:nb].copy(), "n_used_total": tot.value, "rng": rng}


def draw_block_weights(rng: MT, R: int, num_blocks: int):
    w = np.zeros((R, num_blocks), dtype=np.int32)
    lib().oracle_draw_block_weights(C.byref(rng), R, num_blocks, w)
    return w


def stage2(weights, blk, age: float = 0.0, norm_1e3: bool = False):
    """norm_1e3: every front-end but tmp/tmp divides both vectors by 1e3 afterwards (coal.cpp:3453-3463)."""
    R, nb = weights.shape
    counts = np.zeros((R, 2, NBINS))
    lib().oracle_stage2(R, nb, np.ascontiguousarray(weights), np.ascontiguousarray(blk["shared"]),
                        np.ascontiguousarray(blk["notshared"]), np.ascontiguousarray(blk["shared_emp"]),
                        np.ascontiguousarray(blk["notshared_emp"]), age, age_bins(), counts)
    if norm_1e3:
        lib().oracle_stage2_norm(R, counts)
    return counts


def ages(target_age=None, reference_age=None, years_per_gen=None):
    ypg = C.c_double(0)
    a = lib().oracle_age_generations(target_age.encode() if target_age else None,
                                     reference_age.encode() if reference_age else None,
                                     years_per_gen if years_per_gen is not None else 0.0,
                                     1 if years_per_gen is not None else 0, C.byref(ypg))
    return a, ypg.value


def epochs_from_bins(bins: str, age: float = 0.0, years_per_gen: float = 28.0):
    ep = np.zeros(1024)
    null = C.c_int(0)
    n = lib().oracle_epochs_from_bins(bins.encode(), age, years_per_gen, ep, 1024, C.byref(null))
    if n < 0:
        raise ValueError("bad --bins")
    return ep[:n].copy(), null.value


def epochs_from_coal_line(line: str, age: float = 0.0):
    ep = np.zeros(1024)
    n = lib().oracle_epochs_from_coal_line(line.encode(), age, ep, 1024)
    if n < 0:
        raise ValueError("bad --coal epoch line")
    return ep[:n].copy()


def estep(shared: bool, epochs, rates, t: float):
    E = len(epochs)
    num, den = np.zeros(E), np.zeros(E)
    ll = lib().oracle_estep(1 if shared else 0, E, np.ascontiguousarray(epochs, dtype=np.float64),
                            np.ascontiguousarray(rates, dtype=np.float64), t, num, den)
    return ll, num, den


def em_run(epochs, rates_init, counts, max_iter=100000, age_bin=None):
    """age_bin: the point ages of the 185 bins (default: the exact grid; mut() started from a .colate_mat cache uses the
    grid read back from the file, coal.cpp:3481-3483)."""
    E = len(epochs)
    out = np.zeros(E)
    ll = C.c_double(0)
    it = lib().oracle_em_run(E, np.ascontiguousarray(epochs, dtype=np.float64),
                             np.ascontiguousarray(rates_init, dtype=np.float64),
                             age_bins() if age_bin is None else np.ascontiguousarray(age_bin, dtype=np.float64),
                             np.ascontiguousarray(counts, dtype=np.float64), max_iter, out, C.byref(ll))
    return out, it, ll.value


def libm(which: str, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.zeros_like(x)
    lib().oracle_libm({"exp": 0, "log": 1, "log1p": 2}[which], x.shape[0], x, y)
    return y


def write_coal(path, epochs, rates, is_ancient=False, ep_null=0):
    rates = np.ascontiguousarray(rates, dtype=np.float64).copy()
    R, E = rates.shape
    rc = lib().oracle_write_coal(path.encode(), R, E, np.ascontiguousarray(epochs, dtype=np.float64), rates,
                                 1 if is_ancient else 0, ep_null)
    if rc:
        raise OSError(path)
    return rates


# ---------------------------------------------------------------- compiled reference
_ref = None
REF_LIB = os.path.join(HERE, "_ref", "libcolate_ref.so")


def ref_available() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libcolate_ref.so"))


def ref_cli() -> str | None:
    p = os.path.join(HERE, "_ref", "Colate")
    return p if os.path.exists(p) else None


def ref():
    global _ref
    if _ref is None:
        L = C.CDLL(os.path.join(HERE, "_ref", "libcolate_ref.so"))
        L.ref_parse_tmptmp.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_char_p, C.c_char_p,
                                       C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_double, C.c_double, C.c_int,
                                       _f64, _f64, _f64, _f64, _f64, _u32, _u32]
        for nm in ("ref_em_shared", "ref_em_notshared"):
            f = getattr(L, nm)
            f.argtypes = [C.c_int, _f64, _f64, C.c_double, C.c_double, _f64, _f64]
            f.restype = C.c_double
        for nm in ("ref_em_simplified_shared", "ref_em_simplified_notshared"):
            f = getattr(L, nm)
            f.argtypes = [C.c_int, _f64, _f64, C.c_double, _f64, _f64]
            f.restype = C.c_double
        L.ref_uniform_real.argtypes = [C.c_int, C.c_long, C.c_int, _f64]
        L.ref_uniform_int.argtypes = [C.c_int, C.c_long, C.c_int, C.c_int, _i32]
        L.ref_mt_words.argtypes = [C.c_int, C.c_long, C.c_int, _u32]
        L.ref_bin_of_float_age.argtypes = [C.c_float]
        L.ref_bin_of_double_age.argtypes = [C.c_double]
        L.ref_parse_onebambam.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_char_p, C.c_char_p,
                                          C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_int,
                                          _f64, _f64, _f64, _f64, _f64, _u32]
        L.ref_bam_pileup.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_long, _i32, _i32, _u8]
        _ref = L
    return _ref


_REF_READ_MUT = r"""
import ctypes as C, sys, numpy as np
L = C.CDLL(sys.argv[1])
cap = int(sys.argv[3])
pos = np.zeros(cap, np.int32); ab = np.zeros(cap, np.float32); ae = np.zeros(cap, np.float32)
fl = np.zeros(cap, np.int32); nb = np.zeros(cap, np.int32); ty = np.zeros((cap, 16), np.uint8); tl = np.zeros(cap, np.int32)
L.ref_read_mut.restype = C.c_long
p = lambda a: a.ctypes.data_as(C.c_void_p)
n = L.ref_read_mut(sys.argv[2].encode(), C.c_long(cap), p(pos), p(ab), p(ae), p(fl), p(nb), p(ty), p(tl))
if n < 0:
    sys.exit(3)
np.savez(sys.argv[4], n=n, pos=pos[:n], age_begin=ab[:n], age_end=ae[:n], flipped=fl[:n], n_branch=nb[:n], type15=ty[:n], type_len=tl[:n])
"""


def ref_read_mut(path, cap=1 << 20):
    """The reference's own Mutations::Read (mutations.cpp:56-283) on one .mut file, in a child process (it calls exit(1) on a
    field std::stoi cannot convert and dies of an uncaught exception on one std::stof cannot).  Returns the fields the
    tmp/tmp path uses, or None when the reference did not survive the file."""
    import subprocess
    import sys
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "rows.npz")
        r = subprocess.run([sys.executable, "-c", _REF_READ_MUT, REF_LIB, path, str(cap), out], capture_output=True)
        if r.returncode != 0 or not os.path.exists(out):
            return None
        z = np.load(out)
        return {k: z[k] for k in z.files}


def ref_meta(rows):
    """Packed site word (colate_site_meta) from the reference reader's rows: coal.cpp:2150-2176 minus the masks."""
    n = int(rows["n"])
    meta = np.zeros(n, np.uint32)
    for i in range(n):
        t = bytes(rows["type15"][i]).split(b"\0")[0]
        ok = rows["flipped"][i] == 0 and rows["n_branch"][i] == 1 and rows["age_begin"][i] < rows["age_end"][i] and rows["age_end"][i] >= 0
        ok = ok and rows["type_len"][i] == 3 and len(t) == 3 and t[1:2] == b"/" and t[0:1] in b"ACGT0" and t[2:3] in b"ACGT1"
        if ok:
            meta[i] = 1 | (t[0] << 8) | (t[2] << 16)
    return meta


def ref_parse_tmptmp(dirname, chr_names, prefix, target, reference, seed=1, tmask=None, rmask=None):
    """Runs the reference's parse_tmptmp on files written by colate_b200.synth.write_dataset."""
    L = ref()
    n = len(chr_names)
    arr = lambda xs: (C.c_char_p * n)(*[x.encode() for x in xs])
    names = arr(chr_names)
    muts = arr([os.path.join(dirname, f"{prefix}_chr{c}.mut") for c in chr_names])
    tm = arr([os.path.join(dirname, f"{tmask}_chr{c}.fa") for c in chr_names]) if tmask else None
    rm = arr([os.path.join(dirname, f"{rmask}_chr{c}.fa") for c in chr_names]) if rmask else None
    z = lambda: np.zeros((MAX_BLOCKS, NBINS))
    S, N, SE, NE = z(), z(), z(), z()
    rest = np.zeros(2)
    mt = np.zeros(625, dtype=np.uint32)
    nxt = np.zeros(8, dtype=np.uint32)
    nb = L.ref_parse_tmptmp(n, names, muts, os.path.join(dirname, target + ".colate.in").encode(),
                            os.path.join(dirname, reference + ".colate.in").encode(), tm, rm, 0.0, 0.0, seed,
                            S, N, SE, NE, rest, mt, nxt)
    return {"num_blocks": nb, "shared": S[:nb].copy(), "notshared": N[:nb].copy(), "shared_emp": SE[:nb].copy(),
            "notshared_emp": NE[:nb].copy(), "emp_rest": rest, "mt": mt, "next_words": nxt}


def ref_estep(shared: bool, epochs, rates, t: float):
    E = len(epochs)
    num, den = np.zeros(E), np.zeros(E)
    f = ref().ref_em_shared if shared else ref().ref_em_notshared
    ll = f(E, np.ascontiguousarray(epochs, dtype=np.float64), np.ascontiguousarray(rates, dtype=np.float64), t, t, num, den)
    return ll, num, den


def ref_estep_simplified(shared: bool, epochs, rates, t: float):
    E = len(epochs)
    num, den = np.zeros(E), np.zeros(E)
    f = ref().ref_em_simplified_shared if shared else ref().ref_em_simplified_notshared
    ll = f(E, np.ascontiguousarray(epochs, dtype=np.float64), np.ascontiguousarray(rates, dtype=np.float64), t, num, den)
    return ll, num, den


# ---- the bam/bam front-end on synthetic reads (oracle/hts_stubs.c serves the reference's bam_parser) ----
def write_fake_bam(path, chr_names, reads):
    """reads: iterable of (tid, pos0, mapq, reverse, seq bytes, qual uint8 array), sorted by (tid, pos0).
    Format: oracle/hts_stubs.c."""
    import struct
    with open(path, "wb") as f:
        f.write(b"FBAM" + struct.pack("<i", len(chr_names)))
        for nm in chr_names:
            b = nm.encode()
            f.write(struct.pack("<i", len(b)) + b)
        for tid, pos0, mapq, rev, seq, qual in reads:
            f.write(struct.pack("<iiBBi", tid, pos0, mapq, 1 if rev else 0, len(seq)) + bytes(seq) + bytes(bytearray(qual)))


def ref_parse_onebambam(dirname, chr_names, prefix, target_bam, ref_bam, ref_genome, seed=1, params="20,30,10", tmask=None, rmask=None):
    """The reference's parse_onebambam (coal.cpp:1799-2069) on <dirname>/<prefix>_chr<c>.mut, two fake BAMs and the
    per-chromosome reference genome <ref_genome>_chr<c>.fa."""
    L = ref()
    n = len(chr_names)
    arr = lambda xs: (C.c_char_p * n)(*[x.encode() for x in xs])
    names = arr(chr_names)
    muts = arr([os.path.join(dirname, f"{prefix}_chr{c}.mut") for c in chr_names])
    gen = arr([os.path.join(dirname, f"{ref_genome}_chr{c}.fa") for c in chr_names])
    tm = arr([os.path.join(dirname, f"{tmask}_chr{c}.fa") for c in chr_names]) if tmask else None
    rm = arr([os.path.join(dirname, f"{rmask}_chr{c}.fa") for c in chr_names]) if rmask else None
    z = lambda: np.zeros((MAX_BLOCKS, NBINS))
    S, N, SE, NE = z(), z(), z(), z()
    rest = np.zeros(2)
    mt = np.zeros(625, dtype=np.uint32)
    nb = L.ref_parse_onebambam(params.encode(), n, names, muts, os.path.join(dirname, target_bam).encode(),
                               os.path.join(dirname, ref_bam).encode(), tm, rm, gen, seed, S, N, SE, NE, rest, mt)
    return {"num_blocks": nb, "shared": S[:nb].copy(), "notshared": N[:nb].copy(), "shared_emp": SE[:nb].copy(),
            "notshared_emp": NE[:nb].copy(), "emp_rest": rest, "mt": mt}


def ref_bam_pileup(bam_path, contig, ref_genome_path, bp, params="20,30,10"):
    """What the reference's bam_parser holds at the 1-based positions bp (ascending) of one contig: counts [n][4] (A, C, G, T)."""
    bp = np.ascontiguousarray(bp, dtype=np.int32)
    counts = np.zeros((bp.shape[0], 4), dtype=np.int32)
    cov = np.zeros(bp.shape[0], dtype=np.uint8)
    ref().ref_bam_pileup(params.encode(), bam_path.encode(), contig.encode(), ref_genome_path.encode(), bp.shape[0], bp, counts, cov)
    return counts, cov


def pileup_from_reads(pos_rows, reads, ref_seq, filters=(20, 30, 10)):
    """Restatement of what bam_parser (include/vcf/htslib.cpp:60-168, 379-437, 490-573) holds at the 1-based positions `pos_rows`
    of ONE contig after reading its `reads` [(pos0, mapq, seq bytes, qual uint8 array)] in file order: counts [n][4] (A, C, G, T).
    Pure Python (small cases).  The first read of the contig is counted by assign_contig, which leaves `q` pointing at the record's
    packed 4-bit sequence instead of its qualities (htslib.cpp:549 against 406): "quality i" of that read is byte i of
    {packed sequence, qualities}."""
    code = {c: i for i, c in enumerate(b"=ACMGRSVTWYHKDBN")}
    rows = {int(p) - 1: i for i, p in enumerate(pos_rows)}
    out = np.zeros((len(pos_rows), 4), dtype=np.int32)
    for k, (p, mq, seq, q) in enumerate(reads):
        l = len(seq)
        if k == 0:
            nb = (l + 1) // 2
            packed = [(code.get(seq[2 * i], 15) << 4) | (code.get(seq[2 * i + 1], 15) if 2 * i + 1 < l else 0) for i in range(nb)]
            q = np.array(packed + [int(x) for x in q[:l - nb]], dtype=np.int64)
        if mq < filters[0] or l < filters[1]:
            continue
        tot = mat = 0
        for i in range(3, l - 3):
            if p + i >= len(ref_seq):
                break
            if q[i] >= 30:
                tot += 1
                mat += int(ref_seq[p + i] == seq[i])
        if not (tot > 0 and tot - mat <= filters[2]):
            continue
        for i in range(3, l - 3):
            if q[i] >= 30 and (p + i) in rows:
                col = b"ACGT".find(bytes([seq[i]]))
                if col >= 0:
                    out[rows[p + i], col] += 1
    return out


# ---- the bcf/bcf front-end on synthetic genotype records (oracle/hts_stubs.c serves the reference's vcf_parser) ----
def write_fake_bcf(path, n_samples, ploidy, records):
    """records: iterable of (pos0, [allele bytes, ...], [allele index per haplotype] of length n_samples * ploidy), ascending pos0."""
    import struct
    with open(path, "wb") as f:
        f.write(b"FBCF" + struct.pack("<ii", n_samples, ploidy))
        for pos0, alleles, gt in records:
            f.write(struct.pack("<ii", pos0, len(alleles)))
            for a in alleles:
                f.write(struct.pack("<i", len(a)) + bytes(a))
            f.write(struct.pack("<%di" % (n_samples * ploidy), *[int(x) for x in gt]))


def ref_parse_vcfvcf(dirname, chr_names, prefix, target_bcf, ref_bcf, seed=1, ref_genome=None, tmask=None, rmask=None):
    """The reference's parse_vcfvcf (coal.cpp:907-1228) on <dirname>/<prefix>_chr<c>.mut and the fake BCFs <target_bcf>_chr<c>.bcf /
    <ref_bcf>_chr<c>.bcf (+ <ref_genome>_chr<c>.fa, masks)."""
    L = ref()
    L.ref_parse_vcfvcf.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.POINTER(C.c_char_p),
                                   C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_int, _f64, _f64, _f64, _f64, _f64, _u32]
    n = len(chr_names)
    arr = lambda xs: (C.c_char_p * n)(*[x.encode() for x in xs])
    per = lambda stem, ext: arr([os.path.join(dirname, f"{stem}_chr{c}.{ext}") for c in chr_names])
    z = lambda: np.zeros((MAX_BLOCKS, NBINS))
    S, N, SE, NE = z(), z(), z(), z()
    rest = np.zeros(2)
    mt = np.zeros(625, dtype=np.uint32)
    nb = L.ref_parse_vcfvcf(n, per(prefix, "mut"), per(target_bcf, "bcf"), per(ref_bcf, "bcf"), per(tmask, "fa") if tmask else None,
                            per(rmask, "fa") if rmask else None, per(ref_genome, "fa") if ref_genome else None, seed, S, N, SE, NE, rest, mt)
    return {"num_blocks": nb, "shared": S[:nb].copy(), "notshared": N[:nb].copy(), "shared_emp": SE[:nb].copy(),
            "notshared_emp": NE[:nb].copy(), "emp_rest": rest, "mt": mt}


def decode_vcfvcf(pos_rows, anc_rows, der_rows, usable_rows, recs_target, recs_ref, n_target, n_ref, refg_at_rows=None):
    """The DECODER half of parse_vcfvcf (coal.cpp:997-1137) for ONE chromosome, restated: what the two genotype streams say about
    every .mut row, as the per-row counts colate_set_row_counts / stage1_pileup (width 2) take: (AAF, DAF) of the row's two
    alleles per genome, AAF = N - DAF; zeros = the row is not usable for this genome.
      pos_rows / anc_rows / der_rows: the rows (1-based positions, allele codes); usable_rows: the row passes the filter in front
      of the lookups (coal.cpp:966-995: meta bit 0 and ancestral != derived) -- the cursors only move for such rows;
      recs_*: [(pos0, [alleles], [gt])] ascending; n_*: haplotypes per record (samples x ploidy); refg_at_rows: base of the
      reference genome at every row (--ref_genome given) or None.
    Masks are not the decoder's business (they gate `use` before the lookups; the cursors lose nothing by not moving)."""
    n = len(pos_rows)
    tc = np.zeros((n, 2), np.int32)
    rc = np.zeros((n, 2), np.int32)

    class Cur:                       # vcf_parser + the bp_* variable of parse_vcfvcf: read_snp() keeps the last record at end of file
        def __init__(self, recs):
            self.recs, self.k, self.rec = recs, 0, None
            self.read()
            self.bp = self.rec[0] + 1 if self.rec else 0       # coal.cpp:953-954 / 958-959

        def read(self):
            if self.k < len(self.recs):
                self.rec = self.recs[self.k]; self.k += 1
                return 0
            return -1

        def seek(self, bp_mut):      # coal.cpp:1001-1006
            if self.bp < bp_mut:
                while self.read() == 0:
                    self.bp = self.rec[0] + 1
                    if self.bp >= bp_mut:
                        break

    cr, ct = Cur(recs_ref), Cur(recs_target)
    for m in range(n):
        if not usable_rows[m]:
            continue
        bp, anc, der = int(pos_rows[m]), bytes([anc_rows[m]]), bytes([der_rows[m]])
        use = True
        cr.seek(bp)
        daf_r = 0
        if cr.bp == bp and cr.rec is not None:
            al = cr.rec[1]
            a0, a1 = bytes(al[0]), bytes(al[1]) if len(al) > 1 else None      # (the synthetic reference records always carry two alleles)
            if (a0 == anc and a1 == der) or (a0 == der and a1 == anc):
                flip = a0 == der and a1 == anc
                gt = cr.rec[2]
                daf_r = int(sum(gt))
                if all(g <= 1 for g in gt):
                    if flip:
                        daf_r = n_ref - daf_r
                else:
                    use = False
            else:
                use = False
        else:
            if refg_at_rows is not None:
                if der == bytes([refg_at_rows[m]]):
                    daf_r = n_ref
                else:
                    use = False
            else:
                use = False
        if daf_r == 0:
            use = False
        if not use:
            continue
        ct.seek(bp)
        daf_t = 0
        if ct.bp == bp and ct.rec is not None:
            al = ct.rec[1]
            a0 = bytes(al[0]) if len(al) > 0 else b""
            a1 = bytes(al[1]) if len(al) > 1 else b""
            accept = fixed = tflip = False
            if a0 != b"" and a1 == b"":
                if a0 == anc or a0 == der:
                    accept = True
                if a0 == der:
                    tflip = True
                fixed = True
            elif (a0 == der and a1 == anc) or (a0 == anc and a1 == der):
                accept = True
                tflip = a0 == der and a1 == anc
            if accept:
                gt = ct.rec[2]
                daf_t = int(sum(gt))
                if not all(g <= 1 for g in gt):
                    accept = False
                if fixed and daf_t != 0:
                    accept = False
            if not accept:
                use = False
            if tflip:
                daf_t = n_target - daf_t
        else:
            if refg_at_rows is not None:
                b = bytes([refg_at_rows[m]])
                if der == b:
                    daf_t = n_target
                elif anc == b:
                    daf_t = 0
                else:
                    use = False
            else:
                use = False
        if not use:
            continue
        rc[m] = (n_ref - daf_r, daf_r)
        tc[m] = (n_target - daf_t, daf_t)
    return tc, rc
