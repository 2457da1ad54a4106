"""Seeded synthetic inputs of the shapes BASELINE.json names.

Produces the parsed structure-of-arrays the C-ABI consumes and, for file-level tests and
the reference arm, the same data as the reference's on-disk formats:
Relate ``.mut`` text, ``.colate.in`` binary records, P/N fasta masks and a ``--chr`` list.

Generator: numpy ``default_rng(seed)`` (PCG64).  Ages are generated as float32 and written
with their shortest round-trip repr, so ``strtof`` of the file gives the array bit for bit.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

# GRCh37-like autosome lengths (bp), used for the whole-genome shapes
AUTOSOME_LEN = [249250621, 243199373, 198022430, 191154276, 180915260, 171115067, 159138663,
                146364022, 141213431, 135534747, 135006516, 133851895, 115169878, 107349540,
                102531392, 90354753, 81195210, 78077248, 59128983, 63025520, 48129895, 51304566]

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


@dataclass
class Sites:
    """Rows of the per-chromosome .mut files, concatenated in --chr order."""
    chr_names: list
    site_off: np.ndarray      # int64 [n_chr+1]
    pos: np.ndarray           # int32
    age_begin: np.ndarray     # float32
    age_end: np.ndarray       # float32
    flipped: np.ndarray       # uint8
    n_branch: np.ndarray      # int32
    anc: np.ndarray           # uint8 first char of the ancestral allele string
    der: np.ndarray           # uint8 first char of the derived allele string
    odd: np.ndarray           # uint8: 0 normal "X/Y"; 1 multi-base ancestral; 2 no '/' ("NA")
    chrom_len: list = field(default_factory=list)

    @property
    def n(self):
        return int(self.pos.shape[0])

    def meta(self) -> np.ndarray:
        """Packed site word (bit0 = row passes coal.cpp:2150/2166/2175-2176, byte1 anc, byte2 der)."""
        ok = (self.flipped == 0) & (self.n_branch == 1) & (self.age_begin < self.age_end) & (self.age_end >= 0)
        ok &= self.odd == 0
        ok &= np.isin(self.anc, np.frombuffer(b"ACGT0", dtype=np.uint8))
        ok &= np.isin(self.der, np.frombuffer(b"ACGT1", dtype=np.uint8))
        m = ok.astype(np.uint32) | (self.anc.astype(np.uint32) << 8) | (self.der.astype(np.uint32) << 16)
        return np.where(ok, m, 0).astype(np.uint32)


@dataclass
class Genome:
    """Records of one .colate.in file in file order."""
    chrom: np.ndarray   # int32 index into chr_names (>= n_chr: a name not in the list)
    bp: np.ndarray      # int32
    anc: np.ndarray     # uint8
    der: np.ndarray     # uint8
    aaf: np.ndarray     # int32
    daf: np.ndarray     # int32

    @property
    def n(self):
        return int(self.bp.shape[0])


def _loguniform(rng, lo, hi, n):
    return np.exp(rng.uniform(np.log(lo), np.log(hi), n))


def make_sites(seed: int, rows_per_chr, chrom_len, chr_names=None, weird: float = 0.0) -> Sites:
    """SURVEY.md 8d config-1 distributions.  ``weird`` adds a fraction of rows that exercise
    the filter edge cases (multi-base alleles, missing '/', age_begin >= age_end, negative
    age_begin, 0/1 allele codes)."""
    rng = np.random.default_rng(seed)
    n_chr = len(rows_per_chr)
    chr_names = chr_names or [str(i + 1) for i in range(n_chr)]
    off = np.zeros(n_chr + 1, dtype=np.int64)
    off[1:] = np.cumsum(rows_per_chr)
    n = int(off[-1])
    pos = np.empty(n, dtype=np.int32)
    for c in range(n_chr):
        k = int(rows_per_chr[c])
        L = int(chrom_len[c])
        if k > L - 1:
            raise ValueError("more rows than positions")
        if k * 4 < L:
            p = np.unique(rng.integers(1, L, size=int(k * 1.05) + 16))
            while p.shape[0] < k:
                p = np.unique(np.concatenate([p, rng.integers(1, L, size=k)]))
            p = np.sort(rng.choice(p, size=k, replace=False)) if p.shape[0] > k else p
        else:
            p = np.sort(rng.choice(np.arange(1, L), size=k, replace=False))
        pos[off[c]:off[c + 1]] = p
    ab = np.where(rng.random(n) < 0.3, 0.0, _loguniform(rng, 10.0, 2e4, n)).astype(np.float32)
    ae = (ab.astype(np.float64) + _loguniform(rng, 50.0, 5e4, n)).astype(np.float32)
    flipped = (rng.random(n) < 0.02).astype(np.uint8)
    n_branch = np.where(rng.random(n) < 0.03, 2, 1).astype(np.int32)
    a = rng.integers(0, 4, n)
    d = (a + rng.integers(1, 4, n)) % 4
    anc, der = ACGT[a].copy(), ACGT[d].copy()
    odd = np.zeros(n, dtype=np.uint8)
    if weird > 0:
        w = rng.random(n)
        odd[w < weird * 0.2] = 1
        odd[(w >= weird * 0.2) & (w < weird * 0.3)] = 2
        sw = (w >= weird * 0.3) & (w < weird * 0.5)          # age_begin >= age_end
        ae[sw] = ab[sw]
        ng = (w >= weird * 0.5) & (w < weird * 0.7)          # negative age_begin (clamped to 0)
        ab[ng] = -np.abs(ab[ng]) - np.float32(1.5)
        zo = (w >= weird * 0.7) & (w < weird * 0.85)         # '0'/'1' allele codes
        anc[zo], der[zo] = ord("0"), ord("1")
        bad = (w >= weird * 0.85) & (w < weird)              # allele outside ACGT
        anc[bad] = ord("N")
    return Sites(chr_names, off, pos, ab, ae, flipped, n_branch, anc, der, odd, list(chrom_len))


def add_deep_rows(sites: Sites, seed: int, frac: float = 0.02) -> np.ndarray:
    """Gives a fraction of the rows age intervals that reach past the age grid (bin 185 starts at ~9.3e6 generations)
    with age_begin > 0: for such rows the reference redraws every sample that falls beyond the grid (coal.cpp:2279-2294),
    consuming two more engine words per redraw.  Returns the indices of the modified rows."""
    rng = np.random.default_rng(seed)
    idx = np.nonzero(rng.random(sites.n) < frac)[0]
    ab = rng.uniform(2e6, 8.5e6, idx.shape[0])
    sites.age_begin[idx] = ab.astype(np.float32)
    sites.age_end[idx] = (ab + _loguniform(rng, 5e5, 4e7, idx.shape[0])).astype(np.float32)
    return idx


def make_genome(seed: int, sites: Sites, p_present: float = 0.7, mean_extra_reads: float = 1.0,
                p_derived: float = 0.3, weird: float = 0.0) -> Genome:
    """Record present w.p. ``p_present``; N = 1 + floor(Exp(mean_extra_reads)) reads, each derived
    w.p. ``p_derived``.  ``weird`` adds allele-swapped records, records at positions absent from
    the .mut file and zero-count records."""
    rng = np.random.default_rng(seed)
    n = sites.n
    keep = rng.random(n) < p_present
    idx = np.nonzero(keep)[0]
    k = idx.shape[0]
    chrom = (np.searchsorted(sites.site_off, idx, side="right") - 1).astype(np.int32)
    bp = sites.pos[idx].copy()
    anc = sites.anc[idx].copy()
    der = sites.der[idx].copy()
    N = 1 + np.floor(rng.exponential(mean_extra_reads, k)).astype(np.int64) if mean_extra_reads > 0 \
        else np.ones(k, dtype=np.int64)
    daf = rng.binomial(N, p_derived).astype(np.int32)
    aaf = (N - daf).astype(np.int32)
    if weird > 0:
        w = rng.random(k)
        sw = w < weird * 0.3
        anc[sw], der[sw] = der[sw].copy(), anc[sw].copy()
        z = (w >= weird * 0.3) & (w < weird * 0.5)
        daf[z] = 0
        aaf[(w >= weird * 0.4) & (w < weird * 0.5)] = 0
        # extra records at positions that are not rows of the .mut file
        ex = np.nonzero((w >= weird * 0.5) & (w < weird))[0]
        ebp = bp[ex] + 1
        ok = ~np.isin(ebp.astype(np.int64) + (chrom[ex].astype(np.int64) << 32),
                      sites.pos.astype(np.int64) + ((np.searchsorted(sites.site_off, np.arange(n), side="right") - 1).astype(np.int64) << 32))
        ex, ebp = ex[ok], ebp[ok]
        chrom = np.concatenate([chrom, chrom[ex]])
        bp = np.concatenate([bp, ebp])
        anc = np.concatenate([anc, anc[ex]])
        der = np.concatenate([der, der[ex]])
        aaf = np.concatenate([aaf, aaf[ex]])
        daf = np.concatenate([daf, daf[ex]])
        order = np.lexsort((bp, chrom))
        chrom, bp, anc, der, aaf, daf = (x[order] for x in (chrom, bp, anc, der, aaf, daf))
    return Genome(chrom.astype(np.int32), bp.astype(np.int32), anc.astype(np.uint8), der.astype(np.uint8),
                  aaf.astype(np.int32), daf.astype(np.int32))


def make_mask(seed: int, length: int, frac_n: float, run_lo=1000, run_hi=50000, lower=False) -> bytes:
    """P/N mask with ``frac_n`` of the bases 'N' in runs of run_lo..run_hi bp (config 4)."""
    rng = np.random.default_rng(seed)
    m = np.full(length, ord("P"), dtype=np.uint8)
    target = int(frac_n * length)
    done = 0
    while done < target:
        L = int(rng.integers(run_lo, run_hi + 1))
        s = int(rng.integers(0, max(1, length - L)))
        m[s:s + L] = ord("N")
        done += L
    if lower:
        m = np.where(m == ord("P"), ord("p"), ord("n")).astype(np.uint8)
    return m.tobytes()


def rows_for_genome(total_rows: int, lengths=AUTOSOME_LEN):
    tot = float(sum(lengths))
    rows = [int(total_rows * L / tot) for L in lengths]
    rows[0] += total_rows - sum(rows)
    return rows


# ------------------------------------------------------------------ file writers
def _fmt_f32(x) -> str:
    return np.format_float_positional(np.float32(x), unique=True, trim="-")


def write_mut(path: str, sites: Sites, c: int):
    """Relate .mut text for chromosome index c (mutations.cpp:56-257; writer 298-328)."""
    lo, hi = int(sites.site_off[c]), int(sites.site_off[c + 1])
    with open(path, "w") as f:
        f.write("snp;pos_of_snp;dist;rs-id;tree_index;branch_indices;is_not_mapping;is_flipped;age_begin;age_end;ancestral_allele/alternative_allele;upstream_allele;downstream_allele;\n")
        for i in range(lo, hi):
            k = i - lo
            br = "17" if sites.n_branch[i] == 1 else "17 23"
            if sites.odd[i] == 1:
                mt = chr(sites.anc[i]) + "T/" + chr(sites.der[i])
            elif sites.odd[i] == 2:
                mt = "NA"
            else:
                mt = chr(sites.anc[i]) + "/" + chr(sites.der[i])
            dist = int(sites.pos[i + 1] - sites.pos[i]) if i + 1 < hi else 1
            f.write(f"{k};{int(sites.pos[i])};{dist};rs{k};{k // 3};{br};{int(sites.n_branch[i] > 1)};"
                    f"{int(sites.flipped[i])};{_fmt_f32(sites.age_begin[i])};{_fmt_f32(sites.age_end[i])};{mt};A;C;\n")


_synthio = None


def mut_text_fast(sites: Sites, c: int) -> np.ndarray:
    """The bytes write_mut() writes for chromosome index c, as a uint8 array, produced by the C writer
    (colate_b200/csrc/synth_io.c -> libsynthio.so): 10 M rows in a few seconds instead of minutes."""
    global _synthio
    import ctypes as C
    if _synthio is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libsynthio.so")
        L = C.CDLL(path)
        L.synth_mut_text.restype = C.c_longlong
        L.synth_mut_text.argtypes = [C.c_longlong, C.c_longlong] + [C.c_void_p] * 9 + [C.c_longlong]
        _synthio = L
    lo, hi = int(sites.site_off[c]), int(sites.site_off[c + 1])
    cap = 256 + 200 * (hi - lo + 1)
    out = np.empty(cap, dtype=np.uint8)
    a = [np.ascontiguousarray(x, dtype=dt) for x, dt in ((sites.pos, np.int32), (sites.age_begin, np.float32), (sites.age_end, np.float32),
                                                          (sites.flipped, np.uint8), (sites.n_branch, np.int32), (sites.anc, np.uint8),
                                                          (sites.der, np.uint8), (sites.odd, np.uint8))]
    n = _synthio.synth_mut_text(lo, hi, *[x.ctypes.data for x in a], out.ctypes.data, cap)
    if n < 0:
        raise RuntimeError("synth_mut_text: buffer too small")
    return out[:n]


def mut_texts_fast(sites: Sites, threads: int | None = None):
    """mut_text_fast for every chromosome, on a thread pool (the C writer runs without the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    n = len(sites.chr_names)
    with ThreadPoolExecutor(max_workers=threads or min(n, os.cpu_count() or 1)) as ex:
        return list(ex.map(lambda c: mut_text_fast(sites, c), range(n)))


def colate_in_image(g: Genome, chr_names) -> np.ndarray:
    """The bytes write_colate_in_fast() writes, as a uint8 array."""
    parts = []
    for c, nm in enumerate(chr_names):
        sel = np.nonzero(g.chrom == c)[0]
        if sel.shape[0] == 0:
            continue
        nb = nm.encode()
        rec = np.dtype([("l", "<i4"), ("nm", f"S{len(nb)}"), ("bp", "<i4"), ("a", "u1"), ("d", "u1"), ("aaf", "<i4"), ("daf", "<i4")])
        arr = np.empty(sel.shape[0], dtype=rec)
        arr["l"], arr["nm"], arr["bp"] = len(nb), nb, g.bp[sel]
        arr["a"], arr["d"], arr["aaf"], arr["daf"] = g.anc[sel], g.der[sel], g.aaf[sel], g.daf[sel]
        parts.append(np.frombuffer(arr.tobytes(), dtype=np.uint8))
    return np.concatenate(parts) if parts else np.zeros(0, np.uint8)


def write_colate_in(path: str, g: Genome, chr_names, extra_names=None):
    """.colate.in record stream (coal.cpp:2505-2514): {i32 lchrom, chrom, i32 bp, anc, der, i32 AAF, i32 DAF}."""
    names = list(chr_names) + list(extra_names or [])
    enc = [nm.encode() for nm in names]
    out = bytearray()
    i32 = np.dtype("<i4")
    for k in range(g.n):
        nm = enc[int(g.chrom[k])]
        out += np.array([len(nm)], dtype=i32).tobytes() + nm
        out += np.array([g.bp[k]], dtype=i32).tobytes()
        out += bytes([int(g.anc[k]), int(g.der[k])])
        out += np.array([g.aaf[k], g.daf[k]], dtype=i32).tobytes()
    with open(path, "wb") as f:
        f.write(bytes(out))


def write_colate_in_fast(path: str, g: Genome, chr_names):
    """Vectorised writer for large genomes (all names must have the same byte length per chromosome)."""
    with open(path, "wb") as f:
        for c, nm in enumerate(chr_names):
            sel = np.nonzero(g.chrom == c)[0]
            if sel.shape[0] == 0:
                continue
            nb = nm.encode()
            rec = np.dtype([("l", "<i4"), ("nm", f"S{len(nb)}"), ("bp", "<i4"), ("a", "u1"), ("d", "u1"),
                            ("aaf", "<i4"), ("daf", "<i4")])
            arr = np.empty(sel.shape[0], dtype=rec)
            arr["l"], arr["nm"], arr["bp"] = len(nb), nb, g.bp[sel]
            arr["a"], arr["d"], arr["aaf"], arr["daf"] = g.anc[sel], g.der[sel], g.aaf[sel], g.daf[sel]
            f.write(arr.tobytes())


def write_mask(path: str, seq: bytes, width: int = 60000):
    with open(path, "wb") as f:
        f.write(b">mask\n")
        for i in range(0, len(seq), width):
            f.write(seq[i:i + width] + b"\n")


def write_dataset(dirname: str, sites: Sites, genomes: dict, masks: dict | None = None, prefix="syn"):
    """Writes <prefix>_chr<name>.mut, chr.txt, <gname>.colate.in and <mname>_chr<name>.fa."""
    os.makedirs(dirname, exist_ok=True)
    with open(os.path.join(dirname, "chr.txt"), "w") as f:
        for nm in sites.chr_names:
            f.write(nm + "\n")
    for c, nm in enumerate(sites.chr_names):
        write_mut(os.path.join(dirname, f"{prefix}_chr{nm}.mut"), sites, c)
    for gname, g in genomes.items():
        write_colate_in(os.path.join(dirname, f"{gname}.colate.in"), g, sites.chr_names)
    for mname, per_chr in (masks or {}).items():
        for c, nm in enumerate(sites.chr_names):
            write_mask(os.path.join(dirname, f"{mname}_chr{nm}.fa"), per_chr[c])
    return dirname
