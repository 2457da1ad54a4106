"""CPU: the multi-GPU host logic (colate_b200/dist.py) under gloo with world_size 2.  The compute
backend is replaced by an oracle-backed stand-in (test infrastructure), so what is tested is the
sharding: chromosome split, generator-stream offsets, block bases, zero-padded all-reduce,
replicate assignment and gathers.  The result must be bit-identical to the single-process run."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackend:
    """Stand-in for api.Handle on one rank: same stage contract, computed by the CPU oracle."""

    def __init__(self, sites, gt, gr, chr_lo, chr_hi):
        from colate_b200 import synth
        lo, hi = int(sites.site_off[chr_lo]), int(sites.site_off[chr_hi])
        off = sites.site_off[chr_lo:chr_hi + 1] - sites.site_off[chr_lo]
        self.sites = synth.Sites(sites.chr_names[chr_lo:chr_hi], off, sites.pos[lo:hi], sites.age_begin[lo:hi], sites.age_end[lo:hi],
                                 sites.flipped[lo:hi], sites.n_branch[lo:hi], sites.anc[lo:hi], sites.der[lo:hi], sites.odd[lo:hi],
                                 sites.chrom_len[chr_lo:chr_hi])

        def sub(g):  # records of the rank's chromosomes, renumbered (the seek only ever looks at listed names)
            ch = g.chrom.astype(np.int64) - chr_lo
            ch = np.where((ch >= 0) & (ch < chr_hi - chr_lo), ch, chr_hi - chr_lo + 5)
            return synth.Genome(ch.astype(np.int32), g.bp, g.anc, g.der, g.aaf, g.daf)
        self.gt, self.gr = sub(gt), sub(gr)
        self._counts = None

    def flags(self):
        from oracle import pyoracle as po
        used, blocks = [], []
        if not self.sites.chr_names:                    # a rank that owns no chromosome (world > n_chr)
            return np.zeros(0, np.int64), np.zeros(0, np.int32)
        for c in range(len(self.sites.chr_names)):
            o = OracleBackend._one(self, c)
            used.append(o["n_used_total"]); blocks.append(o["num_blocks"])
        return np.array(used, np.int64), np.array(blocks, np.int32)

    def _one(self, c):
        from colate_b200 import synth
        from oracle import pyoracle as po
        s = self.sites
        lo, hi = int(s.site_off[c]), int(s.site_off[c + 1])
        one = synth.Sites([s.chr_names[c]], np.array([0, hi - lo]), s.pos[lo:hi], s.age_begin[lo:hi], s.age_end[lo:hi], s.flipped[lo:hi],
                          s.n_branch[lo:hi], s.anc[lo:hi], s.der[lo:hi], s.odd[lo:hi], [s.chrom_len[c]])

        def sub(g):
            ch = np.where(g.chrom == c, 0, 7).astype(np.int32)
            return synth.Genome(ch, g.bp, g.anc, g.der, g.aaf, g.daf)
        return po.stage1(one, sub(self.gt), sub(self.gr), seed=1)

    def sample(self, mt_state, used_rank_base, block_base, n_blocks):
        from colate_b200 import api
        from oracle import pyoracle as po
        if not self.sites.chr_names:
            after = api.mt_seed(self._seed)
            api.lib().colate_mt_generate(after, 200 * used_rank_base, np.zeros(max(1, 200 * used_rank_base), np.uint32))
            assert n_blocks == 0
            return np.zeros((0, 4, 185)), np.zeros((0, 3, 185), np.int64), after
        # an oracle generator positioned at this rank's offset in the reference's stream
        g = po.mt_seed(self._seed)
        for _ in range(200 * used_rank_base):
            po.lib().oracle_mt_next(g)
        o = po.stage1(self.sites, self.gt, self.gr, rng=g)
        nb = o["num_blocks"]
        stats = np.stack([o["shared"], o["notshared"], o["shared_emp"], o["notshared_emp"]], axis=1)
        tallies = np.stack([o["n_shared"], o["n_notshared"], o["n_emp"]], axis=1)
        after = api.mt_seed(self._seed)
        burn = np.zeros(max(1, 200 * (used_rank_base + o["n_used_total"])), np.uint32)
        api.lib().colate_mt_generate(after, 200 * (used_rank_base + o["n_used_total"]), burn)
        assert nb == n_blocks
        return stats, tallies, after

    def bootstrap(self, weights, block_stats, age):
        from oracle import pyoracle as po
        blk = {k: np.ascontiguousarray(block_stats[:, i]) for i, k in enumerate(("shared", "notshared", "shared_emp", "notshared_emp"))}
        return po.stage2(np.ascontiguousarray(weights), blk, age)

    def em(self, R, epochs, rates_init, counts, max_iter):
        from oracle import pyoracle as po
        out = [po.em_run(epochs, rates_init, counts[r], max_iter) for r in range(R)]
        return np.stack([o[0] for o in out]), np.array([o[1] for o in out]), np.array([o[2] for o in out])


ROWS = {"even": ([900, 600, 1400, 700, 1100], [2.4e8, 5e7, 1.3e8, 6.1e7, 9e7]),
        "skewed": ([40, 40, 2500], [5e7, 5e7, 2.4e8]),          # everything nearest the target sits in the last chromosome
        "two": ([800, 1200], [9e7, 2.4e8])}                     # fewer chromosomes than ranks at world 3


def _worker(rank, world, port, seed, R, q, shape="even"):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from colate_b200 import dist as cdist, synth
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sites = synth.make_sites(seed, *ROWS[shape], weird=0.05)
    gt = synth.make_genome(seed + 100, sites, 0.8)
    gr = synth.make_genome(seed + 200, sites, 0.8)
    parts = cdist.split_chromosomes(np.diff(sites.site_off), world)
    be = OracleBackend(sites, gt, gr, *parts[rank])
    be._seed = seed
    res = cdist.mut_sharded(be, seed, bins="3,7,0.2", num_bootstraps=R, max_iter=30, device="cpu")
    stats, tallies = res.stats()
    q.put((rank, res.num_blocks, res.n_used, stats, tallies, res.mt_state, res.rates, res.iters))
    dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.mark.parametrize("R,world,shape", [(1, 2, "even"), (5, 2, "even"), (3, 2, "skewed"), (2, 3, "two")])
def test_chromosome_and_replicate_sharding(built, R, world, shape):
    """world 2 on five chromosomes; a skewed split (no rank may end up empty while world <= n_chr); three
    ranks on two chromosomes (one rank owns nothing and still joins every exchange)."""
    from colate_b200 import synth
    from colate_b200.dist import split_chromosomes
    from oracle import pyoracle as po
    seed = 4
    parts = split_chromosomes(ROWS[shape][0], world)
    if world <= len(ROWS[shape][0]):
        assert all(hi > lo for lo, hi in parts), parts
    else:
        assert sum(hi > lo for lo, hi in parts) == len(ROWS[shape][0])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, seed, R, q, shape)) for r in range(world)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=150) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # single-process run of the whole thing
    sites = synth.make_sites(seed, *ROWS[shape], weird=0.05)
    gt = synth.make_genome(seed + 100, sites, 0.8)
    gr = synth.make_genome(seed + 200, sites, 0.8)
    o = po.stage1(sites, gt, gr, seed=seed)
    w = po.draw_block_weights(o["rng"], R, o["num_blocks"])
    counts = po.stage2(w, o, 0.0)
    ep, _ = po.epochs_from_bins("3,7,0.2")
    want = np.stack([po.em_run(ep, np.full(len(ep), 1 / 20000.), counts[r], 30)[0] for r in range(R)])
    ref_stats = np.stack([o["shared"], o["notshared"], o["shared_emp"], o["notshared_emp"]], axis=1)
    for rank, nb, nu, stats, tallies, st, rates, iters in outs:
        assert nb == o["num_blocks"] and nu == o["n_used_total"]
        assert np.array_equal(stats, ref_stats)                       # bit-identical to one process
        assert np.array_equal(tallies[:, 1], o["n_notshared"])
        assert np.array_equal(rates, want) and (iters == 30).all()
    for o2 in outs[1:]:
        assert np.array_equal(outs[0][5], o2[5])                      # same generator state everywhere


def test_split_chromosomes():
    from colate_b200.dist import split_chromosomes
    assert split_chromosomes([10, 10, 10, 10], 2) == [(0, 2), (2, 4)]
    assert split_chromosomes([5], 4) == [(0, 1), (1, 1), (1, 1), (1, 1)]
    assert split_chromosomes([1, 1, 100], 2) == [(0, 2), (2, 3)]                       # was [(0, 3), (3, 3)]
    assert split_chromosomes([10, 10, 10, 1000], 4) == [(0, 1), (1, 2), (2, 3), (3, 4)]
    for w in (1, 2, 3, 8):
        parts = split_chromosomes([249, 243, 198, 191, 180, 171, 159, 146, 141, 135, 135, 133, 115, 107, 102, 90, 81, 78, 59, 63, 48, 51], w)
        assert parts[0][0] == 0 and parts[-1][1] == 22 and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))


class OracleHandle:
    """Stand-in for api.Handle with genomes in slots (the contract pairs.all_pairs uses), computed by the CPU oracle."""

    def __init__(self, sites, genomes):
        self.sites, self.genomes = sites, genomes

    def set_stream_cache(self, enable):
        pass

    def stage1(self, mt_state, target_slot=0, reference_slot=1, fetch=True):
        from types import SimpleNamespace
        from colate_b200 import api
        from oracle import pyoracle as po
        o = po.stage1(self.sites, self.genomes[target_slot], self.genomes[reference_slot], seed=self.seed)
        after = api.mt_seed(self.seed)
        burn = np.zeros(max(1, 200 * o["n_used_total"]), np.uint32)
        api.lib().colate_mt_generate(after, 200 * o["n_used_total"], burn)
        stats = np.stack([o["shared"], o["notshared"], o["shared_emp"], o["notshared_emp"]], axis=1)
        self._resident = stats                      # api.Handle keeps the block histograms on the device (fetch=False)
        return SimpleNamespace(num_blocks=o["num_blocks"], n_used=o["n_used_total"], block_stats=stats if fetch else None, mt_state=after)

    def stage2_bootstrap_dev(self, w, stats_ptr, num_blocks, age=0.0, fetch=False):
        assert stats_ptr is None and num_blocks == self._resident.shape[0]
        return self.stage2_bootstrap(w, self._resident, age)

    def stage2_bootstrap(self, w, block_stats, age):
        from oracle import pyoracle as po
        blk = {k: np.ascontiguousarray(block_stats[:, i]) for i, k in enumerate(("shared", "notshared", "shared_emp", "notshared_emp"))}
        return po.stage2(np.ascontiguousarray(w), blk, age)

    def stage3_em(self, R, epochs, init, counts, max_iter):
        from oracle import pyoracle as po
        out = [po.em_run(epochs, init, counts[r], max_iter) for r in range(R)]
        return np.stack([o[0] for o in out]), np.array([o[1] for o in out], np.int32), np.array([o[2] for o in out])


def _pairs_inputs(seed):
    from colate_b200 import synth
    sites = synth.make_sites(seed, [700, 500], [2.4e8, 1.3e8])
    return sites, [synth.make_genome(seed + 50 + g, sites, 0.8) for g in range(4)]


def _pairs_worker(rank, world, port, seed, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from colate_b200 import pairs as pairs_mod
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sites, genomes = _pairs_inputs(seed)
    h = OracleHandle(sites, genomes)
    h.seed = seed
    res = pairs_mod.all_pairs_sharded(h, len(genomes), seed, bins="3,7,0.2", device="cpu", max_iter=20)
    q.put((rank, res["pairs"], res["rates"], res["iters"], res["ll"], res["num_blocks"], res["n_used"], res["counts"]))
    dist.destroy_process_group()


def test_all_pairs_sharded_world2(built):
    """configs[4] host logic: pairs dealt round-robin to two ranks, tables combined -- identical to one process."""
    from colate_b200 import pairs as pairs_mod
    seed, world = 6, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pairs_worker, args=(r, world, port, seed, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=150) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    sites, genomes = _pairs_inputs(seed)
    h = OracleHandle(sites, genomes)
    h.seed = seed
    one = pairs_mod.all_pairs(h, len(genomes), seed, bins="3,7,0.2", max_iter=20)
    assert one["pairs"].shape == (6, 2) and len({int(x) for x in one["n_used"]}) > 1
    for rank, prs, rates, iters, ll, nb, nu, counts in outs:
        assert np.array_equal(prs, one["pairs"])
        assert np.array_equal(rates.view(np.int64), one["rates"].view(np.int64))       # bit patterns
        assert np.array_equal(ll.view(np.int64), one["ll"].view(np.int64))
        assert np.array_equal(iters, one["iters"]) and np.array_equal(nb, one["num_blocks"]) and np.array_equal(nu, one["n_used"])
        assert np.array_equal(counts, one["counts"])


def _shared_text_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from colate_b200 import pairs
    rng = np.random.default_rng(3)                     # the SAME texts on every rank (the cohort's .mut files)
    texts = [rng.integers(32, 127, size=n, dtype=np.uint8) for n in (1000, 1, 777, 4096, 255, 256, 257)]
    st = pairs.SharedMutText(texts, torch.device("cpu"), world, rank)
    ptr_sizes = st.exchange()
    got = [st.dev.numpy()[p - st.dev.data_ptr():p - st.dev.data_ptr() + n].copy() for p, n in ptr_sizes]
    q.put((rank, all(np.array_equal(a, b) for a, b in zip(got, texts)), st.h2d_bytes, [int(p - st.dev.data_ptr()) % 256 for p, _ in ptr_sizes]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_shared_mut_text_exchange(built, world):
    """pairs.SharedMutText: every rank contributes its 1 / world slice of the concatenated .mut texts, the all-gather leaves the
    whole text on every rank, every chromosome's text at a 256-byte aligned offset (gloo stand-in for the NVLink all-gather)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shared_text_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs: p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs: p.join(timeout=60)
    assert sorted(r[0] for r in res) == list(range(world))
    for rank, ok, h2d, align in res:
        assert ok, rank
        assert all(a == 0 for a in align)
        assert h2d * world >= 1000 + 1 + 777 + 4096 + 255 + 256 + 257 and h2d % 256 == 0
