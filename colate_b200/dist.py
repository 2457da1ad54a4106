"""Multi-GPU driver of the path: one process per GPU (torch.distributed, NCCL over NVLink; gloo in
CPU tests), mirroring SURVEY.md 8(e).

  sites      sharded by chromosome: rank r owns a contiguous range of the --chr list.  The only
             exchanges are (1) an all-gather of two integers per rank (used rows, genomic blocks:
             they give every rank its offset into the reference's generator stream and its first
             block) and (2) one all-reduce of the zero-padded [500, 4, 185] block histograms
             (adding zeros is exact, so the result is bit-identical to a single-GPU run);
  replicates sharded round-robin for the bootstrap + EM; every rank draws the full weight table from
             the same generator state (it is the reference's stream), results are all-gathered;
  pairs      independent: no collective (bench.py --gpus N).

The compute backend is anything with the four stage methods of `api.Handle`; tests inject a
CPU stand-in to exercise this host logic under gloo with world_size 2.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import api
from ._lib import MAX_BLOCKS, NBINS


CALLS: list = []   # collectives issued since the last clear(): (name, shape, dtype) -- bench.py prints them


def _log(name, t):
    CALLS.append([name, list(t.shape), str(t.dtype).replace("torch.", "")])


def split_chromosomes(rows_per_chr, world: int):
    """Contiguous ranges [lo, hi) of the --chr list per rank, balanced by row count: boundary r is the cumulative row
    count nearest to r/world of the total, moved if necessary so that every rank keeps at least one chromosome
    (world <= n_chr).  With more ranks than chromosomes the first n_chr ranks own one chromosome each and the rest
    own none: an empty rank still joins every exchange with zero used rows and zero blocks."""
    rows = np.asarray(rows_per_chr, dtype=np.int64)
    n = rows.shape[0]
    cum = np.concatenate([[0], np.cumsum(rows)])
    total = cum[-1]
    bounds = [0]
    for r in range(1, world):
        lo = min(n, bounds[-1] + 1)                    # at least one chromosome for rank r - 1 ...
        hi = max(lo, n - (world - r))                  # ... and for each of the ranks r .. world - 1
        hi = min(hi, n)
        target = total * r / world
        c = lo + int(np.argmin(np.abs(cum[lo:hi + 1] - target)))
        bounds.append(c)
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def load_shard(handle: api.Handle, sites, target, reference, world: int, rank: int):
    """This rank's chromosomes of a dataset (synth.Sites / synth.Genome) into `handle`: the rows of its contiguous
    range of the --chr list, and both genomes with the record ranges of those chromosomes -- the chromosome seek
    (coal.cpp:2125-2145) is emulated on the WHOLE record stream, a rank only keeps the ranges of its own names."""
    lo, hi = split_chromosomes(np.diff(sites.site_off), world)[rank]
    s0, s1 = int(sites.site_off[lo]), int(sites.site_off[hi])
    meta = sites.meta()
    handle.set_sites(sites.site_off[lo:hi + 1] - sites.site_off[lo], sites.pos[s0:s1], sites.age_begin[s0:s1], sites.age_end[s0:s1], meta[s0:s1])
    for slot, g in ((0, target), (1, reference)):
        first, end = api.chr_ranges(len(sites.chr_names), g.chrom)
        al = g.anc.astype(np.uint16) | (g.der.astype(np.uint16) << 8)
        api.check(api.lib().colate_set_genome(handle._h, slot, g.n, api.ptr(np.ascontiguousarray(first[lo:hi])),
                                              api.ptr(np.ascontiguousarray(end[lo:hi])), api.ptr(np.ascontiguousarray(g.bp)),
                                              api.ptr(np.ascontiguousarray(g.aaf)), api.ptr(np.ascontiguousarray(g.daf)), api.ptr(al), 0))
    return lo, hi


class CudaBackend:
    """The product path: api.Handle on this rank's GPU.  The block histograms stay on the device from the sampling
    kernels through the all-reduce into the bootstrap kernel (`sample_into` / `bootstrap_dev`)."""

    def __init__(self, handle: api.Handle):
        self.h = handle

    def flags(self):
        return self.h.stage1_flags()

    def sample_into(self, mt_state, used_rank_base, block_base, n_blocks, stats_t, tallies_t):
        """Rows [block_base, block_base + n_blocks) of the zero-padded device tensors receive this rank's blocks."""
        return self.h.stage1_sample_dev(mt_state, used_rank_base, block_base,
                                        stats_t.data_ptr() + block_base * 4 * NBINS * 8 if n_blocks else None,
                                        tallies_t.data_ptr() + block_base * 3 * NBINS * 8 if n_blocks else None)

    def extra_words(self):
        return self.h.extra_words()

    def bootstrap_dev(self, weights, stats_t, num_blocks, age):
        self.h.stage2_bootstrap_dev(weights, stats_t.data_ptr(), num_blocks, age)

    def em(self, R, epochs, rates_init, counts, max_iter):
        return self.h.stage3_em(R, epochs, rates_init, counts, max_iter)


@dataclass
class DistResult:
    num_blocks: int
    n_used: int
    block_stats: np.ndarray | None     # [num_blocks, 4, 185], identical on every rank (fetched on first use: .stats())
    block_tallies: np.ndarray | None   # [num_blocks, 3, 185]
    mt_state: np.ndarray         # generator state after stage i, identical on every rank
    rates: np.ndarray | None     # [R, E] on every rank
    iters: np.ndarray | None
    epochs: np.ndarray | None
    ep_null: int = 0
    seconds: dict = field(default_factory=dict)   # this rank's wall time of the EM launch
    stats_dev: object = None     # torch [500, 4, 185] f64 / [500, 3, 185] i64 on the compute device
    tallies_dev: object = None

    def stats(self):
        if self.block_stats is None:
            self.block_stats = self.stats_dev[:self.num_blocks].cpu().numpy()
            self.block_tallies = self.tallies_dev[:self.num_blocks].cpu().numpy()
        return self.block_stats, self.block_tallies


def _dist():
    import torch
    import torch.distributed as dist
    return torch, dist


def stage1_sharded(backend, seed_state: np.ndarray, device="cuda"):
    """Stage i with the sites sharded by chromosome.  Every rank returns the full block histograms."""
    torch, dist = _dist()
    world, rank = dist.get_world_size(), dist.get_rank()
    used_chr, blocks_chr = backend.flags()
    mine = torch.tensor([int(np.sum(used_chr)), int(np.sum(blocks_chr))], dtype=torch.int64, device=device)
    allv = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    _log("all_gather", mine)
    allv = np.stack([v.cpu().numpy() for v in allv])           # [world, 2]
    used_base = int(allv[:rank, 0].sum())
    block_base = int(allv[:rank, 1].sum())
    n_used, num_blocks = int(allv[:, 0].sum()), int(allv[:, 1].sum())
    if num_blocks > MAX_BLOCKS:
        raise api._lib.ColateError(-3, "more than 500 genomic blocks")
    n_local = int(allv[rank, 1])
    pad = torch.zeros((MAX_BLOCKS, 4, NBINS), dtype=torch.float64, device=device)
    pad_n = torch.zeros((MAX_BLOCKS, 3, NBINS), dtype=torch.int64, device=device)
    if hasattr(backend, "sample_into"):
        if pad.is_cuda:
            torch.cuda.current_stream().synchronize()          # the zero fill, before the handle's stream writes into the tensors
        state_after = backend.sample_into(seed_state, used_base, block_base, n_local, pad, pad_n)
    else:
        stats, tallies, state_after = backend.sample(seed_state, used_base, block_base, n_local)
        if n_local:
            pad[block_base:block_base + n_local] = torch.from_numpy(np.ascontiguousarray(stats[:n_local])).to(device)
            pad_n[block_base:block_base + n_local] = torch.from_numpy(np.ascontiguousarray(tallies[:n_local])).to(device)
    if hasattr(backend, "extra_words"):
        # rows the reference rejection-samples (coal.cpp:2279-2294) consume extra generator words: the stream offsets of
        # the ranks behind would depend on them -- such inputs run on one GPU
        ex = torch.tensor([backend.extra_words()], dtype=torch.int64, device=device)
        dist.all_reduce(ex, op=dist.ReduceOp.MAX)
        _log("all_reduce(max)", ex)
        if int(ex.item()) != 0:
            raise api._lib.ColateError(-2, "rows with redrawn samples (age interval beyond the age grid, coal.cpp:2279-2294) "
                                           "cannot be sharded by chromosome: run this input on one GPU")
    dist.all_reduce(pad, op=dist.ReduceOp.SUM)                 # disjoint supports: x + 0.0 == x
    dist.all_reduce(pad_n, op=dist.ReduceOp.SUM)
    _log("all_reduce(sum)", pad); _log("all_reduce(sum)", pad_n)
    # the generator state after the last used row lives on the last rank
    st = torch.from_numpy(state_after.astype(np.int64)).to(device)
    dist.broadcast(st, src=world - 1)
    _log("broadcast(src=last)", st)
    res = DistResult(num_blocks, n_used, None, None, st.cpu().numpy().astype(np.uint32), None, None, None)
    res.stats_dev, res.tallies_dev = pad, pad_n                # [500, ...] device tensors, identical on every rank
    return res


def em_sharded(backend, res: DistResult, num_bootstraps: int, epochs, rates_init, age=0.0, max_iter=100000, device="cuda"):
    """Stages ii + iii with the replicates dealt round-robin to the ranks; all-gathers the rates."""
    torch, dist = _dist()
    world, rank = dist.get_world_size(), dist.get_rank()
    state = res.mt_state.copy()
    w = api.draw_block_weights(state, num_bootstraps, res.num_blocks)   # same table on every rank
    mine = np.arange(rank, num_bootstraps, world)
    E = len(epochs)
    rates = torch.zeros((num_bootstraps, E), dtype=torch.float64, device=device)
    iters = torch.zeros(num_bootstraps, dtype=torch.int64, device=device)
    import time
    t_em = 0.0
    if mine.shape[0]:
        if hasattr(backend, "bootstrap_dev"):
            if res.stats_dev.is_cuda:
                torch.cuda.current_stream().synchronize()      # the all-reduce, before the handle's stream reads the tensor
            backend.bootstrap_dev(np.ascontiguousarray(w[mine]), res.stats_dev, res.num_blocks, age)
            counts = None
        else:
            counts = backend.bootstrap(np.ascontiguousarray(w[mine]), res.stats()[0], age)
        t0 = time.perf_counter()
        r, it, _ = backend.em(mine.shape[0], epochs, rates_init, counts, max_iter)      # synchronous: returns with the rates on the host
        t_em = time.perf_counter() - t0
        idx = torch.from_numpy(mine).to(device)
        rates[idx] = torch.from_numpy(np.ascontiguousarray(r)).to(device)
        iters[idx] = torch.from_numpy(np.asarray(it, dtype=np.int64)).to(device)
    dist.all_reduce(rates, op=dist.ReduceOp.SUM)               # disjoint supports again
    dist.all_reduce(iters, op=dist.ReduceOp.SUM)
    _log("all_reduce(sum)", rates); _log("all_reduce(sum)", iters)
    res.rates, res.iters, res.epochs = rates.cpu().numpy(), iters.cpu().numpy().astype(np.int32), np.asarray(epochs)
    res.seconds["em"] = t_em
    return res


def mut_sharded(backend, seed: int, bins: str = "3,7,0.2", num_bootstraps: int = 1, target_age=None, reference_age=None,
                years_per_gen=None, max_iter=100000, device="cuda") -> DistResult:
    """mut() (coal.cpp:3071-3863) across the ranks of the default process group."""
    age, ypg = api.ages(target_age, reference_age, years_per_gen)
    res = stage1_sharded(backend, api.mt_seed(seed), device)
    epochs, ep_null = api.epochs_from_bins(bins, age, ypg)
    res = em_sharded(backend, res, num_bootstraps, epochs, np.full(len(epochs), 1.0 / 20000.0), age, max_iter, device)
    res.ep_null = ep_null
    return res
