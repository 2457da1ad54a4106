"""Config-5 style run at scale: G genomes over one whole-genome mutation set, every ordered pair i<j
(colate_b200/pairs.py).  Prints per-stage wall times and the aggregate pair-site evaluations/s.
Under torchrun (`python -m torch.distributed.run --nproc-per-node N tools/all_pairs_scale.py G rows`) the pairs are dealt
round-robin to the ranks (pairs.all_pairs_sharded); rank 0 prints."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from colate_b200 import api, pairs, synth

G = int(sys.argv[1]) if len(sys.argv) > 1 else 16
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
t0 = time.time()
sites = synth.make_sites(1, synth.rows_for_genome(rows), synth.AUTOSOME_LEN)
genomes = [synth.make_genome(1000 + g, sites, 0.7) for g in range(G)]
print("synth %.1f s: %d rows, %d genomes of ~%d records" % (time.time() - t0, sites.n, G, genomes[0].n), flush=True)
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
h = api.Handle(local)
t0 = time.perf_counter()
h.set_sites(sites.site_off, sites.pos, sites.age_begin, sites.age_end, sites.meta())
for g, X in enumerate(genomes):
    h.set_genome(g, X.chrom, X.bp, X.aaf, X.daf, X.anc.astype(np.uint16) | (X.der.astype(np.uint16) << 8))
    h.set_mask(g, None)
t_load = time.perf_counter() - t0
run = (lambda: pairs.all_pairs_sharded(h, G, seed=1, bins="3,7,0.1")) if world > 1 else (lambda: pairs.all_pairs(h, G, seed=1, bins="3,7,0.1"))
res = run()                                                    # first call also joins every genome once
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
res = run()                                                    # steady state: joins cached
t_all = time.perf_counter() - t0
if world > 1:
    t = torch.tensor([t_all], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_all = float(t[0])
    if rank != 0:
        dist.destroy_process_group()
        sys.exit(0)
    print("ranks: %d (pairs dealt round-robin, times are the maximum over the ranks)" % world)
P = res["pairs"].shape[0]
print("load %.2f s | %d pairs: stage i+ii %.3f s (%.2f ms/pair), EM (one launch, throughput mode) %.3f s (%.2f ms/pair), total %.3f s"
      % (t_load, P, res["seconds"]["stage12"], 1e3 * res["seconds"]["stage12"] / P, res["seconds"]["em"],
         1e3 * res["seconds"]["em"] / P, t_all))
print("aggregate %.3e pair-site evaluations/s over %d pairs; iterations min/max %d/%d; used rows/pair ~%d"
      % (P * sites.n / t_all, P, res["iters"].min(), res["iters"].max(), int(res["n_used"].mean())))
# spot check: pair 0 alone through api.mut (latency-mode EM) gives the same bits
h.set_option("rejoin", 0)
one = api.mut(h, 1, bins="3,7,0.1") if tuple(res["pairs"][0]) == (0, 1) else None   # genomes 0 / 1 sit in slots 0 / 1
if one is not None:
    print("pair (0,1) alone == batched:", bool(np.array_equal(one["rates"][0], res["rates"][0]) and one["iters"][0] == res["iters"][0]))
if world > 1:
    dist.destroy_process_group()
