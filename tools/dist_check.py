"""Run under torchrun (NCCL): the chromosome/replicate-sharded path on N GPUs must equal the
single-GPU result bit for bit.  rank 0 prints PASS/FAIL."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from colate_b200 import api, synth, dist as cdist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
R = int(sys.argv[2]) if len(sys.argv) > 2 else 16
sites = synth.make_sites(1, synth.rows_for_genome(rows), synth.AUTOSOME_LEN)
gt = synth.make_genome(101, sites, 0.7); gr = synth.make_genome(201, sites, 0.7)
lo, hi = cdist.split_chromosomes(np.diff(sites.site_off), world)[rank]
s0, s1 = int(sites.site_off[lo]), int(sites.site_off[hi])
h = api.Handle(local)
h.set_sites(sites.site_off[lo:hi + 1] - sites.site_off[lo], sites.pos[s0:s1], sites.age_begin[s0:s1], sites.age_end[s0:s1], sites.meta()[s0:s1])
for slot, g in ((0, gt), (1, gr)):
    first, end = api.chr_ranges(len(sites.chr_names), g.chrom)          # the seek is emulated on the WHOLE file
    al = g.anc.astype(np.uint16) | (g.der.astype(np.uint16) << 8)
    api.check(api.lib().colate_set_genome(h._h, slot, g.n, api.ptr(np.ascontiguousarray(first[lo:hi])), api.ptr(np.ascontiguousarray(end[lo:hi])),
                                          api.ptr(g.bp), api.ptr(g.aaf), api.ptr(g.daf), api.ptr(al), 0))
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
res = cdist.mut_sharded(cdist.CudaBackend(h), seed=1, bins="3,7,0.1", num_bootstraps=R, device="cuda")   # cold: allocations, NCCL connections
torch.cuda.synchronize(); dist.barrier()
e0.record()
res = cdist.mut_sharded(cdist.CudaBackend(h), seed=1, bins="3,7,0.1", num_bootstraps=R, device="cuda")
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device="cuda"); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    h1 = api.Handle(local); h1.load(sites, gt, gr)
    one = api.mut(h1, seed=1, bins="3,7,0.1", num_bootstraps=R)
    ok = (res.num_blocks == one["num_blocks"] and np.array_equal(res.stats()[0], one["stage1"].block_stats)
          and np.array_equal(res.rates, one["rates"]) and np.array_equal(res.iters, one["iters"]))
    print(f"dist_check world={world} rows={rows} R={R}: {'PASS' if ok else 'FAIL'} (bit-identical to 1 GPU), sharded pass (second, warm) {ms.item():.1f} ms", flush=True)
dist.barrier(); dist.destroy_process_group()
