"""Stage i of the whole-genome pair (BASELINE.json configs[1] inputs), three passes on device-resident inputs -- the workload for
the ncu captures of the stage-i kernels (skip the launches of the first two passes)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from colate_b200 import api, synth
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sites = synth.make_sites(1, synth.rows_for_genome(rows), synth.AUTOSOME_LEN)
gt = synth.make_genome(101, sites, 0.7); gr = synth.make_genome(201, sites, 0.7)
h = api.Handle(0)
h.load(sites, gt, gr)
h.set_option("rejoin", 1)
l0 = h.launch_count()
for i in range(passes):
    s1 = h.stage1(api.mt_seed(1), fetch=False)
    print("pass", i, "launches so far", h.launch_count(), {k: round(v, 4) for k, v in h.stage1_timing().items() if k.endswith("_ms")}, flush=True)
print("launches before the passes:", l0, "per pass:", (h.launch_count() - l0) // passes, "rows", sites.n, "used", s1.n_used)
