import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from colate_b200 import api, synth
from oracle import pyoracle as po
sites = synth.make_sites(1, [30000, 20000], [2.5e8, 1.2e8])
gt = synth.make_genome(101, sites, 0.7); gr = synth.make_genome(201, sites, 0.7)
o = po.stage1(sites, gt, gr, seed=1)
h = api.Handle(0)
counts = po.stage2(np.ones((1, o["num_blocks"]), np.int32), o, 0.0)
ep, _ = po.epochs_from_bins("3,7,0.2", 0.0, 28.0)
init = np.full(len(ep), 1 / 20000.)
np.set_printoptions(linewidth=200, precision=4)
for mi in (1, 2, 5, 50, 500, 100000):
    rates, iters, ll = h.stage3_em(1, ep, init, counts, max_iter=mi)
    ro, it, llo = po.em_run(ep, init, counts[0], max_iter=mi)
    with np.errstate(all="ignore"):
        rel = np.abs(rates[0] - ro) / np.abs(ro)
    print("max_iter", mi, "iters", iters[0], it, "ll", ll[0], llo, "rel ll %.2e" % abs(ll[0] / llo - 1))
    print("   maxrel %.3e" % np.nanmax(rel), "argmax", np.nanargmax(rel))
    print("   rel", rel)
print("rates oracle", ro)
print("counts S", counts[0, 0][:40])
