"""Scratch timing of the three stages at whole-genome scale (device-resident inputs)."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from colate_b200 import api, synth

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1
bins = sys.argv[3] if len(sys.argv) > 3 else "3,7,0.1"
t0 = time.time()
sites = synth.make_sites(1, synth.rows_for_genome(rows), synth.AUTOSOME_LEN)
gt = synth.make_genome(101, sites, 0.7)
gr = synth.make_genome(201, sites, 0.7)
print("synth %.1fs rows %d recs %d %d" % (time.time() - t0, sites.n, gt.n, gr.n), flush=True)
h = api.Handle(0)
t0 = time.time(); h.load(sites, gt, gr); print("load %.3fs" % (time.time() - t0), flush=True)
for it in range(4):
    for g in (0, 1):
        pass
    t0 = time.time()
    s1 = h.stage1(api.mt_seed(1))
    t1 = time.time()
    print("stage1 wall %.2f ms" % ((t1 - t0) * 1e3), h.stage1_timing(), "blocks", s1.num_blocks, flush=True)
w = api.draw_block_weights(s1.mt_state, R, s1.num_blocks)
t0 = time.time(); counts = h.stage2_bootstrap(w, s1.block_stats, 0.0); print("stage2 %.2f ms" % ((time.time() - t0) * 1e3))
ep, _ = api.epochs_from_bins(bins)
for it in range(2):
    t0 = time.time(); rates, iters, ll = h.stage3_em(R, ep, np.full(len(ep), 1 / 20000.)); dt = time.time() - t0
    print("stage3 R=%d E=%d %.1f ms iters min/max %d %d" % (R, len(ep), dt * 1e3, iters.min(), iters.max()), flush=True)
