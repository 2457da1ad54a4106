import sys, time, os, numpy as np
sys.path.insert(0, "/root/repo")
from colate_b200 import api, synth
from oracle import pyoracle as po
rows = 2_000_000
sites = synth.make_sites(1, synth.rows_for_genome(rows), synth.AUTOSOME_LEN)
gt = synth.make_genome(101, sites, 0.7); gr = synth.make_genome(201, sites, 0.7)
h = api.Handle(0); h.load(sites, gt, gr)
s1 = h.stage1(api.mt_seed(1))
ep, _ = api.epochs_from_bins("3,7,0.1"); init = np.full(len(ep), 1/20000.)
Rs = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1, 8, 37, 125, 148, 250, 1000]
kerns = sys.argv[2].split(",") if len(sys.argv) > 2 else ["default", "cta", "task"]
for R in Rs:
    st = s1.mt_state.copy()
    w = api.draw_block_weights(st, R, s1.num_blocks)
    h.stage2_bootstrap(w, s1.block_stats, 0.0)
    for kern in kerns:
        if kern == "default": os.environ.pop("COLATE_EM_KERNEL", None)
        else: os.environ["COLATE_EM_KERNEL"] = kern
        h.stage3_em(R, ep, init)
        t0 = time.time(); rates, iters, ll = h.stage3_em(R, ep, init); dt = time.time() - t0
        print(f"R={R} kernel={kern} {dt*1e3:.1f} ms iters {iters.min()}..{iters.max()} -> {iters.sum()/dt:.3e} replicate-iterations/s", flush=True)
os.environ["COLATE_EM_PROF"] = "1"
h.stage3_em(Rs[-1], ep, init)
