"""Small end-to-end pass for compute-sanitizer (memcheck / racecheck / synccheck): text ingest, record decode, stage i
(k_sample's bulk-copy ring, k_replay's TMA ring), stage ii, and all three EM kernels (k_em_split on a 16-CTA cluster with
its st.async / mbarrier handshake, k_em_cta, k_em) for a few iterations.  Checked against the oracle so that a sanitizer run
is also a correctness run.   compute-sanitizer --tool racecheck python tools/sanitize_smoke.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from colate_b200 import api, synth
from oracle import pyoracle as po

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 12
sites = synth.make_sites(1, [2500, 1500], [2.49e8, 1.2e8], weird=0.05)
synth.add_deep_rows(sites, 3, 0.01)
gt = synth.make_genome(101, sites, 0.7); gr = synth.make_genome(201, sites, 0.7)
h = api.Handle(0)
h.ingest_mut([synth.mut_text_fast(sites, c).tobytes() for c in range(2)])
for slot, g in ((0, gt), (1, gr)):
    h.ingest_colate_in(slot, synth.colate_in_image(g, sites.chr_names).tobytes(), sites.chr_names)
s1 = h.stage1(api.mt_seed(1))
o = po.stage1(sites, gt, gr, seed=1)
assert s1.n_used == o["n_used_total"] and np.array_equal(s1.block_stats[:, 0], o["shared"]) and np.array_equal(s1.block_stats[:, 1], o["notshared"])
ep, _ = api.epochs_from_bins("3,7,0.2")
init = np.full(len(ep), 1 / 20000.)
for R, env in ((1, {}), (12, {}), (12, {"COLATE_EM_KERNEL": "task"}), (3, {"COLATE_EM_CLUSTER": "8"})):
    for k in ("COLATE_EM_KERNEL", "COLATE_EM_CLUSTER"):
        os.environ.pop(k, None)
    os.environ.update(env)
    st = s1.mt_state.copy()
    w = api.draw_block_weights(st, R, s1.num_blocks)
    counts = h.stage2_bootstrap(w, s1.block_stats, 0.0)
    rates, it, ll = h.stage3_em(R, ep, init, None, max_iter=iters)
    ro, ito, _ = po.em_run(ep, init, counts[R - 1], iters)
    assert it[R - 1] == ito and np.array_equal(rates[R - 1], ro), (R, env)
h.close()
print("sanitize_smoke ok: n_used", s1.n_used, "extra words", "EM iterations", iters)
