"""Whole CLI runs at BASELINE sizes on the GPU box: generated files -> `colate_b200/bin/Colate` vs the unmodified
reference CLI (`oracle/_ref/Colate`) on the same files: wall clocks and byte-for-byte comparison of the .coal files.
usage: cli_wall.py [rows] [n_chr] [num_bootstraps] [--skip-reference] [--devices=0,1,...]
--devices: also run the CLI with chromosomes and replicates dealt to several GPUs and compare its .coal byte for byte."""
import os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from colate_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
n_chr = int(sys.argv[2]) if len(sys.argv) > 2 else 22
R = int(sys.argv[3]) if len(sys.argv) > 3 else 1
skip_ref = "--skip-reference" in sys.argv
lens = synth.AUTOSOME_LEN[:n_chr]
per = [int(rows * L / sum(lens)) for L in lens]
t0 = time.time()
sites = synth.make_sites(1, per, lens)
gt = synth.make_genome(101, sites, 0.7)
gr = synth.make_genome(201, sites, 0.7)
d = tempfile.mkdtemp(prefix="colate_wall_")
synth.write_dataset(d, sites, {"t": gt, "r": gr})
size = sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d))
print("dataset: %d rows on %d chromosomes, %d + %d records, %.0f MB of files, written in %.0f s" % (sites.n, n_chr, gt.n, gr.n, size / 1e6, time.time() - t0), flush=True)
common = ["--mode", "mut", "--mut", d + "/syn", "--chr", d + "/chr.txt", "--target_tmp", d + "/t.colate.in", "--reference_tmp", d + "/r.colate.in",
          "--bins", "3,7,0.1", "--seed", "1", "--num_bootstraps", str(R)]
def run(exe, out, extra=()):
    t0 = time.perf_counter()
    r = subprocess.run([exe] + common + list(extra) + ["-o", out], capture_output=True, text=True)
    dt = time.perf_counter() - t0
    assert r.returncode == 0, r.stderr[-2000:]
    return dt
ours = os.path.join(ROOT, "colate_b200", "bin", "Colate")
run(ours, d + "/warm")                                       # first run pays CUDA context creation + page cache
os.environ["COLATE_TIMING"] = "1"; r = subprocess.run([ours] + common + ["-o", d + "/t"], capture_output=True, text=True); print("".join(l + "\n" for l in r.stderr.splitlines() if l.startswith("[timing]"))); del os.environ["COLATE_TIMING"]
t_gpu = min(run(ours, d + "/gpu") for _ in range(2))
t_gpu_host = run(ours, d + "/gpu_hostparse", ["--host_parse"])
print("colate_b200 CLI: %.2f s wall (GPU ingest); %.2f s with --host_parse" % (t_gpu, t_gpu_host), flush=True)
assert open(d + "/gpu.coal").read() == open(d + "/gpu_hostparse.coal").read()
for a in sys.argv:
    if a.startswith("--devices="):
        run(ours, d + "/multi_warm", ["--devices", a.split("=", 1)[1]])
        t_multi = min(run(ours, d + "/multi", ["--devices", a.split("=", 1)[1]]) for _ in range(2))
        same = open(d + "/gpu.coal").read() == open(d + "/multi.coal").read()
        os.environ["COLATE_TIMING"] = "1"; r = subprocess.run([ours] + common + ["--devices", a.split("=", 1)[1], "-o", d + "/tm"], capture_output=True, text=True); print("".join(l + "\n" for l in r.stderr.splitlines() if l.startswith("[timing]"))); del os.environ["COLATE_TIMING"]
        print("colate_b200 CLI --devices %s: %.2f s wall; .coal byte-identical to one device: %s" % (a.split("=", 1)[1], t_multi, same), flush=True)
if not skip_ref:
    ref = os.path.join(ROOT, "oracle", "_ref", "Colate")
    t_ref = run(ref, d + "/ref")
    same = open(d + "/gpu.coal").read() == open(d + "/ref.coal").read()
    print("reference CLI (1 core): %.1f s wall; .coal files byte-identical: %s; speed-up %.0fx" % (t_ref, same, t_ref / t_gpu))
