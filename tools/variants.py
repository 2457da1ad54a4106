#!/usr/bin/env python3
"""Stage-i tuning variants: build kernels_sites.cu / kernels_mt.cu with extra -D flags into build/variants/ (here, no GPU
needed) and time every phase of stage i on the GPU box at config-2 size, with a checksum of the result so that every
variant is seen to be bit-identical.

  python tools/variants.py build name1:-DRP_Q_=1 name2:-DRP_Q_=2,-DRP_ROWS_=16 ...
  python tools/variants.py run [rows]          (on the GPU box; the in-tree library first)
"""
import glob, hashlib, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "colate_b200", "csrc")
OUT = os.path.join(ROOT, "build", "variants")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def build(specs):
    shutil.rmtree(OUT, ignore_errors=True)
    os.makedirs(OUT, exist_ok=True)
    subprocess.run(["make", "-s", "-j8", "-C", CSRC, "all"], check=True)
    em = any("EMC_" in sp or "EM_" in sp for sp in specs)          # EM variants rebuild kernels_em.cu too
    others = [os.path.join(CSRC, o) for o in ("abi.o", "kernels_ingest.o", "host_mt.o", "host_misc.o") + (() if em else ("kernels_em.o",))]
    for spec in specs:
        name, _, flags = spec.partition(":")
        defs = [f for f in flags.split(",") if f]
        objs = []
        for src in ("kernels_sites", "kernels_mt") + (("kernels_em",) if em else ()):
            objs.append(os.path.join(OUT, f"{src}_{name}.o"))
            subprocess.run(["nvcc", *ARCH, "-O3", "-lineinfo", "-std=c++17", "--fmad=false", "-Xcompiler", "-fPIC,-O2", *defs, "-c",
                            os.path.join(CSRC, src + ".cu"), "-o", objs[-1]], check=True)
        subprocess.run(["nvcc", *ARCH, "-shared", "-o", os.path.join(OUT, f"lib_{name}.so"), *objs, *others, "-lz"], check=True)
        print("built", spec, flush=True)


def run_one(rows):
    sys.path.insert(0, ROOT)
    import numpy as np
    from colate_b200 import api, synth
    sites = synth.make_sites(1, synth.rows_for_genome(rows), synth.AUTOSOME_LEN)
    gt = synth.make_genome(101, sites, 0.7)
    gr = synth.make_genome(201, sites, 0.7)
    h = api.Handle(0)
    h.load(sites, gt, gr)
    h.set_option("rejoin", 1)
    ts = []
    for it in range(7):
        s1 = h.stage1(api.mt_seed(1))
        ts.append(h.stage1_timing())
    keys = [k for k in ts[0] if k.endswith("_ms")]
    best = {k: min(t[k] for t in ts[2:]) for k in keys}
    dig = hashlib.sha256(s1.block_stats.tobytes() + s1.block_tallies.tobytes() + s1.mt_state.tobytes()).hexdigest()[:16]
    print("VARIANT %-28s %s n_used %d sha %s" % (os.environ.get("COLATE_B200_LIB", "in-tree").split("/")[-1],
          " ".join("%s %.4f" % (k[:-3], best[k]) for k in keys), s1.n_used, dig), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    elif sys.argv[1] == "one":
        run_one(int(sys.argv[2]))
    else:
        rows = sys.argv[2] if len(sys.argv) > 2 else "10000000"
        for lib in [""] + sorted(glob.glob(os.path.join(OUT, "lib_*.so"))):
            env = dict(os.environ)
            if lib:
                env["COLATE_B200_LIB"] = lib
            subprocess.run([sys.executable, os.path.abspath(__file__), "one", rows], env=env)
