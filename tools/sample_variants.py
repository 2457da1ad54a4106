#!/usr/bin/env python3
"""k_sample tuning: build kernels_sites.cu with different (warps per CTA, ring slots, words per chunk) into
build/variants/ (here, no GPU needed) and time each on the GPU box at config-2 size, with a checksum of the
stage-i result so that every variant is seen to be bit-identical.

  python tools/sample_variants.py build 4,2,20 4,3,20 4,2,40 ...
  python tools/sample_variants.py run [rows]          (on the GPU box)
"""
import glob, hashlib, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "colate_b200", "csrc")
OUT = os.path.join(ROOT, "build", "variants")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def build(tags):
    os.makedirs(OUT, exist_ok=True)
    subprocess.run(["make", "-s", "-j8", "-C", CSRC, "all"], check=True)
    others = [os.path.join(CSRC, o) for o in ("abi.o", "kernels_em.o", "kernels_ingest.o", "host_mt.o", "host_misc.o")]
    for t in tags:
        w, r, c, *extra = t.split(",")
        tag = f"{w}_{r}_{c}" + "".join("_" + e for e in extra)
        objs = []
        for src in ("kernels_sites", "kernels_mt"):       # both see the chunk width (internal.h: stream layout)
            objs.append(os.path.join(OUT, f"{src}_{tag}.o"))
            subprocess.run(["nvcc", *ARCH, "-O3", "-lineinfo", "-std=c++17", "--fmad=false", "-Xcompiler", "-fPIC,-O2",
                            f"-DS2_WARPS_={w}", f"-DS2_RING_={r}", f"-DS2_CH_WORDS_={c}", *[f"-D{e}" for e in extra], "-c",
                            os.path.join(CSRC, src + ".cu"), "-o", objs[-1]], check=True)
        subprocess.run(["nvcc", *ARCH, "-shared", "-o", os.path.join(OUT, f"lib_{tag}.so"), *objs, *others, "-lz"], check=True)
        print("built", t, flush=True)


def run_one(rows):
    sys.path.insert(0, ROOT)
    import numpy as np
    from colate_b200 import api, synth
    sites = synth.make_sites(1, synth.rows_for_genome(rows), synth.AUTOSOME_LEN)
    gt = synth.make_genome(101, sites, 0.7)
    gr = synth.make_genome(201, sites, 0.7)
    h = api.Handle(0)
    h.load(sites, gt, gr)
    ts = []
    for it in range(6):
        s1 = h.stage1(api.mt_seed(1))
        ts.append(h.stage1_timing()["sample_ms"])
    dig = hashlib.sha256(s1.block_stats.tobytes() + s1.block_tallies.tobytes() + s1.mt_state.tobytes()).hexdigest()[:16]
    print("VARIANT %s sample_ms min %.4f median %.4f n_used %d sha %s" % (
        os.environ.get("COLATE_B200_LIB", "in-tree").split("/")[-1], min(ts[1:]), sorted(ts[1:])[2], s1.n_used, dig), flush=True)


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
