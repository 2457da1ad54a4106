#!/usr/bin/env python
"""bench.py -- pair-site evaluations/s of the `Colate --mode mut` hot path on B200.

One step = one full pass of the path over one synthetic whole-genome sample pair
(BASELINE.json configs[1]: 22 autosomes, 10 M .mut rows, two ~1x .colate.in genomes,
--bins 3,7,0.1, one replicate): stage i (record->row join, row filter + stream look-ahead,
std::mt19937 stream by jump-ahead, 100 Monte-Carlo age draws per used row, per-block
histograms) -> stage ii (block bootstrap + F redistribution) -> stage iii (EM to convergence).

  value  device-resident: SoA inputs already in HBM when the timed region starts
  e2e    the same pass through the C-ABI with HOST buffers: pinned host -> device copies of
         every input and the device -> host read of the rates inside the timed region

N > 1 (torchrun, one process per GPU): the path shards by sample pair (configs[4]); every rank
runs its own pair of the same shape, no data-path collective, weak scaling; the max over ranks
of the device time is taken with one NCCL all-reduce outside the timed steps.

--impl reference times the unmodified reference CLI (oracle/_ref/Colate, compiled from
/root/reference by oracle/Makefile) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "pair_site_evals_per_s"
UNIT = "pair-site evaluations/s"
BINS = "3,7,0.1"
SEED = 1


def workload_name(rows):
    return (f"synthetic whole-genome single pair: 22 autosomes, {rows} .mut rows, two ~1x .colate.in genomes "
            f"(record w.p. 0.7, N=1+floor(Exp(1)), derived w.p. 0.3), --bins {BINS}, --seed {SEED}, 1 replicate "
            f"(BASELINE.json configs[1]; per rank one such pair when --gpus > 1)")


# --------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------
def chr1_sample_files(rows_total, workdir):
    """Chromosome 1 of the whole-genome dataset as reference-format files (the bounded CPU sample)."""
    from colate_b200 import synth
    rows = synth.rows_for_genome(rows_total)[0]
    sites = synth.make_sites(SEED, [rows], [synth.AUTOSOME_LEN[0]], chr_names=["1"])
    gt = synth.make_genome(SEED + 100, sites, 0.7)
    gr = synth.make_genome(SEED + 200, sites, 0.7)
    os.makedirs(workdir, exist_ok=True)
    with open(os.path.join(workdir, "chr.txt"), "w") as f:
        f.write("1\n")
    synth.write_mut(os.path.join(workdir, "syn_chr1.mut"), sites, 0)
    synth.write_colate_in_fast(os.path.join(workdir, "t.colate.in"), gt, sites.chr_names)
    synth.write_colate_in_fast(os.path.join(workdir, "r.colate.in"), gr, sites.chr_names)
    return rows


def run_reference_cli(workdir, out_prefix):
    from oracle import pyoracle as po
    cli = po.ref_cli()
    cmd = [cli, "--mode", "mut", "--mut", os.path.join(workdir, "syn"), "--chr", os.path.join(workdir, "chr.txt"),
           "--target_tmp", os.path.join(workdir, "t.colate.in"), "--reference_tmp", os.path.join(workdir, "r.colate.in"),
           "--bins", BINS, "--seed", str(SEED), "-o", out_prefix]
    return subprocess.Popen(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def cpu_baseline_one_core(rows_total):
    """The unmodified reference CLI, one process (it is single-threaded), on chr1 of the dataset."""
    from oracle import pyoracle as po
    if po.ref_cli() is None:
        return cpu_baseline_port(rows_total)
    d = tempfile.mkdtemp(prefix="colate_cpu_")
    try:
        rows = chr1_sample_files(rows_total, d)
        t0 = time.perf_counter()
        p = run_reference_cli(d, os.path.join(d, "ref"))
        p.wait()
        dt = time.perf_counter() - t0
        return {"value": rows / dt, "unit": UNIT, "cores": 1, "kind": "reference",
                "sample": f"oracle/_ref/Colate --mode mut on chromosome 1 of the dataset ({rows} rows; .mut text parse, "
                          f"site loop and EM included), {dt:.2f} s wall"}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def cpu_baseline_port(rows_total):
    """Fallback when oracle/_ref is absent: the C restatement on in-memory arrays (no text parse)."""
    from colate_b200 import synth
    from oracle import pyoracle as po
    rows = synth.rows_for_genome(rows_total)[0]
    sites = synth.make_sites(SEED, [rows], [synth.AUTOSOME_LEN[0]], chr_names=["1"])
    gt = synth.make_genome(SEED + 100, sites, 0.7)
    gr = synth.make_genome(SEED + 200, sites, 0.7)
    t0 = time.perf_counter()
    o = po.stage1(sites, gt, gr, seed=SEED)
    counts = po.stage2(np.ones((1, o["num_blocks"]), np.int32), o, 0.0)
    ep, _ = po.epochs_from_bins(BINS)
    po.em_run(ep, np.full(len(ep), 1 / 20000.0), counts[0])
    dt = time.perf_counter() - t0
    return {"value": rows / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"oracle/liboracle.so stage i-iii on chromosome 1 of the dataset ({rows} rows, arrays in memory), {dt:.2f} s"}


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation on all host cores: one single-threaded
    reference process per core, each on chromosome 1 of the dataset (a bounded sample of the workload)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import pyoracle as po
    cores = os.cpu_count() or 1
    d = tempfile.mkdtemp(prefix="colate_ref_")
    try:
        if po.ref_cli() is None:
            base = cpu_baseline_port(args.rows)
            line = {"metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                    "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                    "dtype": "f64", "data": "synthetic", "impl": "reference", "config": {"workload": workload_name(args.rows)},
                    "cpu_baseline": base, "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
            print(json.dumps(line))
            return 0
        rows = chr1_sample_files(args.rows, d)
        P = cores

        def step(i):
            procs = [run_reference_cli(d, os.path.join(d, f"out{i}_{k}")) for k in range(P)]
            for p in procs:
                p.wait()

        for i in range(args.warmup):
            step(-1 - i)
        t0 = time.perf_counter()
        for i in range(args.steps):
            step(i)
        dt = (time.perf_counter() - t0) / max(1, args.steps)
        val = P * rows / dt
        base = {"value": val, "unit": UNIT, "cores": P, "kind": "reference",
                "sample": f"{P} concurrent single-threaded oracle/_ref/Colate --mode mut processes per step, each on chromosome 1 "
                          f"of the dataset ({rows} rows: .mut text parse + site loop + EM)"}
        line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "impl": "reference", "config": {"workload": workload_name(args.rows)},
                "cpu_baseline": base, "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0
    finally:
        shutil.rmtree(d, ignore_errors=True)


# --------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000, help="rows of the whole-genome .mut (default: configs[1])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    from colate_b200 import api, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: colate_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # ---- synthetic pair of this rank (pair index = rank)
    sites = synth.make_sites(SEED + 1000 * rank, synth.rows_for_genome(args.rows), synth.AUTOSOME_LEN)
    gt = synth.make_genome(SEED + 100 + 1000 * rank, sites, 0.7)
    gr = synth.make_genome(SEED + 200 + 1000 * rank, sites, 0.7)
    meta = sites.meta()

    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()

    keep = []
    host = {}
    for name, arr in (("pos", sites.pos), ("ab", sites.age_begin), ("ae", sites.age_end), ("meta", meta)):
        t, v = pinned(arr); keep.append(t); host[name] = v
    genomes = []
    for g in (gt, gr):
        al = g.anc.astype(np.uint16) | (g.der.astype(np.uint16) << 8)
        first, end = api.chr_ranges(len(sites.chr_names), g.chrom)
        d = {}
        for name, arr in (("bp", g.bp), ("aaf", g.aaf), ("daf", g.daf), ("al", al)):
            t, v = pinned(arr); keep.append(t); d[name] = v
        d["first"], d["end"] = first, end
        genomes.append(d)
    h2d_bytes = sum(v.nbytes for v in host.values()) + sum(sum(d[k].nbytes for k in ("bp", "aaf", "daf", "al", "first", "end")) for d in genomes)

    h = api.Handle(local)
    h.set_option("rejoin", 1)  # the record->row join is part of every timed pass
    site_off = np.ascontiguousarray(sites.site_off, dtype=np.int64)
    ep, _ = api.epochs_from_bins(BINS)
    E = len(ep)
    rates_init = np.full(E, 1.0 / 20000.0)

    def upload():
        h.set_sites(site_off, host["pos"], host["ab"], host["ae"], host["meta"])
        for slot, d in enumerate(genomes):
            api.check(api.lib().colate_set_genome(h._h, slot, d["bp"].shape[0], api.ptr(d["first"]), api.ptr(d["end"]),
                                                  api.ptr(d["bp"]), api.ptr(d["aaf"]), api.ptr(d["daf"]), api.ptr(d["al"]), 0))

    stage1_acc = {}

    def one_pass():
        s1 = h.stage1(api.mt_seed(SEED))
        for k, v in h.stage1_timing().items():
            stage1_acc.setdefault(k, []).append(v)
        t0 = time.perf_counter()
        w = api.draw_block_weights(s1.mt_state, 1, s1.num_blocks)
        h.stage2_bootstrap(w, s1.block_stats, 0.0, fetch=False)
        t1 = time.perf_counter()
        rates, iters, ll = h.stage3_em(1, ep, rates_init)
        t2 = time.perf_counter()
        stage1_acc.setdefault("bootstrap_wall_ms", []).append((t1 - t0) * 1e3)
        stage1_acc.setdefault("em_wall_ms", []).append((t2 - t1) * 1e3)
        return s1, rates, iters

    ext = torch.cuda.ExternalStream(h.stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ext):
            e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        with torch.cuda.stream(ext):
            e1.record()
        torch.cuda.synchronize()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    upload()
    for _ in range(args.warmup):
        one_pass()
    stage1_acc.clear()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = h.launch_count()
    ms_res, (s1, rates, iters) = timed(one_pass, args.steps)
    launches = h.launch_count() - launches0
    em_probe = []
    t_stage1 = {k: float(np.mean(v)) for k, v in stage1_acc.items() if k.endswith("_ms")}

    def e2e_pass():
        upload()
        return one_pass()

    for _ in range(2):
        e2e_pass()
    ms_e2e, _ = timed(e2e_pass, args.steps)
    clocks = sampler.stop()

    rows = sites.n
    n_used = int(s1.n_used)
    per_step = ms_res / args.steps
    value = world * rows / (per_step * 1e-3)
    e2e_val = world * rows / (ms_e2e / args.steps * 1e-3)
    d2h_bytes = int(rates.nbytes + iters.nbytes + 8 + s1.block_stats.nbytes + s1.block_tallies.nbytes + 624 * 4)

    # ---- roofline of the per-mutation kernel (k_sample), measured live with CUDA events on the
    # handle's stream inside the library (colate_last_stage1_timing), averaged over the timed steps
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    sample_bytes = 800.0 * n_used
    # DRAM traffic of one k_sample launch from the ncu --set full capture of this workload
    # (profiles/r01_ncu_k_sample_full.md: dram__bytes_read.sum + dram__bytes_write.sum = 1.0067 GB + 0.2101 GB at
    # 1,209,943 used rows); it scales with the used rows (832 B read + 192 B written each)
    traffic = 1.21688e9 * n_used / 1209943.0
    ach = sample_bytes / (t_stage1["sample_ms"] * 1e-3) / 1e9
    stage_bytes = 40.0 * rows + 800.0 * n_used
    roofline = {"bound": "hbm", "kernel": "k_sample (per-mutation Monte-Carlo age binning)", "achieved": ach, "peak": peak,
                "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "traffic_source": "ncu --set full, profiles/r01_ncu_k_sample_full.md", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": sample_bytes, "launch_ms": t_stage1["sample_ms"],
                "stage1_all_kernels": {"bytes": stage_bytes, "ms": t_stage1["total_ms"],
                                       "gbs": stage_bytes / (t_stage1["total_ms"] * 1e-3) / 1e9,
                                       "frac": stage_bytes / (t_stage1["total_ms"] * 1e-3) / 1e9 / peak,
                                       "bytes_site_only": 40.0 * rows}}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": workload_name(args.rows), "rows_per_pair": rows, "used_rows_per_pair": n_used,
                           "num_blocks": int(s1.num_blocks), "epochs": E, "em_iterations": int(iters[0]),
                           "sharding": "one sample pair per GPU, no data-path collective" if world > 1 else "single GPU",
                           "l2": "inputs (~0.5 GB SoA + 0.2-1 GB generator stream per pass) exceed the 126 MB L2; no flush needed"},
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": d2h_bytes,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
                "stage_ms": {**t_stage1, "pass_total_ms": per_step}}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_one_core(args.rows)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    h.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
