"""Throughput of the GPU-side .mut ingest (N1) against the host reader, on generated whole-genome text."""
import os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from colate_b200 import api, synth

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
copies = int(sys.argv[2]) if len(sys.argv) > 2 else 10
sites = synth.make_sites(1, [rows], [2.49e8])
d = tempfile.mkdtemp()
p = os.path.join(d, "x.mut")
synth.write_mut(p, sites, 0)
one = open(p, "rb").read()
hdr_end = one.index(b"\n") + 1
text = one[:hdr_end] + one[hdr_end:] * copies            # `copies` x the data lines: a whole-genome sized file
t0 = time.perf_counter(); want = api.read_mut(p); t_host = time.perf_counter() - t0
print("host reader (colate_read_mut, 1 core): %d rows, %.1f MB in %.3f s = %.1f MB/s" % (len(want[0]), len(one) / 1e6, t_host, len(one) / 1e6 / t_host))
import torch
pinned = torch.empty(len(text), dtype=torch.uint8).pin_memory()
pinned.numpy()[:] = np.frombuffer(text, dtype=np.uint8)
h = api.Handle(0)
import ctypes as C
for it in range(3):
    api.check(api.lib().colate_ingest_begin(h._h, 1, rows * copies + 8))
    t0 = time.perf_counter()
    n = api.check(api.lib().colate_ingest_mut_text(h._h, C.c_void_p(pinned.data_ptr()), len(text), 0))
    dt = time.perf_counter() - t0
    api.check(api.lib().colate_ingest_end(h._h))
    st = h.ingest_stats()
    print("GPU ingest: %d rows, %.1f MB: wall %.1f ms (H2D from pinned + kernels) = %.2f GB/s; kernels %.2f ms = %.1f GB/s of text; host-parsed rows %d"
          % (n, len(text) / 1e6, dt * 1e3, len(text) / 1e9 / dt, st["kernel_ms"], len(text) / 1e6 / st["kernel_ms"], st["host_fallback_rows"]))
h.n_site = n
got = h.ingest_fetch(0, len(want[0]))
print("first copy identical to the host reader:", all(np.array_equal(g.view(np.uint32), w.view(np.uint32)) for g, w in zip(got, want)))
