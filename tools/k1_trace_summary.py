#!/usr/bin/env python
"""Summary of a score-pass timeline written with RMGPU_K1_TRACE=<file> (four 64-bit words per claim: ticket * 16 + segment,
claim, start and end time in ns): how long sub-tickets take, how evenly the resident warps progress, and how long a
sub-ticket waited for its own previous segment.
usage: k1_trace_summary.py <trace.bin> [<resident warps, default 2368>]"""
import sys

import numpy as np

a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 4)
warps = int(sys.argv[2]) if len(sys.argv) > 2 else 2368
seg = (a[:, 0] & 15).astype(np.int64)
t0 = a[:, 1].min()
claim, start, end = ((a[:, i] - t0) / 1e6 for i in (1, 2, 3))
wait, dur = start - claim, end - start
nseg = int(seg.max()) + 1
print(f"{len(a)} sub-tickets ({nseg} segments per ticket), launch {end.max():.2f} ms, {warps} resident warps")
print(f"busy {dur.sum() / warps:.1f} ms per warp, waiting for the previous segment {wait.sum() / warps:.2f} ms per warp")
print("duration of a sub-ticket, ms: " + ", ".join(f"{p}% {v:.2f}" for p, v in zip((1, 5, 25, 50, 75, 95, 99, 100),
                                                                                     np.percentile(dur, (1, 5, 25, 50, 75, 95, 99, 100)))))
print(f"  the first {warps} claims (all started together): mean {dur[:warps].mean():.2f} ms; the next {warps}: {dur[warps:2 * warps].mean():.2f} ms")
for s in range(nseg):
    m = seg == s
    print(f"segment {s}: {m.sum()} sub-tickets, wait mean {wait[m].mean():.3f} ms, max {wait[m].max():.2f} ms, duration mean {dur[m].mean():.2f} ms")
print(f"sub-tickets that waited more than 0.5 ms: {(wait > 0.5).sum()}")
marks = [int(len(a) * f) for f in (0, 0.0625, 0.125, 0.25, 0.5, 0.75, 0.999)]
print("claim number -> claim time, ms: " + ", ".join(f"{u}: {claim[u]:.1f}" for u in marks))
