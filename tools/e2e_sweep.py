#!/usr/bin/env python
"""End-to-end ct x ct multiply rate through the host-pointer ABI for one pipeline setting (read from the environment:
HECUDA_PIPELINE_DEPTH, HECUDA_PIPELINE_STAGES).  Usage: python tools/e2e_sweep.py [batch]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "swift-homomorphic-encryption_b200")):
    sys.path.insert(0, p)
import numpy as np

import hecuda

Q8192 = [36028797018652673, 36028797017571329, 36028797017456641, 36028797017276417]
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n, t = 8192, 557057
ctx = hecuda.Context(n, Q8192, t)
L = ctx.L
rng = np.random.default_rng(1)
hl, hr, ho = hecuda.PinnedBuffer((batch, 2, L, n)), hecuda.PinnedBuffer((batch, 2, L, n)), hecuda.PinnedBuffer((batch, 3, L, n))
for buf in (hl, hr):
    for i, q in enumerate(Q8192[:L]):
        buf.array[:, :, i, :] = rng.integers(0, q, size=(batch, 2, n), dtype=np.uint64)
for _ in range(2):
    hecuda.Bfv.mulAssign(ctx, hl.array, hr.array, out=ho.array)
steps = 5
t0 = time.perf_counter()
for _ in range(steps):
    hecuda.Bfv.mulAssign(ctx, hl.array, hr.array, out=ho.array)
dt = (time.perf_counter() - t0) / steps
gb = (hl.array.nbytes + hr.array.nbytes + ho.array.nbytes) / 1e9
print(f"depth={os.environ.get('HECUDA_PIPELINE_DEPTH', 'default')} stages={os.environ.get('HECUDA_PIPELINE_STAGES', 'default')} "
      f"batch={batch}: {batch / dt:9.0f} mult/s  {dt * 1e3:7.2f} ms/step  {gb / dt:6.1f} GB/s over PCIe (both directions)")
