#!/usr/bin/env python
"""Diagnostic: concurrent MulPir queries/s of this process under different process set-ups (see tools/gpu_two_process_pir.sh): plain / after importing torch / with an NCCL group; optional thread count."""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "swift-homomorphic-encryption_b200")):
    sys.path.insert(0, p)
mode = sys.argv[1]
device = int(sys.argv[2]) if len(sys.argv) > 2 else 0
if mode in ("torch", "nccl"):
    import torch
    torch.cuda.set_device(device)
    torch.zeros(1, device="cuda")
if mode == "nccl":
    import torch.distributed as dist
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29533", rank=0, world_size=1)
    dist.barrier()
import numpy as np

import hecuda
from hecuda import pir

hecuda.set_device(device)
MOD = [134176769, 268369921, 268361729]
n, t = 4096, 17
ctx = hecuda.Context(n, MOD, t)
L = ctx.L
rng = np.random.default_rng(3)


def uniform(moduli, prefix):
    out = np.empty(tuple(prefix) + (len(moduli), n), dtype=np.uint64)
    for i, q in enumerate(moduli):
        out[..., i, :] = rng.integers(0, q, size=tuple(prefix) + (n,), dtype=np.uint64)
    return out


param = pir.MulPir.generateParameter(pir.IndexPirConfig(1 << 20, 64, 2, 1, True, "hybridCompression", False), ctx)
count = int(np.prod(param.dimensions))
db = pir.ProcessedDatabase(ctx, rng.integers(0, t, size=(count, n), dtype=np.uint64), None, evalFormat=False)
server = pir.MulPirServer(param, ctx, [db])
key = hecuda.EvaluationKey(ctx, uniform(MOD, (L, 2)))
for e in param.evaluationKeyConfig.galoisElements:
    key.setGaloisKey(e, uniform(MOD, (L, 2)))
query = uniform(MOD[:L], (1, 2))
for _ in range(3):
    server.computeResponse(query, key)
only = int(sys.argv[3]) if len(sys.argv) > 3 else 0
for threads in ((1, 2, 4, 8) if not only else (only, only, 4, only)):
    def worker():
        for _ in range(10):
            server.computeResponse(query, key)
    pool = [threading.Thread(target=worker) for _ in range(threads)]
    t0 = time.perf_counter()
    for th in pool:
        th.start()
    for th in pool:
        th.join()
    dt = time.perf_counter() - t0
    print(f"mode={mode} device={device} threads={threads}: {threads * 10 / dt:8.1f} queries/s", flush=True)
