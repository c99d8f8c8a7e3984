#!/usr/bin/env python
"""BASELINE config 5 on one GPU: PNNS encrypted-vector x plaintext-matrix products, N=8192, 512-dimensional vectors.

    python tools/bench_pnns.py [database_rows] [vector_dim] [batch]

Synthetic: uniform coefficient plaintexts < t for the diagonal-packed database, uniform residues for the query
ciphertexts and Galois keys.  Prints one JSON line: latency of one query vector, throughput of a batch that shares
one evaluation key (hecuda_pnns_mul_transpose_vector + modSwitchDownToSingle through the host-pointer ABI), and the
CPU restatement on a bounded sample.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "swift-homomorphic-encryption_b200")):
    sys.path.insert(0, p)
import numpy as np

import hecuda
from hecuda import pnns

Q8192 = [36028797018652673, 36028797017571329, 36028797017456641, 36028797017276417]


def uniform(rng, moduli, prefix, n):
    out = np.empty(tuple(prefix) + (len(moduli), n), dtype=np.uint64)
    for i, q in enumerate(moduli):
        out[..., i, :] = rng.integers(0, q, size=tuple(prefix) + (n,), dtype=np.uint64)
    return out


def run(rows=100000, dim=512, batch=16, reps=5, cpu=True):
    """One row block of the database per GPU (weak scaling: every rank holds `rows` rows and answers the same batch of
    query vectors; no data-path collective).  Returns the result dict on rank 0 (None elsewhere)."""
    n, t = 8192, 65537
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        hecuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    ctx = hecuda.Context(n, Q8192, t)
    L = ctx.L
    rng = np.random.default_rng(5 + rank)
    bsgs = pnns.BabyStepGiantStep.forVectorDimension(dim)
    results = -(-rows // n)
    count = bsgs.vectorDimension * results
    plain = rng.integers(0, t, size=(count, n), dtype=np.uint64)
    t0 = time.time()
    matrix = pnns.PlaintextMatrix(ctx, pnns.MatrixDimensions(rows, dim), None, bsgs, plaintexts=plain)
    upload_s = time.time() - t0
    key = hecuda.EvaluationKey(ctx, None)
    for e in {pnns.GaloisElement.rotatingColumns(-1, n), pnns.GaloisElement.rotatingColumns(-bsgs.babyStep, n)}:
        key.setGaloisKey(e, uniform(rng, Q8192, (L, 2), n))
    vec = hecuda.PinnedBuffer((batch, 2, L, n))
    vec.array[...] = uniform(rng, Q8192[:L], (batch, 2), n)

    def answer(b):
        return matrix.mulTranspose(vec.array[:b], key, modSwitchDownToSingle=True)

    for _ in range(2):
        answer(1), answer(batch)
    t0 = time.perf_counter()
    for _ in range(reps):
        answer(1)
    latency_ms = (time.perf_counter() - t0) / reps * 1e3
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        answer(batch)
    batch_ms = (time.perf_counter() - t0) / reps * 1e3
    if world > 1:
        worst = torch.tensor([batch_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        batch_ms = float(worst.item())
        dist.destroy_process_group()
        if rank != 0:
            return None
    db_bytes = count * L * n * 8
    out = {
        "metric": "PNNS mulTranspose query vectors/s (matrix resident in HBM)",
        "config": {"workload": f"N={n}, 4 x 55-bit moduli, t={t}, database {rows} x {dim}, babyStep={bsgs.babyStep}, "
                               f"giantStep={bsgs.giantStep}, result ciphertexts/vector={results}, batch={batch}"},
        "database_plaintexts": count, "database_gb": round(db_bytes / 1e9, 3), "database_upload_s": round(upload_s, 2),
        "latency_ms": round(latency_ms, 3), "batch_ms": round(batch_ms, 3), "value": round(batch / (batch_ms / 1e3), 1),
        "unit": "vectors/s", "dot_products_per_s": round(batch / (batch_ms / 1e3) * rows * world, 1), "n_gpus": world,
        "database_rows_per_gpu": rows, "reps": reps,
        "db_scan_gbs_at_value": round(batch / (batch_ms / 1e3) * db_bytes / 1e9, 1), "database_bytes": db_bytes,
        "h2d_bytes_per_batch": int(vec.array.nbytes), "d2h_bytes_per_batch": int(batch * results * 2 * n * 8),
        "gpu_launches": hecuda.kernel_launch_count(),
    }
    if cpu and os.environ.get("PNNS_CPU", "1") == "1" and world == 1:
        from oracle import oracle as orc
        o = orc.Context(n, Q8192, t)
        gk = uniform(rng, Q8192, (L, 2), n)
        ct = np.ascontiguousarray(vec.array[0])
        t0 = time.perf_counter()
        for _ in range(4):
            o.apply_galois(ct, pnns.GaloisElement.rotatingColumns(-1, n), gk, threads=1)
        rot_s = (time.perf_counter() - t0) / 4
        first = np.stack([ct] * bsgs.babyStep)
        cap = min(results * bsgs.giantStep, 8)
        pts = uniform(rng, Q8192[:L], (cap, bsgs.babyStep), n)
        t0 = time.perf_counter()
        o.inner_product_plain(first, pts, None)
        scan_s = (time.perf_counter() - t0) * (results * bsgs.giantStep / cap)
        rotations = (bsgs.babyStep - 1) + results * (bsgs.giantStep - 1)
        total = rotations * rot_s + scan_s
        out["cpu_baseline"] = {"value": round(1.0 / total, 4), "unit": "vectors/s", "cores": orc.num_threads(), "kind": "port",
                               "sample": f"{rotations} rotations x {rot_s * 1e3:.2f} ms (4 timed, one thread each) + inner products "
                                         f"{scan_s:.2f}s (timed on {cap} of {results * bsgs.giantStep} rows, scaled; NTTs of the "
                                         f"plaintexts excluded: database taken as already in Eval format)"}
    return out


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    dim = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    batch = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    out = run(rows, dim, batch)
    if out is not None:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
