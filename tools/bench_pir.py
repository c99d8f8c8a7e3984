#!/usr/bin/env python
"""BASELINE config 4 on one GPU: MulPir server computeResponse over a synthetic index-PIR database held in HBM.

    python tools/bench_pir.py [entry_count] [entry_size_bytes] [threads]

Encryption parameters = the reference's PIR benchmark default (EncryptionParametersConfig.defaultPir,
_BenchmarkUtilities/BenchmarkMetricExtensions.swift:60-63: N=4096, log t = 5, log q = 27/28/28), index-PIR config = the
benchmark's (PirBenchmarkUtilities.swift:147-158: 2 dimensions, uneven, hybrid key compression, batch 1).
Synthetic data: uniform residues for the query ciphertexts and the evaluation key (the server's arithmetic does not
depend on them being well-formed), uniform coefficients < t for the database.  Prints one JSON line: latency of one
query through the host-pointer C ABI, throughput with `threads` concurrent queries (one stream each), the size of
the resident database and the first-dimension scan rate; plus the CPU restatement (oracle) on a bounded sample.
"""
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "swift-homomorphic-encryption_b200")):
    sys.path.insert(0, p)
import numpy as np

import hecuda
from hecuda import pir

PIR_MODULI = [134176769, 268369921, 268361729]  # n_4096_logq_27_28_28_logt_5 (EncryptionParameters.swift:357-367)


def uniform(rng, moduli, shape_prefix, n):
    out = np.empty(tuple(shape_prefix) + (len(moduli), n), dtype=np.uint64)
    for i, q in enumerate(moduli):
        out[..., i, :] = rng.integers(0, q, size=tuple(shape_prefix) + (n,), dtype=np.uint64)
    return out


def run(entries=1 << 20, entry_size=64, threads=8, per_thread=30, cpu=True):
    """Returns the result dict on rank 0 (None on the other ranks)."""
    n, t = 4096, 17
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:  # one database shard per GPU (KeywordDatabase shards are independent): weak scaling, no data-path collective
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        hecuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    ctx = hecuda.Context(n, PIR_MODULI, t)
    L = ctx.L
    rng = np.random.default_rng(3 + rank)
    config = pir.IndexPirConfig(entries, entry_size, 2, 1, True, "hybridCompression", False)
    param = pir.MulPir.generateParameter(config, ctx)
    chunk_count = -(-param.encodedEntrySize // pir.bytesPerPlaintext(ctx))
    count = chunk_count * int(np.prod(param.dimensions))
    t0 = time.time()
    rows = rng.integers(0, t, size=(count, n), dtype=np.uint64)
    db = pir.ProcessedDatabase(ctx, rows, None, evalFormat=False)
    process_s = time.time() - t0
    del rows
    server = pir.MulPirServer(param, ctx, [db])
    elements = param.evaluationKeyConfig.galoisElements
    if world > 1:  # the client's evaluation key reaches every shard by one NCCL broadcast
        from hecuda.distributed import broadcast_evaluation_key
        relin = uniform(rng, PIR_MODULI, (L, 2), n) if rank == 0 else None
        gal = {e: uniform(rng, PIR_MODULI, (L, 2), n) for e in elements} if rank == 0 else None
        key = broadcast_evaluation_key(ctx, relin, 0, gal, elements)
    else:
        key = hecuda.EvaluationKey(ctx, uniform(rng, PIR_MODULI, (L, 2), n))
        for e in elements:
            key.setGaloisKey(e, uniform(rng, PIR_MODULI, (L, 2), n))
    query_cts = -(-param.expandedQueryCount // n)
    query = uniform(rng, PIR_MODULI[:L], (query_cts, 2), n)

    def one():
        return server.computeResponse(query, key)

    for _ in range(3):
        one()
    reps = 10
    t0 = time.perf_counter()
    for _ in range(reps):
        one()
    latency_ms = (time.perf_counter() - t0) / reps * 1e3

    # the same query as bytes: seeded serialized ciphertexts in, skipLSBs-packed reply out (hecuda_mulpir_compute_response_wire)
    packed = hecuda.Bfv.serialize(ctx, query[:, 0])
    seeds = rng.integers(0, 256, size=(query_cts, 32), dtype=np.uint8)
    for _ in range(3):
        pir.PirWire.computeResponse(server, packed, seeds, key)
    t0 = time.perf_counter()
    for _ in range(reps):
        wire_reply, wire_skips = pir.PirWire.computeResponse(server, packed, seeds, key)
    wire_latency_ms = (time.perf_counter() - t0) / reps * 1e3

    def worker(count=per_thread):
        for _ in range(count):
            one()
    # warm the per-thread streams and the stream-ordered memory pool at this concurrency (first use of a stream grows
    # the pool, which synchronises the device) -- a server keeps them warm
    pool = [threading.Thread(target=worker, args=(5,)) for _ in range(threads)]
    for th in pool:
        th.start()
    for th in pool:
        th.join()
    pool = [threading.Thread(target=worker) for _ in range(threads)]
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for th in pool:
        th.start()
    for th in pool:
        th.join()
    concurrent_s = time.perf_counter() - t0
    if world > 1:
        worst = torch.tensor([concurrent_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        concurrent_s = float(worst.item())
    qps = world * threads * per_thread / concurrent_s
    if rank != 0:
        dist.destroy_process_group()
        return None

    import ctypes as C
    ptr, nbytes = C.c_void_p(), C.c_uint64(0)
    hecuda.load_library().hecuda_pir_database_device_buffer(db._h, C.byref(ptr), C.byref(nbytes))
    db_bytes = int(nbytes.value) or count * L * n * 8  # uint32 rows for the small default moduli
    out = {
        "metric": "MulPir computeResponse queries/s (index PIR, database resident in HBM)",
        "config": {"workload": f"N={n}, q=27/28/28 bit, t={t}, entries={entries} x {entry_size} B, dims={param.dimensions}, "
                               f"chunks={chunk_count}, galois={param.evaluationKeyConfig.galoisElements}"},
        "database_plaintexts": count, "database_gb": round(db_bytes / 1e9, 3), "database_upload_s": round(process_s, 2),
        "latency_ms": round(latency_ms, 3),
        "wire": {"latency_ms": round(wire_latency_ms, 3), "request_bytes": int(packed.nbytes + seeds.nbytes),
                 "reply_bytes": int(wire_reply.nbytes), "skip_lsbs": wire_skips,
                 "unpacked_request_bytes": int(query.nbytes), "unpacked_reply_bytes": int(2 * n * 8 * chunk_count)},
        "threads": threads, "queries_per_thread": per_thread, "concurrent_s": concurrent_s, "n_gpus": world, "scaling": "weak (one shard per GPU)",
        "value": round(qps, 1), "unit": "queries/s",
        "db_scan_gbs_at_value": round(qps * db_bytes / 1e9, 1),
        "gpu_launches": hecuda.kernel_launch_count(),
    }
    if world > 1:
        dist.destroy_process_group()
    if cpu and os.environ.get("PIR_CPU", "1") == "1" and world == 1:
        from oracle import oracle as orc
        from oracle import pir_oracle as opir
        o = orc.Context(n, PIR_MODULI, t)
        oparam = opir.IndexPirParameter(entries, entry_size, list(param.dimensions), 1,
                                        list(param.evaluationKeyConfig.galoisElements), False)
        # bounded CPU sample: the same query against a database cut to the first `cap` columns of the first chunk,
        # scaled back linearly in the database size (the expansion, which does not scale, is timed separately)
        okeys = {e: uniform(rng, PIR_MODULI, (L, 2), n) for e in oparam.galois_elements}
        relin = uniform(rng, PIR_MODULI, (L, 2), n)
        qlist = [query[i] for i in range(query_cts)]
        t0 = time.perf_counter()
        expanded = opir.expand(o, qlist, oparam.expanded_query_count, okeys)
        expand_s = time.perf_counter() - t0
        dim0 = param.dimensions[0]
        first = np.stack([np.stack([orc.ntt_forward(n, o.q, ct[p]) for p in range(2)]) for ct in expanded[:dim0]])
        cap = min(int(np.prod(param.dimensions[1:])) * chunk_count, 64)
        pts = uniform(rng, PIR_MODULI[:L], (cap, dim0), n)
        t0 = time.perf_counter()
        o.inner_product_plain(first, pts, None)
        scan_s = (time.perf_counter() - t0) * (count / (cap * dim0))
        pairs = param.dimensions[1]
        lhs = np.stack(expanded[dim0:dim0 + pairs])[None]
        t0 = time.perf_counter()
        prod = o.inner_product(lhs, lhs)
        o.relinearize(prod[0], relin)
        tail_s = (time.perf_counter() - t0) * chunk_count
        total = expand_s + scan_s + tail_s
        out["cpu_baseline"] = {"value": round(1.0 / total, 4), "unit": "queries/s", "cores": orc.num_threads(), "kind": "port",
                               "sample": f"expand {expand_s:.2f}s + first-dimension scan {scan_s:.2f}s (timed on {cap * dim0} "
                                         f"of {count} plaintexts, scaled) + ct x ct / relinearize {tail_s:.2f}s"}
    return out


def main():
    entries = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    entry_size = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    threads = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    out = run(entries, entry_size, threads)
    if out is not None:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
