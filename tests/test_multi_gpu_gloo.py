"""world_size-2 gloo tests of the multi-GPU plumbing (hecuda/distributed.py) on CPU: batch sharding, key broadcast and
result gathering.  The per-shard engine call is injected; here it is the oracle (the CUDA engine cannot run without a
GPU), so this covers exactly the host-side logic that bench.py --gpus N and the N>1 deployment use."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    from hecuda.distributed import shard_range

    for batch in (0, 1, 2, 7, 1024, 4097):
        for world in (1, 2, 3, 8):
            spans = [shard_range(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "swift-homomorphic-encryption_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hecuda import distributed as hd
    from oracle import oracle as orc

    n = 64
    moduli = orc.generate_primes([50, 50, 50], False, n)
    t = orc.generate_primes([12], True, 1)[0]
    ctx = orc.Context(n, moduli, t)  # per-rank replica, built deterministically: no communication
    L = ctx.L
    batch = 7  # not divisible by the world size
    a = orc.fill_uniform(1, ctx.q, n, batch * 2 * L).reshape(batch, 2, L, n)
    b = orc.fill_uniform(2, ctx.q, n, batch * 2 * L).reshape(batch, 2, L, n)
    # key material exists on rank 0 only and is broadcast once at setup
    shape = (L, 2, L + 1, n)
    rk = ctx.keygen(5)[1] if rank == 0 else None
    rk = hd.broadcast_key_bytes(rk, shape, src=0).reshape(shape)
    prod = hd.sharded_apply(lambda x, y: ctx.mul(x, y, threads=1), a, b)
    relin = hd.sharded_apply(lambda x: ctx.relinearize(x, rk, threads=1), prod)
    # PIR: one database shard per rank, the client's Galois keys broadcast once, every query answered on its shard
    from oracle import pir_oracle as opir
    entries, entry_size = 90, 3
    lo, hi = hd.shard_databases(entries, rank, world)
    pctx = orc.Context(16, orc.generate_primes([55, 52, 62, 58], False, 16), 1153)
    oparam = opir.generate_parameter(opir.IndexPirConfig(hi - lo, entry_size, 2, 1, False, "noCompression", False), 16, 1153)
    database = [bytes([(7 * i + k) % 256 for k in range(entry_size)]) for i in range(entries)]
    sk, prk = pctx.keygen(9)
    gshape = (pctx.L, 2, pctx.L + 1, 16)
    gkeys = {e: pctx.galois_keygen(40 + e, sk, e) for e in oparam.galois_elements} if rank == 0 else None
    gkeys = {e: k.reshape(gshape) for e, k in hd.broadcast_galois_keys_bytes(gkeys, oparam.galois_elements, gshape, 0).items()}
    prk = hd.broadcast_key_bytes(prk if rank == 0 else None, gshape, src=0).reshape(gshape)
    shard = opir.process_database(pctx, oparam, database[lo:hi])
    answers = {}
    for index in (3, 44, 45, 89):
        if lo <= index < hi:
            query = opir.generate_query(pctx, oparam, [index - lo], sk, 300 + index)
            reply = opir.compute_response(pctx, query, 1, gkeys, prk, [shard], oparam)
            answers[index] = opir.decrypt_response(pctx, oparam, reply, [index - lo], sk)[0]
    np.save(os.path.join(out_dir, f"pir_{rank}.npy"), np.array([[i] + list(v) for i, v in answers.items()], dtype=np.int64))
    np.save(os.path.join(out_dir, f"relin_{rank}.npy"), relin)
    np.save(os.path.join(out_dir, f"key_{rank}.npy"), rk)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_pipeline(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from oracle import oracle as orc

    n = 64
    moduli = orc.generate_primes([50, 50, 50], False, n)
    t = orc.generate_primes([12], True, 1)[0]
    ctx = orc.Context(n, moduli, t)
    L = ctx.L
    a = orc.fill_uniform(1, ctx.q, n, 7 * 2 * L).reshape(7, 2, L, n)
    b = orc.fill_uniform(2, ctx.q, n, 7 * 2 * L).reshape(7, 2, L, n)
    rk = ctx.keygen(5)[1]
    expect = ctx.relinearize(ctx.mul(a, b), rk)
    served = {}
    for r in range(world):
        for row in np.load(tmp_path / f"pir_{r}.npy"):
            served[int(row[0])] = bytes(int(v) for v in row[1:])
    assert served == {i: bytes([(7 * i + k) % 256 for k in range(3)]) for i in (3, 44, 45, 89)}
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"key_{r}.npy"), rk)
        assert np.array_equal(np.load(tmp_path / f"relin_{r}.npy"), expect)
