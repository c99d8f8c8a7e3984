"""Multi-GPU plumbing for the batched hot path.

The unit of sharding is the ciphertext: every BEHZ / key-switch step couples all RNS rows of a coefficient, so
limb-sharding would need an all-to-all per op, while ciphertexts are independent.  One process per GPU
(torch.distributed); each rank owns the contiguous slice [rank*B/W, (rank+1)*B/W) of a batch and a private replica
of the immutable device context (built deterministically from (N, moduli, t) -- no communication).  The only
collective is the one-time broadcast of the evaluation key from rank 0 (NCCL over NVLink on GPUs, gloo in the CPU
tests); nothing is exchanged on the per-ciphertext path.

The reference has no distributed layer; this mirrors how its callers shard work offline
(KeywordPir shards, Sources/PrivateInformationRetrieval/KeywordPir/KeywordDatabase.swift:56-110).
"""
from __future__ import annotations

from typing import Callable, Tuple

import numpy as np


def shard_range(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of `batch` items over `world` ranks (first `batch % world` ranks get one more)."""
    if world < 1 or not (0 <= rank < world) or batch < 0:
        raise ValueError("invalid shard request")
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class _CudaView:
    """Zero-copy torch view of a raw device buffer via __cuda_array_interface__."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes // 8,), "typestr": "<i8", "data": (ptr, False), "version": 2}


def broadcast_evaluation_key(context, relin_key_host, src: int = 0, galois_keys_host=None, galois_elements=None):
    """Creates the EvaluationKey on every rank; the key material comes from rank `src`.

    GPU ranks: the key is uploaded on `src` and broadcast device-to-device into each rank's key buffer (NCCL).
    Returns a hecuda.EvaluationKey.  `relin_key_host` and `galois_keys_host` ({element: key}) are only read on rank
    `src` (others may pass None); `galois_elements` (the EvaluationKeyConfig, known to every rank) lists the Galois
    keys to broadcast.  has_relin: pass relin_key_host=None on every rank for a Galois-only key."""
    import torch
    import torch.distributed as dist

    import hecuda

    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    elements = list(galois_elements if galois_elements is not None else (galois_keys_host or {}))
    if world == 1:
        key = hecuda.EvaluationKey(context, relin_key_host)
        for e in elements:
            key.setGaloisKey(e, galois_keys_host[e])
        return key
    flag = torch.tensor([1 if (rank == src and relin_key_host is not None) else 0], device=torch.device("cuda", torch.cuda.current_device()))
    dist.broadcast(flag, src=src)
    has_relin = bool(flag.item())
    if rank == src:
        key = hecuda.EvaluationKey(context, relin_key_host)
        for e in elements:
            key.setGaloisKey(e, galois_keys_host[e])
    else:
        key = hecuda.EvaluationKey(context, None)  # empty device buffers of the right size
    device = torch.device("cuda", torch.cuda.current_device())
    buffers = ([key.deviceBuffer()] if has_relin else []) + [key.galoisDeviceBuffer(e) for e in elements]
    for ptr, nbytes in buffers:
        dist.broadcast(torch.as_tensor(_CudaView(ptr, nbytes), device=device), src=src)
    torch.cuda.synchronize()
    return key


def broadcast_key_bytes(relin_key_host, shape, src: int = 0) -> np.ndarray:
    """Host-side (gloo) variant used when the transport is CPU: returns the key as a numpy array on every rank."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return np.ascontiguousarray(relin_key_host, dtype=np.uint64)
    if dist.get_rank() == src:
        t = torch.from_numpy(np.ascontiguousarray(relin_key_host, dtype=np.uint64).view(np.int64).reshape(shape))
    else:
        t = torch.zeros(shape, dtype=torch.int64)
    dist.broadcast(t, src=src)
    return t.numpy().view(np.uint64)


def broadcast_galois_keys_bytes(galois_keys_host, elements, shape, src: int = 0) -> dict:
    """Host-side (gloo) variant for the Galois keys of an evaluation key: `elements` (the EvaluationKeyConfig) is known on
    every rank, the key material only on rank `src`.  Returns {element: key} on every rank."""
    return {int(e): broadcast_key_bytes(None if galois_keys_host is None else galois_keys_host[e], shape, src=src)
            for e in elements}


def shard_databases(entry_count: int, rank: int, world: int) -> Tuple[int, int]:
    """The entries [lo, hi) of an index-PIR database that rank `rank` serves: one contiguous shard per GPU, like the
    reference's offline KeywordDatabase sharding (KeywordDatabase.swift:56-110); a query for entry i goes to the rank
    whose range contains it and is answered there without any exchange."""
    return shard_range(entry_count, rank, world)


def sharded_apply(fn: Callable[..., np.ndarray], *batched_inputs: np.ndarray, gather: bool = True):
    """Runs `fn` on this rank's contiguous slice of every batched input; optionally all-gathers the outputs.

    `fn` is the per-shard engine call, e.g. lambda a, b: Bfv.mulAssign(ctx, a, b).  With gather=True every rank
    returns the full-batch result (outputs are independent ciphertexts, so this is a plain concatenation)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    batch = batched_inputs[0].shape[0]
    lo, hi = shard_range(batch, rank, world)
    local = fn(*[x[lo:hi] for x in batched_inputs])
    if world == 1 or not gather:
        return local
    item_shape = local.shape[1:]
    sizes = [shard_range(batch, r, world) for r in range(world)]
    max_items = max(h - l for l, h in sizes)
    pad = np.zeros((max_items,) + item_shape, dtype=np.uint64)
    pad[: hi - lo] = local
    mine = torch.from_numpy(pad.view(np.int64))
    if dist.get_backend() == "nccl":
        mine = mine.cuda()
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    out = np.concatenate([p.cpu().numpy().view(np.uint64)[: h - l] for p, (l, h) in zip(parts, sizes)], axis=0)
    return out
