#!/usr/bin/env python
"""Device-side rate of the wire-format codec (PolyRq.load / serialize) on C2-shaped polynomials: prints GB/s of
unpacked words produced / consumed.  python tools/bench_codec.py [polys]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "swift-homomorphic-encryption_b200")):
    sys.path.insert(0, p)
import torch

import hecuda

Q8192 = [36028797018652673, 36028797017571329, 36028797017456641, 36028797017276417]
polys = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = 8192
ctx = hecuda.Context(n, Q8192, 557057)
L = ctx.L
lib = hecuda.load_library()
size = hecuda.Bfv.serializationByteCount(ctx, L)
dev = torch.device("cuda", 0)
words = torch.randint(0, 1 << 54, (polys, L, n), dtype=torch.int64, device=dev)
packed = torch.empty((polys, size), dtype=torch.uint8, device=dev)
back = torch.empty_like(words)
s = torch.cuda.current_stream()


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(reps):
        fn()
    e1.record(s)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ser = timed(lambda: lib.hecuda_poly_serialize_device(ctx._h, 0, words.data_ptr(), 0, packed.data_ptr(), L, polys, s.cuda_stream))
load = timed(lambda: lib.hecuda_poly_load_device(ctx._h, 0, packed.data_ptr(), 0, back.data_ptr(), L, polys, s.cuda_stream))
assert torch.equal(back, words)
gb = words.numel() * 8 / 1e9
print(json.dumps({"metric": "PolyRq wire-format codec on device", "config": {"workload": f"{polys} polys of {L} x {n}, 55-bit moduli"},
                  "serialize_ms": round(ser, 3), "load_ms": round(load, 3), "serialize_gbs_words": round(gb / ser * 1e3, 1),
                  "load_gbs_words": round(gb / load * 1e3, 1), "bytes_per_poly_packed": size, "bytes_per_poly_words": L * n * 8}))
