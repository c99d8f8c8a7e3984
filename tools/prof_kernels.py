#!/usr/bin/env python
"""Launches the kernels that the headline bench does not reach, for ncu captures (development tool):
lazy ct x pt inner product (MulPir first-dimension scan shape), ct x ct inner product (tensor_sum), key switch at C3 shape.
   ncu --set full -k regex:... python tools/prof_kernels.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "swift-homomorphic-encryption_b200"))
import numpy as np
import torch
import hecuda
from bench import Q8192, Q16384

dev = torch.device("cuda", 0)
lib = hecuda.load_library()
s = torch.cuda.current_stream()
gen = torch.Generator(device=dev); gen.manual_seed(1)

def uniform(shape, moduli):
    qs = torch.tensor(list(moduli), dtype=torch.int64, device=dev).view(*([1] * (len(shape) - 2)), len(moduli), 1)
    return (torch.randint(0, 1 << 62, shape, generator=gen, device=dev, dtype=torch.int64) % qs).contiguous()

def ck(rc):
    assert rc == 0, lib.hecuda_last_error()

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "ip"):
    n, t, moduli = 8192, 557057, Q8192[:4]
    ctx = hecuda.Context(n, moduli, t); L = ctx.L
    terms, rows, pairs, groups = 64, 256, 16, 32
    cts = uniform((terms, 2, L, n), moduli[:L]); pts = uniform((rows, terms, L, n), moduli[:L])
    out = torch.empty((rows, 2, L, n), dtype=torch.int64, device=dev)
    for _ in range(3):
        ck(lib.hecuda_bfv_inner_product_plaintexts_device(ctx._h, cts.data_ptr(), 2, L, terms, pts.data_ptr(), None, out.data_ptr(), rows, s.cuda_stream))
    lhs = uniform((groups, pairs, 2, L, n), moduli[:L]); rhs = uniform((groups, pairs, 2, L, n), moduli[:L])
    out3 = torch.empty((groups, 3, L, n), dtype=torch.int64, device=dev)
    for _ in range(3):
        ck(lib.hecuda_bfv_inner_product_device(ctx._h, lhs.data_ptr(), rhs.data_ptr(), out3.data_ptr(), pairs, groups, s.cuda_stream))
    torch.cuda.synchronize()
if which in ("all", "c3"):
    n, t, moduli = 16384, 557057, Q16384
    ctx = hecuda.Context(n, moduli, t); L, K = ctx.L, ctx.L + 1
    batch = 256
    ct3 = uniform((batch, 3, L, n), moduli[:L])
    evk = hecuda.EvaluationKey(ctx, uniform((L, 2, K, n), moduli).cpu().numpy().view(np.uint64))
    relin = torch.empty((batch, 2, L, n), dtype=torch.int64, device=dev)
    down = torch.empty((batch, 2, L - 1, n), dtype=torch.int64, device=dev)
    for _ in range(3):
        ck(lib.hecuda_bfv_relinearize_device(ctx._h, evk._h, ct3.data_ptr(), L, relin.data_ptr(), batch, s.cuda_stream))
        ck(lib.hecuda_bfv_mod_switch_down_device(ctx._h, relin.data_ptr(), 2, L, down.data_ptr(), batch, s.cuda_stream))
    torch.cuda.synchronize()
print("done")
