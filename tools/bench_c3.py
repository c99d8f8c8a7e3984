#!/usr/bin/env python
"""BASELINE config 3 on one GPU: Bfv relinearize + modSwitchDown at N=16384 with 8 coefficient moduli (L=7, K=8).
Prints one JSON line (not the headline bench; see bench.py).  Synthetic uniform 3-poly ciphertexts and key."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "swift-homomorphic-encryption_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch

import hecuda

Q16384 = [36028797017456641, 36028797016178689, 36028797014704129, 36028797014573057, 36028797014376449,
          36028797014081537, 36028797013327873, 36028797013098497]  # largest 55-bit primes = 1 mod 32768


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    steps = 5
    n, t = 16384, 557057
    ctx = hecuda.Context(n, Q16384, t)
    lib = hecuda.load_library()
    L, K = ctx.L, ctx.L + 1
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(7)
    qs = torch.tensor(Q16384[:L], dtype=torch.int64, device=dev).view(1, 1, L, 1)
    kq = torch.tensor(Q16384, dtype=torch.int64, device=dev).view(1, 1, K, 1)
    ct3 = (torch.randint(0, 1 << 62, (batch, 3, L, n), generator=gen, device=dev, dtype=torch.int64) % qs).contiguous()
    key = (torch.randint(0, 1 << 62, (L, 2, K, n), generator=gen, device=dev, dtype=torch.int64) % kq).cpu().numpy().view(np.uint64)
    evk = hecuda.EvaluationKey(ctx, key)
    relin = torch.empty((batch, 2, L, n), dtype=torch.int64, device=dev)
    down = torch.empty((batch, 2, L - 1, n), dtype=torch.int64, device=dev)
    s = torch.cuda.current_stream()

    def step():
        rc = lib.hecuda_bfv_relinearize_device(ctx._h, evk._h, ct3.data_ptr(), L, relin.data_ptr(), batch, s.cuda_stream)
        rc |= lib.hecuda_bfv_mod_switch_down_device(ctx._h, relin.data_ptr(), 2, L, down.data_ptr(), batch, s.cuda_stream)
        assert rc == 0, lib.hecuda_last_error()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(steps):
        step()
    e1.record(s)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    bytes_per_unit = (2 * L * K + 5 * L + 8 * K) * n * 8 + (4 * L - 2) * n * 8  # 31 064 064 B
    peak = 6575.4
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    rate = batch / (ms / 1e3)
    print(json.dumps({"workload": f"C3: relinearize + modSwitchDown N={n}, 8 coefficient moduli (L=7), batch={batch}",
                      "units_per_s": rate, "ms_per_batch": ms, "stage_model_bytes_per_unit": bytes_per_unit,
                      "achieved_gbs": bytes_per_unit * rate / 1e9, "frac_of_hbm_peak": bytes_per_unit * rate / 1e9 / peak}))


if __name__ == "__main__":
    main()
