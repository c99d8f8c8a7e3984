#!/usr/bin/env python
"""bench.py -- BFV ct x ct multiply throughput on B200 (BASELINE.json metric / config 2).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores

A "step" = one pass of the hot path (Bfv.mulAssign, Bfv+Multiply.swift:18-21) over one batch of 1024 synthetic
ciphertext pairs at N=8192 with 4 coefficient moduli (L=3 ciphertext moduli + the key-switch modulus).
`value` is device-resident throughput (inputs in HBM before the clock starts, CUDA events, max over ranks);
`e2e` is the same work through the host-pointer C-ABI call with pinned host buffers (H2D + D2H inside the clock).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "swift-homomorphic-encryption_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

# Parameter sets: the largest 55-bit NTT-friendly primes, descending -- what the reference's
# generatePrimes(significantBitCounts: [55,...], preferringSmall: false, nttDegree: N) returns (Scalar.swift:113-154);
# the first three at N=8192 are its predefined n_8192_logq_3x55 set (EncryptionParameters.swift:406-410); t = 557057
# is the 20-bit NTT-friendly plaintext modulus of RlweBenchmark (EncryptionParameters.swift:383).
Q8192 = [36028797018652673, 36028797017571329, 36028797017456641, 36028797017276417, 36028797017014273]
WORKLOADS = {
    # name: (N, coefficient moduli [q_0..q_{L-1}, q_ks], t, default batch)
    "C2": (8192, Q8192[:4], 557057, 1024),
    "C2-L4": (8192, Q8192[:5], 557057, 1024),
}


def workload_params(name):
    n, moduli, t, batch = WORKLOADS[name]
    return n, list(moduli), t, batch


def stage_model_bytes(n, L):
    """Algorithmic bytes per multiply: SURVEY.md section 8(d) stage model (28R + 3L) N w."""
    R = 2 * L + 1
    return (28 * R + 3 * L) * n * 8


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.rows = []
        self.idx = device_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for ts, r in self.rows if t0 <= ts <= t1 + 0.1] or [r for _, r in self.rows[-3:]]
        sm = sorted(int(float(r[1])) for r in rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in rows:
            for k, name in enumerate(names):
                if len(r) > 5 + k and r[5 + k].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(rows)}


def host_threads():
    """All host threads this process may use (torchrun pins OMP_NUM_THREADS=1, which is not what we want here)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_quota():
    """CPU bandwidth limit of this container in cores (cgroup v2 cpu.max), or None when unlimited / unknown."""
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        return None if quota == "max" else max(1, -(-int(quota) // int(period)))
    except Exception:
        return None


def best_thread_count(ctx, n, L):
    """All logical CPUs, half of them (one per physical core) or the container's CPU quota, whichever runs the oracle
    fastest on a short probe -- the CPU baseline should get its best configuration."""
    from oracle import oracle as orc

    full = host_threads()
    candidates = {full, max(1, full // 2)}
    if cpu_quota():
        candidates.add(min(full, cpu_quota()))
    candidates = sorted(candidates, reverse=True)
    best, best_rate = full, 0.0
    for c in candidates:
        k = 2 * c
        a = orc.fill_uniform(11, ctx.q, n, k * 2 * L).reshape(k, 2, L, n)
        b = orc.fill_uniform(12, ctx.q, n, k * 2 * L).reshape(k, 2, L, n)
        ctx.mul(a[:c], b[:c], threads=c)  # warm the threads / page in
        t0 = time.perf_counter()
        ctx.mul(a, b, threads=c)
        rate = k / (time.perf_counter() - t0)
        if rate > best_rate:
            best, best_rate = c, rate
    return best


def cpu_reference_throughput(n, moduli, t, budget_s=12.0, threads=0):
    """Times the oracle (C restatement of the Swift reference) on a bounded sample of the same workload."""
    from oracle import oracle as orc

    ctx = orc.Context(n, moduli, t)
    L = ctx.L
    cores = threads or best_thread_count(ctx, n, L)
    probe = max(1, min(cores, 8))
    a = orc.fill_uniform(1, ctx.q, n, probe * 2 * L).reshape(probe, 2, L, n)
    b = orc.fill_uniform(2, ctx.q, n, probe * 2 * L).reshape(probe, 2, L, n)
    t0 = time.perf_counter()
    ctx.mul(a, b, threads=cores)
    per_round = time.perf_counter() - t0  # `probe` multiplies in parallel
    sample = int(max(cores, min(4096, budget_s / max(per_round, 1e-4) * probe)))
    sample = (sample // cores) * cores or cores
    a = orc.fill_uniform(3, ctx.q, n, sample * 2 * L).reshape(sample, 2, L, n)
    b = orc.fill_uniform(4, ctx.q, n, sample * 2 * L).reshape(sample, 2, L, n)
    t0 = time.perf_counter()
    ctx.mul(a, b, threads=cores)
    dt = time.perf_counter() - t0
    return sample / dt, cores, sample, dt


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; no Swift toolchain here) on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, moduli, t, batch = workload_params(args.workload)
    from oracle import oracle as orc

    ctx = orc.Context(n, moduli, t)
    L = ctx.L
    cores = best_thread_count(ctx, n, L)
    sample = max(cores, 2 * cores)  # bounded per-step sample of the batch
    a = orc.fill_uniform(3, ctx.q, n, sample * 2 * L).reshape(sample, 2, L, n)
    b = orc.fill_uniform(4, ctx.q, n, sample * 2 * L).reshape(sample, 2, L, n)
    for _ in range(args.warmup):
        ctx.mul(a, b, threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.mul(a, b, threads=cores)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": "BFV ct*ct mults/sec at N=8192, 4 coefficient moduli", "value": value,
        "unit": "mult/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: Bfv<UInt64> ct*ct multiply N={n}, {len(moduli)} coefficient moduli "
                               f"(L={L}), CPU sample of {sample} ciphertext pairs per step"},
        "cpu_baseline": {"value": value, "unit": "mult/s", "cores": cores, "kind": "port", "cpu_quota_cores": cpu_quota(),
                         "sample": f"{sample} pairs/step x {args.steps} steps, OpenMP over pairs, C restatement of the "
                                   "Swift reference (no Swift toolchain on this box)"},
        "e2e": {"value": value, "unit": "mult/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import hecuda

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or hecuda.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    hecuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n, moduli, t, batch = workload_params(args.workload)
    if args.batch:
        batch = args.batch
    ctx = hecuda.Context(n, moduli, t)
    lib = hecuda.load_library()
    L = ctx.L
    dev = torch.device("cuda", local_rank)

    # synthetic device-resident inputs: uniform residues in [0, q_i) (valid ring elements; SURVEY.md 8(d) flavour i)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    qs = torch.tensor(moduli[:L], dtype=torch.int64, device=dev).view(1, 1, L, 1)

    def uniform(shape):
        x = torch.randint(0, 1 << 62, shape, generator=gen, device=dev, dtype=torch.int64)
        return (x % qs).contiguous()

    lhs, rhs = uniform((batch, 2, L, n)), uniform((batch, 2, L, n))
    out = torch.empty((batch, 3, L, n), dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream()

    def step():
        rc = lib.hecuda_bfv_multiply_device(ctx._h, lhs.data_ptr(), rhs.data_ptr(), out.data_ptr(), batch,
                                            stream.cuda_stream)
        if rc != 0:
            raise RuntimeError(lib.hecuda_last_error().decode())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = hecuda.kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.perf_counter()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    w1 = time.perf_counter()
    ms = ev0.elapsed_time(ev1)
    launches = hecuda.kernel_launch_count() - launches0
    clocks = sampler.stop(w0, w1) if rank == 0 else None
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    value = world * batch * args.steps / (ms_max / 1e3)

    # ---- roofline of the dominant kernel (forward NTT over the extended base: 28 of the 49 NTTs of a multiply),
    # timed alone with CUDA events on the launch stream, same launch shape as inside a step's pipeline stage.
    R = 2 * L + 1
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    ntt_polys = min(batch, 64) * 4
    buf = uniform((ntt_polys // 4 * 4, 1, L, n))[:, 0]  # any canonical residues do; reshape as R-row polys below
    ext = torch.zeros((ntt_polys, R, n), dtype=torch.int64, device=dev)
    ext[:, :L] = buf[:ntt_polys]
    reps = 20
    for _ in range(3):
        lib.hecuda_ntt_forward_device(ctx._h, hecuda.BASE_Q_AUX, ext.data_ptr(), R, ntt_polys, stream.cuda_stream)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    k0.record(stream)
    for _ in range(reps):
        lib.hecuda_ntt_forward_device(ctx._h, hecuda.BASE_Q_AUX, ext.data_ptr(), R, ntt_polys, stream.cuda_stream)
    k1.record(stream)
    torch.cuda.synchronize()
    ntt_ms = k0.elapsed_time(k1) / reps
    ntt_rows = ntt_polys * R
    ntt_bytes = ntt_rows * 2 * n * 8
    ntt_gbs = ntt_bytes / (ntt_ms / 1e3) / 1e9
    traffic = None
    tr_path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tr_path):
        tr = json.load(open(tr_path))
        if "ntt_forward_dram_bytes_per_row_narrow" in tr:  # per-row figures from the ncu --set full capture
            traffic = (tr["ntt_forward_dram_bytes_per_row_narrow"] * ntt_polys * L +
                       tr["ntt_forward_dram_bytes_per_row_wide"] * ntt_polys * (L + 1))
    roofline = {"bound": "hbm", "kernel": "ntt_forward (extended base [Q,Bsk])", "achieved": ntt_gbs, "peak": peak,
                "unit": "GB/s", "frac": ntt_gbs / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": ntt_bytes, "rows_per_launch": ntt_rows, "ntt_per_s": ntt_rows / (ntt_ms / 1e3),
                "whole_step_stage_model": {
                    "bytes_per_mult": stage_model_bytes(n, L),
                    "achieved_gbs": stage_model_bytes(n, L) * (value / world) / 1e9,
                    "frac": stage_model_bytes(n, L) * (value / world) / 1e9 / peak}}
    del ext, buf

    # ---- extra (not the headline): relinearize and multiply+relinearize on the same batch.  The relinearization
    # key is synthetic (uniform residues, valid Eval-format rows): rank 0 creates it and it reaches the other ranks
    # by one NCCL broadcast into their key buffers (the only collective of the deployment).
    extra = {}
    try:
        from hecuda import distributed as hd

        K = L + 1
        if rank == 0:
            kq = torch.tensor(moduli, dtype=torch.int64, device=dev).view(1, 1, K, 1)
            key_host = (torch.randint(0, 1 << 62, (L, 2, K, n), generator=gen, device=dev, dtype=torch.int64) % kq)
            key_host = key_host.cpu().numpy().view(np.uint64)
        else:
            key_host = None
        evk = hd.broadcast_evaluation_key(ctx, key_host, src=0)
        relin_out = torch.empty((batch, 2, L, n), dtype=torch.int64, device=dev)

        def relin_step():
            rc = lib.hecuda_bfv_relinearize_device(ctx._h, evk._h, out.data_ptr(), L, relin_out.data_ptr(), batch,
                                                   stream.cuda_stream)
            if rc != 0:
                raise RuntimeError(lib.hecuda_last_error().decode())

        for name, fn in (("relinearize_per_s", relin_step), ("multiply_relinearize_per_s", lambda: (step(), relin_step()))):
            for _ in range(2):
                fn()
            barrier()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record(stream)
            reps = max(3, min(args.steps, 10))
            for _ in range(reps):
                fn()
            r1.record(stream)
            barrier()
            tt = torch.tensor([r0.elapsed_time(r1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            extra[name] = world * batch * reps / (float(tt.item()) / 1e3)
        # lazy ct x pt inner product (MulPir first-dimension scan): 64 query ciphertexts against a
        # 256-row plaintext database streamed once from HBM -- the one HBM-bound kernel of the path
        ip_terms, ip_rows = 64, 256
        ip_cts = uniform((ip_terms, 2, L, n))
        ip_pts = (torch.randint(0, 1 << 62, (ip_rows, ip_terms, L, n), generator=gen, device=dev, dtype=torch.int64)
                  % qs.view(1, 1, L, 1)).contiguous()
        ip_out = torch.empty((ip_rows, 2, L, n), dtype=torch.int64, device=dev)

        def ip_step():
            rc = lib.hecuda_bfv_inner_product_plaintexts_device(ctx._h, ip_cts.data_ptr(), 2, L, ip_terms, ip_pts.data_ptr(),
                                                                None, ip_out.data_ptr(), ip_rows, stream.cuda_stream)
            if rc != 0:
                raise RuntimeError(lib.hecuda_last_error().decode())

        for _ in range(2):
            ip_step()
        torch.cuda.synchronize()
        i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        i0.record(stream)
        for _ in range(5):
            ip_step()
        i1.record(stream)
        torch.cuda.synchronize()
        ip_ms = i0.elapsed_time(i1) / 5
        ip_bytes = (ip_pts.numel() + ip_cts.numel() + ip_out.numel()) * 8
        extra["inner_product_plaintexts"] = {
            "shape": f"{ip_rows} rows x {ip_terms} terms, l={L}, N={n}", "ms": ip_ms,
            "ct_pt_products_per_s": ip_rows * ip_terms / (ip_ms / 1e3),
            "achieved_gbs": ip_bytes / (ip_ms / 1e3) / 1e9, "frac_of_hbm_peak": ip_bytes / (ip_ms / 1e3) / 1e9 / peak,
            "algorithmic_bytes": ip_bytes}
        del ip_pts, ip_cts, ip_out
        extra["ntt_forward_per_s_per_gpu"] = ntt_rows / (ntt_ms / 1e3)
        extra["key_broadcast"] = "nccl" if world > 1 else "local"
        evk.close()
        del relin_out
    except Exception as exc:  # the headline number must survive a failure of the extras
        extra["error"] = repr(exc)

    # ---- e2e: host buffers (pinned), H2D + D2H inside the timed region, through the host-pointer C-ABI call
    e2e = None
    if not args.no_e2e:
        e2e_batch = batch
        hl = hecuda.PinnedBuffer((e2e_batch, 2, L, n))
        hr = hecuda.PinnedBuffer((e2e_batch, 2, L, n))
        ho = hecuda.PinnedBuffer((e2e_batch, 3, L, n))
        hl.array[...] = lhs.cpu().numpy().view(np.uint64)
        hr.array[...] = rhs.cpu().numpy().view(np.uint64)
        e2e_steps = max(2, min(args.steps, 5))
        hecuda.Bfv.mulAssign(ctx, hl.array, hr.array, out=ho.array)  # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            hecuda.Bfv.mulAssign(ctx, hl.array, hr.array, out=ho.array)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tdt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tdt, op=dist.ReduceOp.MAX)
        e2e_ok = bool(np.array_equal(ho.array[:2], out[:2].cpu().numpy().view(np.uint64)))
        e2e = {"value": world * e2e_batch * e2e_steps / float(tdt.item()), "unit": "mult/s",
               "h2d_bytes_per_step": int(hl.array.nbytes + hr.array.nbytes), "d2h_bytes_per_step": int(ho.array.nbytes),
               "steps": e2e_steps, "timer": "host wall clock around blocking C-ABI calls, max over ranks",
               "matches_device_result": e2e_ok}
        hl.free(), hr.free(), ho.free()

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            v, cores, sample, dt = cpu_reference_throughput(n, moduli, t)
            cpu = {"value": v, "unit": "mult/s", "cores": cores, "kind": "port", "cpu_quota_cores": cpu_quota(),
                   "sample": f"{sample} ciphertext pairs of the same workload in {dt:.1f} s, OpenMP over pairs "
                             "(C restatement of the Swift reference; no Swift toolchain on this box)"}
        line = {
            "metric": "BFV ct*ct mults/sec at N=8192, 4 coefficient moduli", "value": value, "unit": "mult/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: Bfv<UInt64> ct*ct multiply (Bfv.mulAssign) N={n}, {len(moduli)} "
                                   f"coefficient moduli (L={L} ciphertext + key-switch), t={t}, batch={batch} pairs per GPU",
                       "batch_per_gpu": batch, "parallelism": f"batch-sharded x{world}, no data-path collective",
                       "l2": "inputs+outputs per step (1.4 GB) exceed L2 (126 MB); no explicit flush",
                       "pipeline_chunk": int(os.environ.get("HECUDA_CHUNK", "0")) or "auto"},
            "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "extra": extra,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
