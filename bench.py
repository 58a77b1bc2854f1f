#!/usr/bin/env python
"""Benchmark of the similarity-search hot path (BASELINE.json metric: queries/sec of exact top-k
over an N-vector bank, plus % of the HBM / tensor roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c3s]

N = 1 workload (BASELINE.json configs[1], "C2"): 1M-vector x 768 bf16 bank resident in HBM,
64 queries, cosine, top-100.  One step = one pass of the hot path over one query batch
(pack queries -> tcgen05 scoring kernel with fused top-k -> merge of per-CTA candidates).
N > 1 (torchrun, one rank per GPU): every rank holds its own 1M-vector shard (weak scaling; the
bank is N x 1M vectors), each step adds one NCCL all-gather of the per-rank [Q, k] candidates and a
device merge.  `value` counts query-over-1M-vector-shard searches per second over all ranks
(= N*Q/t); `qps_global_bank` is Q/t over the whole N x 1M bank.

--impl reference times the reference's own CPU algorithm (oracle/ref_port.py, a torch port of
/root/reference/utils/similarity.py: the reference is pure Python and cannot travel to the GPU
box) on the host cores, each step a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "queries/sec exact top-k over N-vector bank"
WORKLOADS = {
    # name: (bank rows per GPU, D, Q, k, metric)
    "c2": (1_000_000, 768, 64, 100, "cosine"),
    "c2mse": (1_000_000, 768, 64, 100, "MSE"),
    "small": (100_000, 768, 64, 100, "cosine"),
    "mid": (100_000, 768, 512, 100, "cosine"),
    # the reference's own regime: one query (Q=1), and the widest single SIMT pass (Q=4)
    "q1": (1_000_000, 768, 1, 100, "cosine"),
    "q4": (1_000_000, 768, 4, 100, "cosine"),
    # BASELINE configs[2] ("C3"): 10M-vector bank, 4096-query batch, L2 (= unweighted MSE), tensor-pipe bound
    "c3": (10_000_000, 768, 4096, 100, "MSE"),
    "c3s": (1_000_000, 768, 4096, 100, "MSE"),
    # BASELINE configs[3] ("C4") per-GPU share at 8 GPUs: 12.5M of 100M vectors, 1000 queries, top-1000
    "c4": (12_500_000, 768, 1000, 1000, "cosine"),
    # C3 on 8 GPUs: the 10M-vector bank row-sharded, 1.25M vectors per GPU
    "c3g8": (1_250_000, 768, 4096, 100, "MSE"),
    # the reference's production call (scripts/done/sim.sh: -mp False): 64 patch tokens per item, one weighted
    # query, combine = min; 1M bank rows = 15625 items
    "l64": (1_000_000, 768, 1, 100, "cosine"),
}
TOKENS = {"l64": 64}


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def read_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the scoring kernel, from the committed
    ncu --set full capture of the SAME workload (profiles/traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get(key)
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the GPU works."""

    def __init__(self, index, period=0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.window = None          # (t0, t1) of the timed region, set by the caller
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:   # pragma: no cover
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        while not self._halt.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), mhz, util, r))
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._halt.set()

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        inside = [s for s in self.samples if self.window and self.window[0] <= s[0] <= self.window[1]]
        note = "sampled inside the timed region"
        if len(inside) < 3:
            inside = self.samples
            note = "timed region shorter than the sampling period: warm-up + timed + same-work burst window"
        return {"sm_mhz": statistics.median(s[1] for s in inside), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(inside), "note": note}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return rank, local, world


# ------------------------------------------------------------------------------------------------
# CPU baseline (the reference's algorithm on host cores)
# ------------------------------------------------------------------------------------------------
def cpu_sample(D, k, metric, rows, n_queries, batch=512, seed=0):
    """Time oracle/ref_port.multi_query_loop on `rows` bank rows x n_queries queries."""
    import torch
    from oracle import ref_port
    from sky_embeddings_b200 import synth
    s, m = synth.feature_profile(D)
    g = torch.Generator().manual_seed(seed)
    bank = torch.randn((rows, D), generator=g) * torch.from_numpy(s) + torch.from_numpy(m)
    mu, sd = bank[:batch].mean(0), bank[:batch].std(0)
    z = ((bank - mu) / (sd + 1e-8)).unsqueeze(1)
    q = z[:: max(rows // n_queries, 1), 0][:n_queries] + 0.1 * torch.randn((n_queries, D), generator=g)
    t0 = time.perf_counter()
    ref_port.multi_query_loop(q, None, z, batch, k, metric)
    return time.perf_counter() - t0


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU legs are to use every host core they can."""
    import torch
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def cpu_vectorised_sample(D, k, metric, rows, Q, seed=0):
    """A well-written CPU search, for context next to the reference port: the contraction form of the same metric
    (one fp32 GEMM over the normalised rows, all host threads) + torch.topk.  Not the reference's algorithm."""
    import torch
    from sky_embeddings_b200 import synth
    s, m = synth.feature_profile(D)
    g = torch.Generator().manual_seed(seed)
    bank = torch.randn((rows, D), generator=g) * torch.from_numpy(s) + torch.from_numpy(m)
    z = (bank - bank[:512].mean(0)) / (bank[:512].std(0) + 1e-8)
    q = z[:: max(rows // Q, 1)][:Q] + 0.1 * torch.randn((Q, D), generator=g)
    rn = (z * z).sum(1)
    t0 = time.perf_counter()
    dot = q @ z.T                                                     # [Q, rows]
    if metric == "cosine":
        sc = dot / (q.norm(dim=1, keepdim=True) * rn.sqrt()[None, :] + 1e-6)
        torch.topk(sc, min(k, rows), dim=1, largest=True)
    else:
        sc = ((q * q).sum(1, keepdim=True) - 2 * dot + rn[None, :]) / (D * D)
        torch.topk(sc, min(k, rows), dim=1, largest=False)
    return time.perf_counter() - t0


def cpu_baseline_block(n_bank, D, Q, k, metric, budget_s=15.0):
    import torch
    cores = use_all_host_threads()
    rows = min(n_bank, 100_000)
    cpu_sample(D, k, metric, min(rows, 20_000), 1)            # warm-up
    t1 = cpu_sample(D, k, metric, rows, 1)
    nq = int(max(1, min(16, budget_s / max(t1, 1e-3))))
    t = cpu_sample(D, k, metric, rows, nq) if nq > 1 else t1
    per_query_full = t / nq * (n_bank / rows)
    return {"value": 1.0 / per_query_full, "unit": "queries/s", "cores": cores, "kind": "port",
            "sample": f"oracle/ref_port.py (torch {torch.__version__} CPU port of utils/similarity.py, one pass per "
                      f"query as the reference requires): {nq} of {Q} queries over the first {rows} of {n_bank} "
                      f"bank rows, batch 512, {t:.2f} s, scaled linearly to the full bank"}


def run_reference(args, wl):
    rank, local, world = dist_env()
    if rank != 0:
        return 0
    import torch
    n_bank, D, Q, k, metric = wl
    cores = use_all_host_threads()
    # bounded sample: size one step so that the whole --steps/--warmup run ends within ~2 minutes
    nq = 1
    probe_rows = min(n_bank, 20_000)
    cpu_sample(D, k, metric, probe_rows, nq)
    rows_per_s = probe_rows / max(cpu_sample(D, k, metric, probe_rows, nq), 1e-6)
    budget = 120.0 / max(args.steps + args.warmup, 1)
    rows = int(min(n_bank, 100_000, max(4096, rows_per_s * budget)))
    for _ in range(args.warmup):
        cpu_sample(D, k, metric, rows, nq)
    times = []
    for _ in range(args.steps):
        times.append(cpu_sample(D, k, metric, rows, nq))
    per_step = sum(times) / len(times)
    # one step = nq queries over `rows` rows.  Same unit as the GPU arm's `value`: searches of one query over
    # one n_bank-vector shard per second (the CPU has no shards to spread over, so this does not grow with N)
    qps = nq / (per_step * (n_bank / rows))
    sample = (f"each step: {nq} query over the first {rows} of {n_bank * world} bank rows with oracle/ref_port.py "
              f"(torch {torch.__version__} CPU port of utils/similarity.py), scaled linearly")
    line = {"impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {n_bank}-vector x {D} bf16 bank per GPU ({n_bank * world} total), "
                                   f"{Q} queries, {metric} top-{k}, exact",
                       "bank_vectors_per_gpu": n_bank, "bank_vectors": n_bank * world, "dim": D, "queries": Q, "k": k,
                       "similarity": metric, "note": "reference algorithm on host cores (fp32), bounded sample scaled linearly"},
            "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def build_bank(n_rows, D, dev, row0_chunk=0, dtype="bf16", L=1):
    """Device-generated synthetic shard: chunks [row0_chunk, ...) of the global synthetic bank,
    normalised with the statistics of the first 512 rows of global chunk 0 on every rank.
    L > 1: every L consecutive rows form one item of L patch tokens (n_rows counts rows)."""
    from sky_embeddings_b200 import Bank, synth
    bank = Bank(n_rows // L, L, D, dtype, dev)
    first = synth.device_bank_chunk(0, 512, D, dev)
    bank.fit_norm(first.reshape(512 // L, L, D))
    done = 0
    c = row0_chunk
    while done < n_rows:
        rows = min(synth.CHUNK_ROWS, n_rows - done)
        bank.upload(synth.device_bank_chunk(c, rows, D, dev).reshape(rows // L, L, D), done // L)
        done += rows
        c += 1
    return bank.finalize()


def run_gpu(args, wl):
    import torch
    import torch.distributed as dist
    from sky_embeddings_b200 import _lib, synth
    from sky_embeddings_b200.distributed import CandidateExchange

    rank, local, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the search path has no CPU fallback")
    _lib.load()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_bank, D, Q, k, metric = wl
    L = TOKENS.get(args.workload, 1)
    chunks_per_rank = (n_bank + synth.CHUNK_ROWS - 1) // synth.CHUNK_ROWS
    row_lo = rank * n_bank
    bank = build_bank(n_bank, D, dev, row0_chunk=rank * chunks_per_rank, dtype=args.bank_dtype, L=L)
    esz = 2 if args.bank_dtype == "bf16" else 4

    # queries: planted neighbours of rows of THIS process's rank-0 shard layout (same on every rank:
    # generated from global chunk 0 with the shared statistics)
    probe = build_bank(min(n_bank, synth.CHUNK_ROWS), D, dev, row0_chunk=0)
    stride = probe.n_items // Q
    use_tc = args.path in ("tensor", "batch") or (args.path == "auto" and Q >= 2 and args.bank_dtype == "bf16" and not args.weighted)
    gen = torch.Generator(device=dev).manual_seed(1234)
    planted = [q * stride + stride // 2 for q in range(Q)]
    t_dev = torch.cat([probe.download(r, 1)[:, 0] for r in planted])
    t_dev = t_dev + 0.1 * torch.randn((Q, D), generator=gen, device=dev)
    probe.close()
    w_dev = None
    if args.weighted:     # per-query inverse-variance-like weights, normalised to sum 1 (utils/similarity.py:143-145)
        w_dev = torch.rand((Q, D), generator=gen, device=dev) + 0.5
        w_dev = w_dev / w_dev.sum(1, keepdim=True)
    w_host = w_dev.cpu().pin_memory() if w_dev is not None else None
    t_host = t_dev.cpu().pin_memory()
    out_s_host = torch.empty((Q, k), dtype=torch.float32).pin_memory()
    out_i_host = torch.empty((Q, k), dtype=torch.int64).pin_memory()

    xchg = CandidateExchange(Q, k, dev) if world > 1 else None

    def step_device():
        if world == 1:
            return bank.search(t_dev, w_dev, k=k, metric=metric, path=args.path, idx_offset=row_lo)
        # local top-k straight into the exchange buffer, ONE all-gather (scores + indices), in-place strided merge
        bank.search(t_dev, w_dev, k=k, metric=metric, path=args.path, idx_offset=row_lo,
                    out_scores=xchg.scores, out_idx=xchg.idx)
        return xchg.merge(metric)

    def step_host():
        # public API with HOST buffers: H2D of the queries and D2H of the results inside the call
        if world == 1:
            return bank.search_host(t_host, w_host, k=k, metric=metric, path=args.path, idx_offset=row_lo,
                                    out_scores=out_s_host, out_idx=out_i_host)
        td = t_host.to(dev, non_blocking=True)
        wd = w_host.to(dev, non_blocking=True) if w_host is not None else None
        bank.search(td, wd, k=k, metric=metric, path=args.path, idx_offset=row_lo, out_scores=xchg.scores, out_idx=xchg.idx)
        s, i = xchg.merge(metric)
        out_s_host.copy_(s, non_blocking=True)
        out_i_host.copy_(i, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out_s_host, out_i_host

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()

    # correctness gate before timing: the planted rows must come back first on rank 0's shard
    s, i = step_device()
    torch.cuda.synchronize()
    if L == 1 and not int(os.environ.get("SKY_TC_DEBUG", "0")):      # profiling experiments disable parts of the kernel
        assert i[:, 0].cpu().tolist() == planted, "planted nearest neighbours not returned: refusing to time a wrong kernel"

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    bank.profile(True)
    bank.profile_read(reset=True)
    _lib.launch_count(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    barrier()
    t_wall1 = time.perf_counter()
    sampler.window = (t_wall0, t_wall1)
    launches = _lib.launch_count(reset=True)
    ms_total = ev0.elapsed_time(ev1)
    n_kern, kern_ms = bank.profile_read(reset=True)
    bank.profile(False)
    if world > 1:
        tt = torch.tensor([ms_total], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
    ms_step = ms_total / args.steps

    # end-to-end through the host-buffer API
    for _ in range(3):
        step_host()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_host()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item())
    e2e_step = e2e_ms / args.steps

    # keep the GPU busy with the same work long enough for the clock sampler if the run was short
    if (t_wall1 - t_wall0) < 0.05:
        t_end = time.perf_counter() + 0.3
        while time.perf_counter() < t_end:
            step_device()
        torch.cuda.synchronize()
    sampler.stop()
    sampler.join(timeout=1.0)

    if rank == 0:
        peak, peak_src = read_peaks()
        kern_avg_ms = kern_ms / max(n_kern, 1)
        algo_bytes = float(n_bank) * D * esz                    # one pass over this GPU's shard
        achieved = algo_bytes / (kern_avg_ms * 1e-3) / 1e9 if kern_avg_ms > 0 else 0.0
        use_batch = use_tc and (args.path == "batch" or (args.path == "auto" and Q > 128))
        value = world * Q / (ms_step * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.bank_dtype, "data": "synthetic",
            "config": {"workload": f"{args.workload}: {n_bank}-vector x {D} {args.bank_dtype} bank per GPU ({n_bank * world} total), "
                                   f"{Q} queries, {metric} top-{k}, exact",
                       "bank_vectors_per_gpu": n_bank, "bank_vectors": n_bank * world, "dim": D, "queries": Q, "k": k,
                       "similarity": metric, "weighted": bool(args.weighted), "path": args.path, "parallelism": f"row-shard x{world}",
                       "l2_policy": f"bank shard ({algo_bytes / 1e6:.0f} MB) is larger than the 126 MB L2; no flush needed",
                       "value_definition": "N_gpus * Q / step time (query-over-1M-vector-shard searches/s; at N=1 plain QPS)"},
            "qps_global_bank": Q / (ms_step * 1e-3),
            "clocks": sampler.summary(),
            "e2e": {"value": world * Q / (e2e_step * 1e-3), "unit": "queries/s", "ms_per_step": e2e_step,
                    "h2d_bytes_per_step": t_host.numel() * 4 * (2 if args.weighted else 1), "d2h_bytes_per_step": Q * k * 12,
                    "api": "sky_search_host (C ABI, pinned host buffers)" if world == 1 else
                           "pinned H2D + sky_search + one NCCL all-gather + sky_merge_candidates_strided + D2H"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None, "traffic": read_traffic("tensor" if (use_tc and args.workload == "c2") else
                                                 ("stream_fp32_q1" if (args.workload == "q1" and args.bank_dtype == "fp32") else "none")),
                         "kernel": "tc_search_kernel<64>" if use_tc else "stream_search_kernel",
                         "kernel_ms": kern_avg_ms, "kernel_launches": n_kern, "algorithmic_bytes": algo_bytes,
                         "peak_source": peak_src, "kernel_share_of_step": kern_avg_ms * (n_kern / args.steps) / ms_step},
        }
        if use_batch:
            # tensor-pipe bound: all phase launches of one search together; flops = 2 Q N D
            tf_peak = 1392.2
            try:
                tf_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
            except Exception:
                pass
            kern_search_ms = kern_ms / args.steps
            flops = 2.0 * Q * float(n_bank) * D
            ach = flops / (kern_search_ms * 1e-3) / 1e12
            line["roofline"] = {"bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak,
                                "traffic": None, "kernel": "tc_batch_kernel (all phases of one search)",
                                "kernel_ms": kern_search_ms, "kernel_launches": n_kern, "algorithmic_flops": flops,
                                "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops_sustained)",
                                "kernel_share_of_step": kern_search_ms / ms_step}
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_block(n_bank, D, Q, k, metric)
            if L == 1:      # context only: what a vectorised CPU implementation (GEMM + topk) reaches on the same host
                rows = min(n_bank, 100_000)
                qs = min(Q, 64)
                cpu_vectorised_sample(D, k, metric, rows, qs)
                dt = cpu_vectorised_sample(D, k, metric, rows, qs)
                line["cpu_vectorised"] = {"value": qs / (dt * n_bank / rows), "unit": "queries/s",
                                          "cores": line["cpu_baseline"]["cores"], "kind": "context",
                                          "sample": f"torch fp32 GEMM + topk, {qs} queries over {rows} of {n_bank} rows, "
                                                    f"{dt:.2f} s, scaled linearly; not the reference's algorithm"}
        print(json.dumps(line), flush=True)
    bank.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------------------------------
# pixel-space masked-MSE workloads (BASELINE config 5): raw 5 x 64 x 64 fp32 cutouts, HBM bound
# ------------------------------------------------------------------------------------------------
PIXEL_WORKLOADS = {"c5": (1_000_000, 4, 100), "c5s": (100_000, 4, 100), "c5q1": (1_000_000, 1, 100)}


def pixel_chunk(dev, rows, chunk_index, seed=20240607):
    """Device-generated synthetic cutouts ~ N(0,1) clipped at -3 with 2% NaN pixels and 5% missing bands
    (same statistics as sky_embeddings_b200.synth.cutouts; row content depends only on (seed, chunk))."""
    import torch
    g = torch.Generator(device=dev).manual_seed(seed * 1000003 + chunk_index)
    x = torch.randn((rows, 5, 64, 64), generator=g, device=dev).clamp_(min=-3.0)
    x[torch.rand((rows, 5, 64, 64), generator=g, device=dev) < 0.02] = float("nan")
    x[torch.rand((rows, 5), generator=g, device=dev) < 0.05] = float("nan")
    return x


def run_gpu_pixels(args):
    import numpy as np
    import torch
    from sky_embeddings_b200 import PixelBank, _lib
    rank, local, world = dist_env()
    if world > 1:
        raise SystemExit("pixel workloads are single-GPU bench lines")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the search path has no CPU fallback")
    _lib.load()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    n, Q, k = PIXEL_WORKLOADS[args.workload]
    D, chunk = 5 * 64 * 64, 8192
    bank = PixelBank(n, 5, 64, 64, device=dev)
    for c, s in enumerate(range(0, n, chunk)):
        bank.upload(pixel_chunk(dev, min(chunk, n - s), c), s)
    planted = [(2 * q + 1) * 1000 % n for q in range(Q)]
    first = torch.cat([pixel_chunk(dev, chunk, r // chunk)[r % chunk][None] for r in planted])
    gen = torch.Generator(device=dev).manual_seed(99)
    q_dev = first + 0.2 * torch.randn(first.shape, generator=gen, device=dev)
    q_host = q_dev.cpu().pin_memory()
    out_s = torch.empty((Q, k), dtype=torch.float32).pin_memory()
    out_i = torch.empty((Q, k), dtype=torch.int64).pin_memory()

    def step_device():
        return bank.search(q_dev, None, k=k)

    def step_host():
        s, i = bank.search(q_host.to(dev, non_blocking=True), None, k=k)
        out_s.copy_(s, non_blocking=True)
        out_i.copy_(i, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    s, i = step_device()
    torch.cuda.synchronize()
    assert i[:, 0].cpu().tolist() == planted, "planted nearest cutouts not returned: refusing to time a wrong kernel"
    steps = args.steps
    for _ in range(max(args.warmup, 3)):
        step_device()
    torch.cuda.synchronize()
    bank.profile(True)
    bank.profile_read(reset=True)
    _lib.launch_count(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(steps):
        step_device()
    ev1.record()
    torch.cuda.synchronize()
    sampler.window = (t0, time.perf_counter())
    launches = _lib.launch_count(reset=True)
    ms_step = ev0.elapsed_time(ev1) / steps
    n_kern, kern_ms = bank.profile_read(reset=True)
    bank.profile(False)
    for _ in range(3):
        step_host()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step_host()
    e1.record()
    torch.cuda.synchronize()
    e2e_step = e0.elapsed_time(e1) / steps
    sampler.stop()
    sampler.join(timeout=1.0)
    peak, peak_src = read_peaks()
    kern_avg = kern_ms / max(n_kern, 1)
    algo = float(n) * D * 4
    achieved = algo / (kern_avg * 1e-3) / 1e9
    line = {"metric": METRIC, "value": Q / (ms_step * 1e-3), "unit": "queries/s", "n_gpus": 1, "steps": steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {n} cutouts 5x64x64 fp32 ({algo / 1e9:.1f} GB), {Q} queries, "
                                   f"pixel-space NaN-aware masked MSE, top-{k}, exact",
                       "bank_vectors": n, "dim": D, "queries": Q, "k": k, "similarity": "masked MSE (pixels)",
                       "l2_policy": "bank is larger than the 126 MB L2; no flush needed"},
            "clocks": sampler.summary(),
            "e2e": {"value": Q / (e2e_step * 1e-3), "unit": "queries/s", "ms_per_step": e2e_step,
                    "h2d_bytes_per_step": Q * D * 4, "d2h_bytes_per_step": Q * k * 12,
                    "api": "PixelBank.search with pinned host queries / results"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "kernel": "pixel_search_kernel", "kernel_ms": kern_avg,
                         "kernel_launches": n_kern, "algorithmic_bytes": algo, "peak_source": peak_src,
                         "kernel_share_of_step": kern_avg * (n_kern / steps) / ms_step}}
    if not args.no_cpu:
        from oracle import sky_oracle as O
        rows = 2000
        xs = pixel_chunk(dev, rows, 0).cpu().numpy()
        qs = q_dev[0].cpu().numpy()
        O.pixel_masked_mse(qs, xs[:200])
        t0 = time.perf_counter()
        O.pixel_masked_mse(qs, xs, dtype=np.float32)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": 1.0 / (dt * n / rows), "unit": "queries/s", "cores": 1, "kind": "port",
                                "sample": f"oracle/sky_oracle.pixel_masked_mse (numpy fp32) over {rows} of {n} cutouts, "
                                          f"1 query, {dt:.2f} s, scaled linearly"}
    print(json.dumps(line), flush=True)
    bank.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + sorted(PIXEL_WORKLOADS))
    ap.add_argument("--path", default="auto", choices=["auto", "simt", "tensor", "batch", "generic"])
    ap.add_argument("--bank-dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--weighted", action="store_true", help="per-query feature weights (use_weights=True)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    args = ap.parse_args()
    if args.workload in PIXEL_WORKLOADS:
        if args.impl == "reference":
            raise SystemExit("--impl reference is defined for the headline workload only")
        return run_gpu_pixels(args)
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, wl)
    return run_gpu(args, wl)


if __name__ == "__main__":
    sys.exit(main())
