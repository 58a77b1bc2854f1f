#!/usr/bin/env python
"""Benchmark of the similarity-search hot path (BASELINE.json metric: queries/sec of exact top-k over an
N-vector bank, plus % of the HBM / tensor roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME] [--also LIST]

Headline (every N): BASELINE.json configs[1] ("C2") -- 1M-vector x 768 bf16 bank per GPU resident in HBM, 64
queries, cosine, top-100.  One step = one pass of the hot path over one query batch.  N > 1 (torchrun, one rank
per GPU) is WEAK scaling for the headline: every rank holds its own 1M-vector shard, each step ends with the
candidate exchange over NVLink and the device merge; `value` = N*Q/t, `qps_global_bank` = Q/t.

The same JSON line carries one sub-record per other BASELINE config under "also" (each with its own parity gate,
roofline, e2e, clocks and launch count):
    N = 1:  c2w  (C2 with per-query weights, the reference's use_weights=True), q1w (one weighted fp32 query: the
            reference's own regime), c3 (10M x 768, 4096 queries, L2, top-100), c5q1 (1M cutouts 5x64x64 fp32,
            pixel-space masked MSE)
    N > 1:  c3 and c4 (100M x 768, 1000 queries, top-1000) STRONG-scaled: the global bank row-sharded over the N
            ranks, value = Q/t over the global bank.
Before anything is timed, 8 sampled queries are checked over their full top-k against an fp32 torch scoring of the
stored bank (tests/torch_ref.py), and at N > 1 the exchanged + merged result is compared bit for bit with a torch
merge of the independently all-gathered per-rank candidates.

--impl reference times the reference's OWN mae_simsearch (the unmodified utils/similarity.py from oracle/_ref, see
oracle/ref_harness.py) on the host cores, each step one query over a bounded sample of the same bank.
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
    # name: dict(rows per GPU (weak) or global rows (strong), D, Q, k, metric, ...)
    "c2": dict(n=1_000_000, D=768, Q=64, k=100, metric="cosine"),
    "c2w": dict(n=1_000_000, D=768, Q=64, k=100, metric="cosine", weighted=True),
    "c2mse": dict(n=1_000_000, D=768, Q=64, k=100, metric="MSE"),
    "small": dict(n=100_000, D=768, Q=64, k=100, metric="cosine"),
    "mid": dict(n=100_000, D=768, Q=512, k=100, metric="cosine"),
    # the reference's own regime: one query with per-feature weights on the fp32 embeddings
    "q1w": dict(n=1_000_000, D=768, Q=1, k=100, metric="cosine", weighted=True, dtype="fp32"),
    "q1": dict(n=1_000_000, D=768, Q=1, k=100, metric="cosine"),
    "q1wb": dict(n=1_000_000, D=768, Q=1, k=100, metric="cosine", weighted=True),          # the same on a bf16 bank
    "q4": dict(n=1_000_000, D=768, Q=4, k=100, metric="cosine"),
    # BASELINE configs[2] ("C3"): 10M-vector bank, 4096-query batch, L2 (= unweighted MSE), tensor-pipe bound
    "c3": dict(n=10_000_000, D=768, Q=4096, k=100, metric="MSE", steps=10),
    "c3s": dict(n=1_000_000, D=768, Q=4096, k=100, metric="MSE", steps=20),
    # BASELINE configs[3] ("C4"): 100M vectors sharded over the ranks, 1000 queries, top-1000
    "c4": dict(n=100_000_000, D=768, Q=1000, k=1000, metric="cosine", steps=5),
    # the per-GPU shares of C3 / C4 at 8 GPUs as single-GPU workloads (kernel experiments)
    "c3g8": dict(n=1_250_000, D=768, Q=4096, k=100, metric="MSE", steps=20),
    "c4g8": dict(n=12_500_000, D=768, Q=1000, k=1000, metric="cosine", steps=5),
    "c4g2": dict(n=50_000_000, D=768, Q=1000, k=1000, metric="cosine", steps=3),      # 76.8 GB of bank on one GPU
    # the reference's production call (scripts/done/sim.sh: -mp False): 64 patch tokens per item, one weighted
    # query, combine = min; 1M bank rows = 15625 items
    "l64": dict(n=1_000_000, D=768, Q=1, k=100, metric="cosine", weighted=True, dtype="fp32", L=64),
}
PIXEL_WORKLOADS = {"c5": (1_000_000, 4, 100), "c5s": (100_000, 4, 100), "c5q1": (1_000_000, 1, 100), "c5q1s": (100_000, 1, 100)}
ALSO_N1 = ["c2w", "q1w", "c3", "c5q1"]
ALSO_MULTI = ["c3", "c4"]
KNOBS = ("SKY_PDL", "SKY_PX_QC", "SKY_PX_STAGES", "SKY_SCORE_GENERIC", "SKY_ST_DEBUG", "SKY_ST_POLICY", "SKY_ST_SPIN",
         "SKY_ST_SPLIT", "SKY_ST_STAGES", "SKY_TB_DEBUG", "SKY_TB_DENSE0", "SKY_TB_GROWTH", "SKY_TB_PHASE0", "SKY_TB_SORTED", "SKY_TB_TRACE_PHASE", "SKY_TC_DEBUG",
         "SKY_TC_GRID", "SKY_TC_POLICY", "SKY_TC_STAGES", "SKY_TC_TMA", "SKY_TW2_OCC", "SKY_TW2_POLICY", "SKY_TW2_SPIN",
         "SKY_TW2_STAGES", "SKY_TW_DEBUG", "SKY_TW_PAIR")


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    out = {"hbm": 6650.0, "hbm_src": "fallback (B200_PROFILING.md 6.65 TB/s)", "tf": 1400.0,
           "tf_src": "fallback (B200_PROFILING.md)"}
    try:
        d = json.load(open(p))
        out["hbm"], out["hbm_src"] = float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        out["tf"], out["tf_src"] = float(d["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    except Exception:
        pass
    return out


def read_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the scoring kernel, from the committed
    ncu --set full capture of the SAME workload (profiles/traffic.json) -- not measured in this run -- or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(key)
    except Exception:
        return None


def active_knobs():
    return sorted(k for k in KNOBS if os.environ.get(k))


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the GPU works."""

    def __init__(self, index, period=0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.max_mhz = [], None
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
        while not self._halt.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), mhz, r))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._halt.set()

    def summary(self, window):
        names = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        inside = [s for s in self.samples if window[0] <= s[0] <= window[1]]
        note = "sampled inside the timed region"
        if len(inside) < 3:
            inside = [s for s in self.samples if window[0] - 0.5 <= s[0] <= window[1] + 0.5] or self.samples
            note = "timed region shorter than the sampling period: +-0.5 s window of the same work around it"
        reasons = sorted(n for n, bit in names.items() if any(s[2] & bit for s in inside))
        return {"sm_mhz": statistics.median(s[1] for s in inside), "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(inside), "note": note}


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def workload_config(name, wl, world, strong):
    """The `config` object; identical for the GPU arm and the reference arm of the same workload."""
    n, D, Q, k, metric = wl["n"], wl["D"], wl["Q"], wl["k"], wl["metric"]
    dtype = wl.get("dtype", "bf16")
    per_gpu = n // world if strong else n
    total = n if strong else n * world
    L = wl.get("L", 1)
    return {"workload": f"{name}: {total}-vector x {D} {dtype} bank ({per_gpu} per GPU), {Q} queries, "
                        f"{'weighted ' if wl.get('weighted') else ''}{metric} top-{k}, exact",
            "bank_vectors_per_gpu": per_gpu, "bank_vectors": total, "dim": D, "queries": Q, "k": k, "similarity": metric,
            "weighted": bool(wl.get("weighted")), "tokens_per_item": L, "bank_dtype": dtype,
            "parallelism": f"row-shard x{world}",
            "l2_policy": f"bank shard ({per_gpu * D * (2 if dtype == 'bf16' else 4) / 1e6:.0f} MB) is larger than the 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------------------------
# synthetic data (shared by both arms)
# ------------------------------------------------------------------------------------------------
def raw_chunk(c, rows, D, dev):
    """Rows of global chunk c of the synthetic bank as fp32 on `dev` (a CUDA device, or "cpu" for the CPU-only
    contract test: a different generator, same distribution)."""
    from sky_embeddings_b200 import synth
    return synth.device_bank_chunk(c, rows, D, dev)


def norm_stats(D, dev):
    """First-batch statistics (utils/similarity.py:98-100) of the first 512 rows of chunk 0, torch fp32."""
    first = raw_chunk(0, 512, D, dev)
    return first.mean(0), first.std(0)


def make_queries(n_total, D, Q, dev, weighted, gen_seed=1234):
    """Q planted queries in NORMALISED space: z(row r_q) + 0.1 N(0,1), rows at a fixed stride over the global bank
    (the planted row is the known top-1 answer); optional per-query weights, normalised to
    sum 1 like utils/similarity.py:143-145.  Returns (t [Q,D], w [Q,D] or None, planted rows)."""
    import torch
    from sky_embeddings_b200 import synth
    mu, sd = norm_stats(D, dev)
    stride = max(n_total // Q, 1)
    planted = [(q * stride + stride // 2) % n_total for q in range(Q)]
    cache = {}
    rows = []
    for r in planted:
        c, off = divmod(r, synth.CHUNK_ROWS)
        if c not in cache:
            cache = {c: raw_chunk(c, min(synth.CHUNK_ROWS, n_total - c * synth.CHUNK_ROWS), D, dev)}
        rows.append(cache[c][off].clone())          # a copy: a view would keep the whole 200 MB chunk alive (1000 chunks at C4)
    z = (torch.stack(rows) - mu) / (sd + 1e-8)
    gen = torch.Generator(device=dev).manual_seed(gen_seed)
    t = z + 0.1 * torch.randn((Q, D), generator=gen, device=dev)
    w = None
    if weighted:
        w = torch.rand((Q, D), generator=gen, device=dev) + 0.5
        w = w / w.sum(1, keepdim=True)
    return t.contiguous(), w, planted


def shard_pieces(row0, n_rows, n_total):
    """Rows [row0, row0 + n_rows) of the n_total-row synthetic bank as pieces of its generator chunks:
    (chunk index, rows the chunk is generated with, first row inside the chunk, rows taken, destination row in the shard).
    A chunk is always generated with its canonical size min(CHUNK_ROWS, n_total - chunk * CHUNK_ROWS) -- the values of a
    torch Philox draw depend on its shape -- so a row is the same tensor whichever rank holds it and wherever the shard
    boundaries fall (make_queries draws the planted rows the same way)."""
    from sky_embeddings_b200 import synth
    out, done = [], 0
    while done < n_rows:
        c, off = divmod(row0 + done, synth.CHUNK_ROWS)
        chunk_rows = min(synth.CHUNK_ROWS, n_total - c * synth.CHUNK_ROWS)
        take = min(chunk_rows - off, n_rows - done)
        assert take > 0, "shard reaches past the end of the bank"
        out.append((c, chunk_rows, off, take, done))
        done += take
    return out


def build_bank(n_rows, D, dev, chunk0=0, dtype="bf16", L=1, row0=None, n_total=None):
    """Device-generated synthetic shard, normalised with the statistics of the first 512 rows of global chunk 0 on every
    rank.  Default: n_rows rows from generator chunk `chunk0` on (a stand-alone bank of n_rows rows).  With row0 / n_total:
    rows [row0, row0 + n_rows) of the n_total-row global bank, any row0 (strong scaling: equal shards).  L > 1: every L
    consecutive rows form one item of L patch tokens (n_rows counts rows)."""
    from sky_embeddings_b200 import Bank, synth
    bank = Bank(n_rows // L, L, D, dtype, dev)
    bank.fit_norm(raw_chunk(0, 512, D, dev).reshape(512 // L, L, D))
    if row0 is not None:
        for c, chunk_rows, off, take, dst in shard_pieces(row0, n_rows, n_total):
            bank.upload(raw_chunk(c, chunk_rows, D, dev)[off:off + take].reshape(take // L, L, D), dst // L)
        return bank.finalize()
    done, c = 0, chunk0
    while done < n_rows:
        rows = min(synth.CHUNK_ROWS, n_rows - done)
        bank.upload(raw_chunk(c, rows, D, dev).reshape(rows // L, L, D), done // L)
        done += rows
        c += 1
    return bank.finalize()


# ------------------------------------------------------------------------------------------------
# CPU legs: the reference's own mae_simsearch on the host cores
# ------------------------------------------------------------------------------------------------
def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU legs are to use every host core they can."""
    import torch
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


class CpuReference:
    """One query of the bench workload through the reference's unmodified mae_simsearch (utils/similarity.py:37-132)
    over the first `rows` rows of the SAME synthetic bank the GPU arm holds (batch 512 = the rows that define the
    first-batch normalisation on both sides).  Falls back to oracle/ref_port.search_loop when oracle/_ref is absent."""

    def __init__(self, wl, dev, rows=100_000, batch=512):
        import torch
        from oracle import ref_harness as H
        from sky_embeddings_b200 import synth
        self.torch, self.H = torch, H
        self.wl, self.batch = wl, batch
        self.cores = use_all_host_threads()
        self.ref = H.load_reference()
        self.kind = "reference" if self.ref is not None else "port"
        D = wl["D"]
        self.rows = min(wl["n"], rows)
        parts, done, c = [], 0, 0
        while done < self.rows:
            r = min(synth.CHUNK_ROWS, self.rows - done)
            parts.append(raw_chunk(c, r, D, dev).cpu())
            done += r
            c += 1
        self.bank = torch.cat(parts)
        mu, sd = norm_stats(D, dev)
        self.mu, self.sp = mu.cpu(), (sd + 1e-8).cpu()
        t, w, self.planted = make_queries(wl["n"], D, min(wl["Q"], 8), dev, wl.get("weighted", False))
        self.t, self.w = t.cpu(), (w.cpu() if w is not None else None)

    def one_query(self, q=0, rows=None):
        """Seconds for one query over the first `rows` sample rows; returns (seconds, top-1 index)."""
        torch, H = self.torch, self.H
        bank = self.bank if rows is None else self.bank[:rows]
        q = q % self.t.shape[0]
        w = self.w[q] if self.w is not None else None
        t0 = time.perf_counter()
        if self.ref is not None:
            sc, ix = H.reference_search(self.ref, bank, self.t[q], w, self.mu, self.sp, self.batch, self.wl["k"], self.wl["metric"])
        else:
            from oracle import ref_port
            tgt = H.target_group_for(self.t[q], w, self.mu, self.sp)
            sc, ix = ref_port.search_loop(tgt, bank[:, None, :], self.batch, self.wl["k"], self.wl["metric"], "min",
                                          use_weights=w is not None, cls_token=True)
        return time.perf_counter() - t0, int(ix[0])

    def describe(self, what):
        torch = self.torch
        src = ("the reference's unmodified mae_simsearch (utils/similarity.py:37-132, oracle/_ref) through a latent-cache stub"
               if self.kind == "reference" else "oracle/ref_port.search_loop (port of utils/similarity.py; oracle/_ref absent)")
        return (f"{src}, torch {torch.__version__} CPU fp32, {self.cores} threads: {what} over the first {self.rows} of "
                f"{self.wl['n']} bank rows (same synthetic tensors as the GPU arm), batch {self.batch}, one pass per query "
                f"as the reference requires, scaled linearly to the full bank")


def cpu_baseline_block(wl, dev):
    cpu = CpuReference(wl, dev)
    cpu.one_query(0, rows=min(cpu.rows, 20_000))                 # warm-up
    runs = [cpu.one_query(i) for i in range(3)]
    best = min(r[0] for r in runs)
    mean = sum(r[0] for r in runs) / len(runs)
    scale = wl["n"] / cpu.rows
    return {"value": 1.0 / (best * scale), "unit": "queries/s", "cores": cpu.cores, "kind": cpu.kind,
            "sample": cpu.describe(f"1 warm-up + best of 3 single-query passes ({best:.2f} s best, {mean:.2f} s mean)")}


def run_reference(args, name, wl):
    rank, local, world = dist_env()
    if rank != 0:
        return 0
    import torch
    dev = torch.device("cuda", local) if torch.cuda.is_available() else torch.device("cpu")
    cpu = CpuReference(wl, dev)
    # bounded sample: one step = one query over the sample rows; the whole run should end within ~2 minutes
    probe_rows = min(cpu.rows, 20_000)
    cpu.one_query(0, rows=probe_rows)
    t_probe, _ = cpu.one_query(0, rows=probe_rows)
    budget = 120.0 / max(args.steps + args.warmup, 1)
    rows = int(min(cpu.rows, max(4096, probe_rows * budget / max(t_probe, 1e-6))))
    cpu.rows = rows
    for i in range(args.warmup):
        cpu.one_query(i, rows=rows)
    times = [cpu.one_query(i, rows=rows)[0] for i in range(args.steps)]
    per_step = sum(times) / len(times)
    qps = 1.0 / (per_step * (wl["n"] / rows))
    cb = {"value": qps, "unit": "queries/s", "cores": cpu.cores, "kind": cpu.kind,
          "sample": cpu.describe(f"each step = 1 query ({args.steps} steps, {args.warmup} warm-up, mean {per_step:.2f} s)")}
    line = {"impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic" if dev.type == "cuda" else "synthetic (CPU generator: no CUDA device in this process)",
            "config": workload_config(name, wl, world, False), "cpu_baseline": cb,
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "the CPU has no shards to spread over: this is queries/s over ONE n-vector shard at every N"}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm: embedding-space workloads
# ------------------------------------------------------------------------------------------------
def pick_kernel(wl, path):
    dtype, Q, w = wl.get("dtype", "bf16"), wl["Q"], wl.get("weighted", False)
    L = wl.get("L", 1)
    if path == "batch" or (path == "auto" and dtype == "bf16" and L == 1 and not w and Q > 128):
        return "batch"
    if path == "tensor" or (path == "auto" and dtype == "bf16" and L == 1 and Q > 4):
        return "weighted" if w else "tensor"
    return "stream"


def run_search_workload(name, wl, args, dev, steps, warmup, strong, headline):
    """One embedding-space workload on this rank's GPU (all ranks call it together).  Returns the record (rank 0)."""
    import torch
    import torch.distributed as dist
    from sky_embeddings_b200 import _lib, synth
    from sky_embeddings_b200.distributed import make_exchange, shard_range
    from tests import torch_ref as TR

    rank, local, world = dist_env()
    n, D, Q, k, metric = wl["n"], wl["D"], wl["Q"], wl["k"], wl["metric"]
    dtype, L, weighted = wl.get("dtype", "bf16"), wl.get("L", 1), wl.get("weighted", False)
    esz = 2 if dtype == "bf16" else 4
    if strong:
        # equal shards: boundaries on bank tiles (128 rows), not on generator chunks -- with 65 536-row chunks rank 0 of 8
        # held 1 310 720 of C3's 10 M rows against 1 245 184 on the others and set the pace of every step
        row_lo, row_hi = shard_range(n, rank, world, align=128 * L)
        n_total = n
    else:
        chunks_per_rank = (n + synth.CHUNK_ROWS - 1) // synth.CHUNK_ROWS
        row_lo, row_hi = rank * chunks_per_rank * synth.CHUNK_ROWS, rank * chunks_per_rank * synth.CHUNK_ROWS + n
        n_total = n * world
    n_local = row_hi - row_lo
    if strong:
        bank = build_bank(n_local, D, dev, dtype=dtype, L=L, row0=row_lo, n_total=n)
    else:
        bank = build_bank(n_local, D, dev, chunk0=row_lo // synth.CHUNK_ROWS, dtype=dtype, L=L)
    # queries: planted neighbours of rows spread over the global bank (strong) / rank 0's shard (weak); identical on every rank
    t_dev, w_dev, planted = make_queries(n if (strong or world == 1) else n, D, Q, dev, weighted)
    t_host = t_dev.cpu().pin_memory()
    w_host = w_dev.cpu().pin_memory() if w_dev is not None else None
    out_s_host = torch.empty((Q, k), dtype=torch.float32).pin_memory()
    out_i_host = torch.empty((Q, k), dtype=torch.int64).pin_memory()
    # candidate exchange: peer-memory push + flag-waiting merge (csrc/exchange.cu), or one NCCL all-gather + merge kernel
    xchg = make_exchange(Q, k, dev, kind=args.exchange) if world > 1 else None
    largest = metric == "cosine"

    def local_search(t, w, out_s=None, out_i=None):
        return bank.search(t, w, k=k, metric=metric, path=args.path, idx_offset=row_lo, out_scores=out_s, out_idx=out_i)

    def step_device():
        if world == 1:
            return local_search(t_dev, w_dev)
        if args.exchange == "peer":
            # shard search whose merge kernel writes into every peer, then the flag-waiting merge (sky_search_sharded)
            return xchg.search(bank, t_dev, w_dev, metric=metric, path=args.path, idx_offset=row_lo)
        # local top-k straight into the exchange buffer, then one NCCL all-gather + merge
        local_search(t_dev, w_dev, xchg.scores, xchg.idx)
        return xchg.merge(metric)

    def step_host():
        # public API with HOST buffers: H2D of the queries and D2H of the results inside the call
        if world == 1:
            return bank.search_host(t_host, w_host, k=k, metric=metric, path=args.path, idx_offset=row_lo,
                                    out_scores=out_s_host, out_idx=out_i_host)
        td = t_host.to(dev, non_blocking=True)
        wd = w_host.to(dev, non_blocking=True) if w_host is not None else None
        if args.exchange == "peer":
            s, i = xchg.search(bank, td, wd, metric=metric, path=args.path, idx_offset=row_lo)
        else:
            local_search(td, wd, xchg.scores, xchg.idx)
            s, i = xchg.merge(metric)
        out_s_host.copy_(s, non_blocking=True)
        out_i_host.copy_(i, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out_s_host, out_i_host

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity gate before timing -------------------------------------------------------------------------
    s, i = step_device()
    torch.cuda.synchronize()
    s, i = s.clone(), i.clone()
    parity = {"checked_queries": 0, "k": k, "tolerance": None, "max_rel_err": None, "merge_bit_exact": None}
    if L == 1:
        if True:
            assert i[:, 0].cpu().tolist() == planted, f"{name}: planted nearest neighbours not returned: refusing to time a wrong kernel"
        sample = sorted(set(int(q) for q in torch.linspace(0, Q - 1, min(Q, 8)).round().tolist()))
        rel = 1e-3 if dtype == "bf16" else 1e-5          # north_star tolerances
        ws = w_dev[sample] if w_dev is not None else None
        ref_s, ref_i = TR.fp32_topk(bank, t_dev[sample], ws, metric, k, idx_offset=row_lo)
        if world > 1:
            gs = [torch.empty_like(ref_s) for _ in range(world)]
            gi = [torch.empty_like(ref_i) for _ in range(world)]
            dist.all_gather(gs, ref_s.contiguous())
            dist.all_gather(gi, ref_i.contiguous())
            ref_s, ref_i = TR.merge_sorted_lists(torch.stack(gs), torch.stack(gi), k, largest)
            # the exchanged + merged engine result vs a torch merge of the independently all-gathered local candidates
            ls, li = local_search(t_dev, w_dev)
            torch.cuda.synchronize()
            es = [torch.empty_like(ls) for _ in range(world)]
            ei = [torch.empty_like(li) for _ in range(world)]
            dist.all_gather(es, ls.contiguous())
            dist.all_gather(ei, li.contiguous())
            ms, mi = TR.merge_sorted_lists(torch.stack(es), torch.stack(ei), k, largest)
            exact = bool(torch.equal(ms.view(torch.int32), s.view(torch.int32)) and torch.equal(mi, i))
            assert exact, f"{name}: exchanged + merged top-k differs from the torch merge of the all-gathered candidates"
            parity["merge_bit_exact"] = True
        worst = 0.0
        # scale-relative tolerance: |got - ref| <= rel * max(|ref|, max |ref top-k score|)
        for j, q in enumerate(sample):
            ok, msg, err = TR.check_topk(s[q], i[q], ref_s[j], ref_i[j], rel, largest)
            assert ok, f"{name}: parity gate failed for query {q}: {msg}"
            worst = max(worst, err)
        parity.update(checked_queries=len(sample), tolerance=rel, max_rel_err=worst,
                      reference="fp32 torch scoring of the stored bank with the unrounded queries (tests/torch_ref.py), "
                                "full top-k, tie-aware; max_rel_err is relative to max(|score|, max |top-k score|)")

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(warmup, 3)):
        step_device()
    barrier()
    bank.profile(True)
    bank.profile_read(reset=True)
    _lib.launch_count(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    ev0.record()
    for _ in range(steps):
        step_device()
    ev1.record()
    barrier()
    t_wall1 = time.perf_counter()
    launches = _lib.launch_count(reset=True)
    ms_total = ev0.elapsed_time(ev1)
    n_kern, kern_ms = bank.profile_read(reset=True)
    bank.profile(False)
    kernel_ms_ranks = None
    if world > 1:
        tt = torch.tensor([ms_total], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
        # the scoring kernel's mean duration on every rank: a synchronous exchange makes each step as slow as the slowest
        kk = torch.tensor([kern_ms / max(n_kern, 1) * (n_kern / steps)], device=dev)
        allk = [torch.empty_like(kk) for _ in range(world)]
        dist.all_gather(allk, kk)
        kernel_ms_ranks = [round(float(v.item()), 4) for v in allk]
    ms_step = ms_total / steps

    # end-to-end through the host-buffer API
    for _ in range(3):
        step_host()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step_host()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item())
    e2e_step = e2e_ms / steps
    # keep the GPU busy with the same work long enough for the clock sampler if the run was short
    if (t_wall1 - t_wall0) < 0.05:
        t_end = time.perf_counter() + 0.3
        while time.perf_counter() < t_end:
            step_device()
        torch.cuda.synchronize()
    sampler.stop()
    sampler.join(timeout=1.0)

    rec = None
    if rank == 0:
        peaks = read_peaks()
        kernel = pick_kernel(wl, args.path)
        agg = Q / (ms_step * 1e-3)
        value = agg if (strong or world == 1) else world * agg
        rec = {"metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": steps,
               "warmup": max(warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
               "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
               "config": workload_config(name, wl, world, strong),
               "exchange": (args.exchange if world > 1 else None), "kernel_ms_per_rank": kernel_ms_ranks,
               "value_definition": "Q / step time over the global (row-sharded) bank" if (strong or world == 1) else
                                   "N_gpus * Q / step time (query-over-1M-vector-shard searches/s, weak scaling)",
               "qps_global_bank": agg, "parity": parity, "clocks": sampler.summary((t_wall0, t_wall1)),
               "e2e": {"value": (Q if (strong or world == 1) else world * Q) / (e2e_step * 1e-3), "unit": "queries/s",
                       "ms_per_step": e2e_step, "h2d_bytes_per_step": t_host.numel() * 4 * (2 if weighted else 1),
                       "d2h_bytes_per_step": Q * k * 12,
                       "api": "sky_search_host (C ABI, pinned host buffers: queries read and results written over PCIe by the kernels themselves, inside the timed call)" if world == 1 else
                              "pinned H2D + sky_search + candidate exchange + merge + D2H"},
               "gpu_launches": launches, "gpu_launches_per_step": launches / steps, "env_knobs": active_knobs()}
        if kernel == "batch":
            # tensor-pipe bound: all phase launches of one search together; flops = 2 Q N D of THIS GPU's shard
            kern_search_ms = kern_ms / steps
            flops = 2.0 * Q * float(n_local) * D
            ach = flops / (kern_search_ms * 1e-3) / 1e12
            rec["roofline"] = {"bound": "tensor", "achieved": ach, "peak": peaks["tf"], "unit": "TFLOP/s", "frac": ach / peaks["tf"],
                               "traffic": None, "kernel": "tc_batch_kernel (all phases of one search)", "kernel_ms": kern_search_ms,
                               "kernel_launches": n_kern, "algorithmic_flops": flops, "peak_source": peaks["tf_src"],
                               "kernel_share_of_step": kern_search_ms / ms_step,
                               "step_frac": flops / (ms_step * 1e-3) / 1e12 / peaks["tf"]}
        else:
            kern_avg_ms = kern_ms / max(n_kern, 1)
            algo_bytes = float(n_local) * D * esz                    # one pass over this GPU's shard
            achieved = algo_bytes / (kern_avg_ms * 1e-3) / 1e9 if kern_avg_ms > 0 else 0.0
            kname = {"tensor": "tc_search_kernel<64>", "weighted": "tc_weighted_kernel", "stream": "stream_search_kernel"}[kernel]
            tkey = {"tensor": "tensor", "weighted": "weighted", "stream": "stream_fp32_q1" if dtype == "fp32" else "stream_bf16"}[kernel]
            rec["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm"], "unit": "GB/s",
                               "frac": achieved / peaks["hbm"], "traffic": read_traffic(tkey) if n == 1_000_000 else None,
                               "traffic_source": "committed ncu --set full capture of the same workload (profiles/), not measured in this run",
                               "kernel": kname, "kernel_ms": kern_avg_ms, "kernel_launches": n_kern,
                               "algorithmic_bytes": algo_bytes, "peak_source": peaks["hbm_src"],
                               "kernel_share_of_step": kern_avg_ms * (n_kern / steps) / ms_step,
                               "step_frac": algo_bytes * (n_kern / steps) / (ms_step * 1e-3) / 1e9 / peaks["hbm"]}
        if headline and world == 1 and not args.no_cpu:
            rec["cpu_baseline"] = cpu_baseline_block(wl, dev)
    bank.close()
    del bank
    torch.cuda.empty_cache()
    return rec


# ------------------------------------------------------------------------------------------------
# pixel-space masked-MSE workloads (BASELINE config 5): raw 5 x 64 x 64 fp32 cutouts, HBM bound
# ------------------------------------------------------------------------------------------------
def pixel_chunk(dev, rows, chunk_index, seed=20240607):
    """Device-generated synthetic cutouts ~ N(0,1) clipped at -3 with 2% NaN pixels and 5% missing bands
    (same statistics as sky_embeddings_b200.synth.cutouts; row content depends only on (seed, chunk))."""
    import torch
    g = torch.Generator(device=dev).manual_seed(seed * 1000003 + chunk_index)
    x = torch.randn((rows, 5, 64, 64), generator=g, device=dev).clamp_(min=-3.0)
    x[torch.rand((rows, 5, 64, 64), generator=g, device=dev) < 0.02] = float("nan")
    x[torch.rand((rows, 5), generator=g, device=dev) < 0.05] = float("nan")
    return x


def run_pixel_workload(name, args, dev, steps, warmup):
    import numpy as np
    import torch
    from sky_embeddings_b200 import PixelBank, _lib
    from tests import torch_ref as TR
    rank, local, world = dist_env()
    n, Q, k = PIXEL_WORKLOADS[name]
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

    s, i = step_device()
    torch.cuda.synchronize()
    assert i[:, 0].cpu().tolist() == planted, "planted nearest cutouts not returned: refusing to time a wrong kernel"
    # parity gate: full top-k of (up to) 2 queries against an fp32 torch scoring of all stored cutouts
    worst = 0.0
    checked = list(range(min(Q, 2)))
    for q in checked:
        best_s = torch.full((k,), float("inf"), device=dev)
        best_i = torch.full((k,), -1, dtype=torch.int64, device=dev)
        for c, s0 in enumerate(range(0, n, chunk)):
            x = pixel_chunk(dev, min(chunk, n - s0), c).reshape(-1, D)
            sc = TR.pixel_scores(x, q_dev[q].reshape(D))
            cs = torch.cat([best_s, sc])
            ci = torch.cat([best_i, torch.arange(s0, s0 + x.shape[0], device=dev)])
            best_s, o = cs.topk(k, largest=False)
            best_i = ci[o]
        ok, msg, err = TR.check_topk(s[q], i[q], best_s, best_i, 1e-5, False)
        assert ok, f"{name}: parity gate failed for query {q}: {msg}"
        worst = max(worst, err)
    parity = {"checked_queries": len(checked), "k": k, "tolerance": 1e-5, "max_rel_err": worst,
              "reference": "fp32 torch masked-MSE over all stored cutouts (tests/torch_ref.py), full top-k, tie-aware"}
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(warmup, 3)):
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
    t1 = time.perf_counter()
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
    peaks = read_peaks()
    kern_avg = kern_ms / max(n_kern, 1)
    algo = float(n) * D * 4
    achieved = algo / (kern_avg * 1e-3) / 1e9
    rec = {"metric": METRIC, "value": Q / (ms_step * 1e-3), "unit": "queries/s", "n_gpus": 1, "steps": steps,
           "warmup": max(warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"{name}: {n} cutouts 5x64x64 fp32 ({algo / 1e9:.1f} GB), {Q} queries, "
                                  f"pixel-space NaN-aware masked MSE, top-{k}, exact",
                      "bank_vectors": n, "dim": D, "queries": Q, "k": k, "similarity": "masked MSE (pixels)",
                      "l2_policy": "bank is larger than the 126 MB L2; no flush needed"},
           "parity": parity, "clocks": sampler.summary((t0, t1)),
           "e2e": {"value": Q / (e2e_step * 1e-3), "unit": "queries/s", "ms_per_step": e2e_step,
                   "h2d_bytes_per_step": Q * D * 4, "d2h_bytes_per_step": Q * k * 12,
                   "api": "PixelBank.search with pinned host queries / results"},
           "gpu_launches": launches, "gpu_launches_per_step": launches / steps, "env_knobs": active_knobs(),
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm"], "unit": "GB/s", "frac": achieved / peaks["hbm"],
                        "traffic": read_traffic("pixel_q1") if (n == 1_000_000 and Q == 1) else None,
                        "traffic_source": "committed ncu --set full capture (profiles/), not measured in this run",
                        "kernel": "pixel_search_kernel", "kernel_ms": kern_avg, "kernel_launches": n_kern,
                        "algorithmic_bytes": algo, "peak_source": peaks["hbm_src"],
                        "kernel_share_of_step": kern_avg * (n_kern / steps) / ms_step,
                        "note": "peak is the copy-measured (read+write) figure; a read-only stream can exceed it"}}
    if not args.no_cpu:
        # no reference implementation of a pixel-space search exists: the numpy restatement
        from oracle import sky_oracle as O
        rows = 2000
        xs = pixel_chunk(dev, rows, 0).cpu().numpy()
        qs = q_dev[0].cpu().numpy()
        O.pixel_masked_mse(qs, xs[:200])
        tt0 = time.perf_counter()
        O.pixel_masked_mse(qs, xs, dtype=np.float32)
        dt = time.perf_counter() - tt0
        rec["cpu_baseline"] = {"value": 1.0 / (dt * n / rows), "unit": "queries/s", "cores": 1, "kind": "port",
                               "sample": f"oracle/sky_oracle.pixel_masked_mse (numpy fp32) over {rows} of {n} cutouts, "
                                         f"1 query, {dt:.2f} s, scaled linearly"}
    bank.close()
    del bank
    torch.cuda.empty_cache()
    return rec


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from sky_embeddings_b200 import _lib
    rank, local, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the search path has no CPU fallback")
    knobs = active_knobs()
    if knobs and not args.allow_knobs:
        raise SystemExit(f"experiment knobs are set in the environment ({', '.join(knobs)}): refusing to produce a bench "
                         "line (they exist only in -DSKY_EXPERIMENTS builds; pass --allow-knobs for kernel experiments)")
    _lib.load()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    name = args.workload
    if name in PIXEL_WORKLOADS:
        if world > 1:
            raise SystemExit("pixel workloads are single-GPU bench lines")
        line = run_pixel_workload(name, args, dev, args.steps, args.warmup)
    else:
        wl = WORKLOADS[name]
        strong = args.strong or (name in ("c4",) and world > 1)
        line = run_search_workload(name, wl, args, dev, args.steps, args.warmup, strong, headline=True)
    also = {}
    names = [] if args.also == "none" else ((ALSO_N1 if world == 1 else ALSO_MULTI) if args.also == "default" else args.also.split(","))
    if args.workload != "c2" and args.also == "default":
        names = []
    for nm in names:
        try:
            if nm in PIXEL_WORKLOADS:
                rec = run_pixel_workload(nm, args, dev, 30, 3) if world == 1 else None
            else:
                wl = WORKLOADS[nm]
                rec = run_search_workload(nm, wl, args, dev, wl.get("steps", 100), 3, strong=world > 1, headline=False)
        except AssertionError:
            raise                                            # a parity gate failed: no bench line at all
        except torch.cuda.OutOfMemoryError as e:
            free_b, total_b = torch.cuda.mem_get_info(dev)
            rec = {"skipped": f"out of device memory: {str(e)[:120]}",
                   "memory": {"free_GB": round(free_b / 1e9, 2), "total_GB": round(total_b / 1e9, 2),
                              "torch_allocated_GB": round(torch.cuda.memory_allocated(dev) / 1e9, 2),
                              "torch_reserved_GB": round(torch.cuda.memory_reserved(dev) / 1e9, 2)}}
        except _lib.SkyError as e:
            if e.code != -4:                                 # SKY_ERR_NOMEM
                raise
            rec = {"skipped": f"out of device memory: {str(e)[:120]}"}
        if rank == 0:
            also[nm] = rec
    if rank == 0:
        line["also"] = also
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + sorted(PIXEL_WORKLOADS))
    ap.add_argument("--also", default="default", help="'default', 'none' or a comma list of workloads reported as sub-records")
    ap.add_argument("--exchange", choices=["peer", "nccl"], default="peer",
                    help="N > 1: candidate exchange over peer memory (default) or one NCCL all-gather")
    ap.add_argument("--strong", action="store_true", help="N > 1: shard the workload's bank over the ranks (strong scaling)")
    ap.add_argument("--path", default="auto", choices=["auto", "simt", "tensor", "batch", "generic"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline legs")
    ap.add_argument("--allow-knobs", action="store_true", help="kernel experiments with SKY_* knobs (not a bench line)")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.workload in PIXEL_WORKLOADS:
            raise SystemExit("--impl reference is defined for the embedding-space workloads")
        return run_reference(args, args.workload, WORKLOADS[args.workload])
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
