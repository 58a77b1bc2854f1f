"""A/B of experiment knobs in ONE process per bank shape (needs a -DSKY_EXPERIMENTS build of libskysearch.so: the
release library never reads the environment).  Every setting must return the baseline setting's top-k bit for bit.
usage: ab_knobs.py <case> [<case> ...]   cases: c2 c2w q1w q1wb c3g8 c4g8 c3"""
import os, sys, time
import torch
sys.path.insert(0, ".")
from sky_embeddings_b200 import Bank

CASES = {
    # name: (n, Q, k, metric, weighted, dtype, path, steps, [settings])
    "c2": (1_000_000, 64, 100, "cosine", False, "bf16", "auto", 200, [{}, {"SKY_PDL": "0"}, {}]),
    "c2w": (1_000_000, 64, 100, "cosine", True, "bf16", "auto", 200, [{}, {"SKY_PDL": "0"}]),
    "q1w": (1_000_000, 1, 100, "cosine", True, "fp32", "auto", 200, [{}, {"SKY_PDL": "0"}]),
    "q1wb": (1_000_000, 1, 100, "cosine", True, "bf16", "auto", 200, [{}, {"SKY_PDL": "0"}]),
    "c3g8": (1_250_000, 4096, 100, "MSE", False, "bf16", "batch", 10,
             [{"SKY_TB_DENSE0": "148"}, {}, {"SKY_TB_GROWTH": "8"}, {"SKY_TB_DENSE0": "148", "SKY_TB_GROWTH": "8"},
              {"SKY_TB_DENSE0": "16"}, {"SKY_TB_DENSE0": "62"}]),
    "c4g8": (12_500_000, 1000, 1000, "cosine", False, "bf16", "batch", 4,
             [{"SKY_TB_SORTED": "1"}, {}, {"SKY_TB_GROWTH": "8"}, {"SKY_TB_GROWTH": "6"}]),
    # release-library runs (knobs are ignored there): the same setting twice shows the run-to-run noise
    "c3g8r": (1_250_000, 4096, 100, "MSE", False, "bf16", "batch", 20, [{}, {}]),
    "c4g8r": (12_500_000, 1000, 1000, "cosine", False, "bf16", "batch", 6, [{}, {}]),
    "c3r": (10_000_000, 4096, 100, "MSE", False, "bf16", "batch", 4, [{}, {}]),
    "c2r": (1_000_000, 64, 100, "cosine", False, "bf16", "auto", 300, [{}, {}]),
    "c3": (10_000_000, 4096, 100, "MSE", False, "bf16", "batch", 3,
           [{"SKY_TB_DENSE0": "148"}, {}, {"SKY_TB_GROWTH": "8"}]),
}
KEYS = ("SKY_PDL", "SKY_TB_DENSE0", "SKY_TB_GROWTH", "SKY_TB_SORTED")


def run(name):
    n, Q, k, metric, weighted, dtype, path, steps, settings = CASES[name]
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev); g.manual_seed(11)
    D = 768
    bank = Bank(n, 1, D, dtype=dtype, device=dev)
    chunk = 1 << 18
    for r0 in range(0, n, chunk):
        m = min(chunk, n - r0)
        x = torch.randn(m, 1, D, device=dev, generator=g)
        if r0 == 0:
            bank.fit_norm(x[:4096])
        bank.upload(x, item0=r0)
    bank.finalize()
    t = torch.randn(Q, D, device=dev, generator=g)
    w = None
    if weighted:
        w = torch.rand(Q, D, device=dev, generator=g) + 0.1
        w = w / w.sum(1, keepdim=True)
    t_host = t.cpu().pin_memory()
    w_host = w.cpu().pin_memory() if w is not None else None
    so = torch.empty((Q, k), dtype=torch.float32).pin_memory()
    io = torch.empty((Q, k), dtype=torch.int64).pin_memory()
    base = None
    for st in settings:
        for key in KEYS:
            os.environ.pop(key, None)
        os.environ.update(st)
        for _ in range(3):
            s, i = bank.search(t, w, k=k, metric=metric, path=path)
        torch.cuda.synchronize()
        if base is None:
            base = (s.clone(), i.clone())
            same = "baseline"
        else:
            same = "bit-identical" if (torch.equal(base[0].view(torch.int32), s.view(torch.int32)) and torch.equal(base[1], i)) else "DIFFERENT"
        # device-resident steps, clean (no profile events between the kernels)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            bank.search(t, w, k=k, metric=metric, path=path)
        e1.record(); torch.cuda.synchronize()
        step = e0.elapsed_time(e1) / steps
        # scoring-kernel time (profile events on)
        bank.profile(True); bank.profile_read(reset=True)
        for _ in range(max(2, steps // 4)):
            bank.search(t, w, k=k, metric=metric, path=path)
        torch.cuda.synchronize()
        nl, ms = bank.profile_read(reset=True)
        bank.profile(False)
        kern = ms / max(2, steps // 4)
        # host-buffer steps (the e2e leg), one synchronise per step as a caller would
        for _ in range(3):
            bank.search_host(t_host, w_host, k=k, metric=metric, path=path, out_scores=so, out_idx=io)
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        hs = min(steps, 100)
        for _ in range(hs):
            bank.search_host(t_host, w_host, k=k, metric=metric, path=path, out_scores=so, out_idx=io)
            torch.cuda.current_stream().synchronize()
        host = (time.perf_counter() - t0) / hs * 1e3
        print(f"{name:5s} {str(st):60s} step {step:8.4f} ms  scoring kernels {kern:8.4f} ms ({nl / max(2, steps // 4):.0f}/search)  "
              f"host-buffer step {host:8.4f} ms  {same}", flush=True)
    bank.close()


if __name__ == "__main__":
    for c in sys.argv[1:]:
        run(c)
        torch.cuda.empty_cache()
