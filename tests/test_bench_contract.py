"""CPU: the bench.py contract that needs no GPU -- the reference arm prints one well-formed JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "small",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["metric"] == "queries/sec exact top-k over N-vector bank"
    assert d["value"] > 0 and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_bench_refuses_every_experiment_knob_the_library_reads():
    """bench.py refuses to print a bench line while an experiment knob is set; its list must name every
    env_knob("SKY_...") the sources read (they are only compiled in with -DSKY_EXPERIMENTS)."""
    import glob
    import re
    sys.path.insert(0, ROOT)
    import bench
    names = set()
    for f in glob.glob(os.path.join(ROOT, "sky_embeddings_b200", "csrc", "*.cu*")):
        names |= set(re.findall(r'env_knob\("(SKY_[A-Z0-9_]+)"', open(f).read()))
    assert names, "no env_knob call found: the scan is broken"
    assert names <= set(bench.KNOBS), f"knobs missing from bench.KNOBS: {sorted(names - set(bench.KNOBS))}"


def test_strong_scaling_shards_are_equal_and_hold_the_same_rows():
    """bench.py's strong-scaled banks: tile-aligned shards of equal size, assembled from generator chunks of canonical
    size, so every rank holds exactly the rows of the global synthetic bank (checked on the CPU generator)."""
    import torch
    sys.path.insert(0, ROOT)
    import bench
    from sky_embeddings_b200 import synth
    from sky_embeddings_b200.distributed import shard_range
    C = synth.CHUNK_ROWS
    # shapes of the driver's runs: C3 (10 M) and C4 (100 M) over 2 / 4 / 8 ranks
    for n in (10_000_000, 100_000_000):
        for world in (2, 4, 8):
            sizes, nxt = [], 0
            for rank in range(world):
                lo, hi = shard_range(n, rank, world, align=128)
                assert lo == nxt and lo % 128 == 0
                nxt = hi
                sizes.append(hi - lo)
                pieces = bench.shard_pieces(lo, hi - lo, n)
                assert sum(p[3] for p in pieces) == hi - lo and pieces[0][4] == 0
                for c, chunk_rows, off, take, dst in pieces:
                    assert chunk_rows == min(C, n - c * C) and 0 <= off and off + take <= chunk_rows
                    assert c * C + off == lo + dst and dst % 128 == 0
            assert nxt == n and max(sizes) - min(sizes) <= 128 + n % 128
    # data: a small bank cut at awkward places equals the chunk-wise global bank
    n, D = 2 * C + 12_800, 8
    glob = torch.cat([bench.raw_chunk(c, min(C, n - c * C), D, "cpu") for c in range(3)])
    for lo, hi in ((0, n), (128, C + 256), (C - 128, n), (C, 2 * C), (2 * C + 128, n)):
        got = torch.cat([bench.raw_chunk(c, cr, D, "cpu")[off:off + take] for c, cr, off, take, _ in bench.shard_pieces(lo, hi - lo, n)])
        assert torch.equal(got, glob[lo:hi])
