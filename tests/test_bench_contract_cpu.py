"""bench.py's reference arm runs without a GPU: check the stdout contract (exactly one JSON line) and the keys the
driver reads.  (The GPU arm's line is the same dictionary plus roofline / clocks / gpu_launches; it is exercised on the
GPU box by tools/gpu_final.sh.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--no-render"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, f"stdout must carry exactly one line, got {len(lines)}"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["metric"].startswith("voxel*views/s") and d["unit"] == "voxel*views/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_our_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, "bench.py", "--steps", "1", "--warmup", "0"], cwd=ROOT, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_committed_k1_traffic_capture_matches_the_kernel_sources():
    """roofline.traffic comes from profiles/r02_k1_256_traffic.json (tools/summarize_ncu.py traffic <.ncu-rep>), which
    records the hash of the kernel sources the capture was taken from: a capture that is stale against the tree fails
    here (and bench.py would print traffic = null)."""
    import json
    sys.path.insert(0, ROOT)
    import bench
    assert os.path.exists(bench.K1_TRAFFIC_PROFILE), "no committed ncu traffic capture of K1"
    rec = json.load(open(bench.K1_TRAFFIC_PROFILE))
    assert rec["kernel_source_sha256_16"] == bench.k1_source_hash(), "K1 sources changed since the ncu capture: re-capture"
    assert rec["dram_bytes_read"] > 0 and rec["dram_bytes_write"] > 0.5 * 256 ** 3 * 36
