"""bench.py's reference arm (the CPU restatement of the path on the host cores) runs without a GPU and prints ONE JSON
line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "canonical_neighborhoods_per_sec" and d["unit"] == "neighborhoods/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] >= 1
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_gather_roofline_arithmetic():
    """bench.gather_roofline: bytes of the gather's own formulation per rank over its time, against the measured HBM peak."""
    import bench

    r = bench.gather_roofline(nodes=1_000_000, directed_edges=22_000_000, queries=29, world=1, gather_ms=11.8)
    alg = 20.0 * 22_000_000 * 29 + 520.0 * 1_000_000 * 29
    assert r["algorithmic_bytes_per_step_per_rank"] == alg and r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["achieved"] - alg / 11.8e-3 / 1e9) < 1e-6 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert 0.2 < r["frac"] < 0.6
    r8 = bench.gather_roofline(10_000_000, 220_000_000, 29, 8, 17.4)
    assert abs(r8["algorithmic_bytes_per_step_per_rank"] - (20.0 * 220e6 * 29 + 520.0 * 10e6 * 29) / 8) < 1.0
    assert bench.gather_roofline(1, 1, 1, 1, 0.0)["achieved"] == 0.0


def test_pack_int32_block_layout():
    """pipeline.pack_int32_block: parts back to back, each on an aligned offset, contents intact, int32 only."""
    import numpy as np
    import pytest

    from desco_b200.pipeline import pack_int32_block

    arrs = [np.arange(5, dtype=np.int32), np.arange(300, dtype=np.int32) * 3, np.zeros(0, dtype=np.int32), np.array([7], dtype=np.int32)]
    block, parts = pack_int32_block(arrs, align=128)
    assert parts == [(0, 5), (128, 300), (512, 0), (512, 1)] and block.numel() == 640
    for (o, n), a in zip(parts, arrs):
        assert np.array_equal(block[o:o + n].numpy(), a)
    with pytest.raises(ValueError):
        pack_int32_block([np.arange(3, dtype=np.int64)])
