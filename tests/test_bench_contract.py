"""The bench.py JSON-line contract, checked on the committed result lines (no GPU needed): every key
the driver reads is present and consistent with the others."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as fh:
        return json.loads(fh.read().strip().splitlines()[-1])


@pytest.mark.parametrize("name,n", [("bench_r1_n1_final.json", 1), ("bench_r1_n2_final.json", 2), ("bench_r1_n8_final.json", 8)])
def test_committed_bench_lines_follow_the_contract(name, n):
    d = _line(name)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks"):
        assert key in d, key
    assert d["n_gpus"] == n and d["unit"] == "Gnnz/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in r, key
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    # value is whole-job throughput: all ranks' input elements per step over the (max over ranks) step time
    nnz = d["config"]["nnz_a"] + d["config"]["nnz_b"]
    assert abs(d["value"] - n * nnz / (d["ms_per_step"] * 1e-3) / 1e9) / d["value"] < 0.02
    assert d["gpu_launches"] >= d["steps"]           # at least one of our kernels per step
    assert d["clocks"]["reasons"] == [] and d["clocks"]["sm_mhz"] > 0.9 * d["clocks"]["sm_max_mhz"]
    if n == 1:
        c = d["cpu_baseline"]
        for key in ("value", "unit", "cores", "kind", "sample"):
            assert key in c, key
        assert c["kind"] in ("reference", "port") and c["cores"] >= 1
    else:
        assert "sharded_exchange" in d
