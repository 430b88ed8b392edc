"""The JSON line bench.py prints (driver contract): the newest committed line under profiles/ has every key the
contract names, with consistent values.  (bench.py itself needs a GPU; this guards the schema on CPU.)"""
import glob
import json
import os

from tests import util


def _newest():
    files = sorted(glob.glob(os.path.join(util.ROOT, "profiles", "r*_bench.json")))
    assert files, "no bench line committed under profiles/"
    return json.load(open(files[-1])), files[-1]


def test_bench_line_has_the_contract_keys():
    d, path = _newest()
    for k in ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks", "cpu_baseline"]:
        assert k in d, f"{path}: missing {k}"
    assert d["metric"] == "crops_per_second" and d["unit"] == "crops/s" and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["data"] == "synthetic" and d["dtype"] == "f32" and d["warmup"] >= 3
    assert "workload" in d["config"] and "model" not in d["config"]
    e = d["e2e"]
    assert e["unit"] == "crops/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
    assert "traffic" in r
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    # value = crops of the timed region / its duration; launches: at least one, at most one per frame
    crops = d["config"]["crops_per_step"] * d["steps"]
    assert abs(d["value"] - crops / (d["ms_per_step"] * d["steps"] * 1e-3)) / d["value"] < 1e-6
    assert 0 < d["gpu_launches"] <= d["config"]["frames_per_step"] * d["steps"]
