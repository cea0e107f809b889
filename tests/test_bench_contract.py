"""The bench JSON contract, checked on the committed lines of the last measured run (profiles/r1_bench_lines.jsonl):
a renamed or dropped key in bench.py shows up here before it reaches the driver."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def lines():
    with open(os.path.join(ROOT, "profiles", "r1_bench_lines.jsonl")) as f:
        return [json.loads(l) for l in f if l.strip()]


def test_b200_lines_carry_the_contract_keys():
    mine = [d for d in lines() if d.get("impl") != "reference"]
    assert mine and mine[0]["config"]["workload"].startswith("hmc_cfg2")
    for d in mine:
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "roofline", "gpu_launches", "clocks"):
            assert k in d, k
        assert d["metric"] == "chain-steps/sec" and d["higher_is_better"] is True and d["scaling"] == "weak"
        assert d["vs_baseline"] is None and d["data"].startswith("synthetic") and "workload" in d["config"]
        assert d["warmup"] >= 3 and d["gpu_launches"] >= 1
        r = d["roofline"]
        for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
            assert k in r, k
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    head = mine[0]
    assert head["roofline"]["bound"] == "tensor" and head["roofline"]["engine"] == "tcgen05"
    assert head["roofline"]["traffic"] and head["n_gpus"] == 1
    e = head["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] == 256 * 200 * 751 * 4 and e["value"] < head["value"]
    c = head["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]


def test_reference_arm_line():
    ref = [d for d in lines() if d.get("impl") == "reference"]
    assert ref
    d = ref[0]
    assert d["metric"] == "chain-steps/sec" and d["config"]["workload"].startswith("hmc_cfg2")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"] == d["cpu_baseline"]["value"]


def test_weak_scaling_lines():
    by_n = {d["n_gpus"]: d["value"] for d in lines() if d.get("impl") != "reference" and
            d["config"]["workload"].startswith("hmc_cfg2") and d["roofline"].get("engine") == "tcgen05"}
    assert {1, 2, 8} <= set(by_n)
    assert by_n[2] > 1.9 * by_n[1] and by_n[8] > 7.5 * by_n[1]
