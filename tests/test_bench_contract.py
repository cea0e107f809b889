"""The bench JSON contract, checked on the committed lines of the last measured run (profiles/r2_bench_lines.jsonl):
a renamed or dropped key in bench.py shows up here before it reaches the driver."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def lines():
    with open(os.path.join(ROOT, "profiles", "r2_bench_lines.jsonl")) as f:
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
    assert head["roofline"]["bound"] == "tensor" and head["roofline"]["engine"].startswith("tcgen05")
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
    """hmc_cfg2 is weak-scaled with no data-path collective; the committed lines come from different boxes of the pool
    (+-5 % box to box on the same build), so the single-GPU yard-stick is the slowest N = 1 line."""
    mine = [d for d in lines() if d.get("impl") != "reference" and d["config"]["workload"].startswith("hmc_cfg2") and
            str(d["roofline"].get("engine", "")).startswith("tcgen05")]
    by_n = {}
    for d in mine:
        by_n.setdefault(d["n_gpus"], []).append(d["value"])
    assert {1, 2, 8} <= set(by_n)
    one = min(by_n[1])
    assert max(by_n[2]) > 1.9 * one and max(by_n[8]) > 7.5 * one


def test_round2_additions():
    """SGLD next to HMC in the driver-run line, e2e through the reference-signature API, step size in config, and -- at
    N > 1 -- BASELINE configs[4] sharded with the on-device bit-identity check and the warm moments all_gather."""
    mine = [d for d in lines() if d.get("impl") != "reference"]
    head = mine[0]
    assert "step_size" in head["config"] and "step_size=" in head["config"]["workload"]
    assert "pbnn_b200.mcmc.hmc" in head["e2e"]["api"] and head["e2e"]["ops_value"] > 0
    for name in ("sgld_cfg1x", "sgld_cfg3", "sgld_cfg5"):
        w = head["workloads"][name]
        assert w["value"] > 0 and w["chains_finite"] and w["roofline"]["frac"] > 0 and w["predictive"]["ms"] > 0
        assert abs(w["roofline"]["frac"] - w["roofline"]["achieved"] / w["roofline"]["peak"]) < 1e-9
    assert head["workloads"]["sgld_cfg5"]["roofline"]["engine"] == "tcgen05_bf16x3_mlp2"
    multi = [d for d in mine if d["n_gpus"] > 1]
    assert multi
    for d in multi:
        s = d["sharded_cfg5"]
        assert s["sharding_check"]["bit_identical_to_solo_run"] is True and s["sharding_check"]["ranks"] == d["n_gpus"]
        assert 0 < s["moments_allgather_us"] < 1000 and s["value"] > 0 and s["scaling"] == "weak"
    assert any(d["n_gpus"] == 8 and "8192 SGLD chains" in d["sharded_cfg5"]["what"] for d in multi)
