"""The JSON lines bench.py printed on the B200 (kept under profiles/) carry every key the driver's
contract names; and the reference arm runs here, on the CPU, end to end on a tiny sample."""
import json
import os
import subprocess
import sys

import pytest

import util

PROFILES = os.path.join(util.ROOT, "profiles")
BASE_KEYS = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"]


@pytest.mark.parametrize("name", ["r01_bench_n1.json", "r01_bench_n2.json", "r01_bench_n4.json", "r01_bench_n8.json",
                                  "r02_bench_n1.json", "r02_bench_n2.json", "r02_bench_n8.json"])
def test_recorded_bench_lines_follow_the_contract(name):
    d = json.load(open(os.path.join(PROFILES, name)))
    for k in BASE_KEYS + ["roofline", "clocks"]:
        assert k in d, k
    assert d["unit"] == "Gbit/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["dtype"] == "u8"
    assert d["vs_baseline"] is None and "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.5 < r["frac"] < 1.0
    assert abs(d["value"] - d["config"]["total_symbols"] / (d["ms_per_step"] / 1e3) / 1e9) < 1e-6 * d["value"]
    assert d["gpu_launches"] > 0 and d["warmup"] >= 3
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if d["n_gpus"] == 1:
        c = d["cpu_baseline"]
        for k in ("value", "unit", "cores", "kind", "sample"):
            assert k in c, k
        e = d["e2e"]
        assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    if name.startswith("r02"):
        # round 2: the roofline is the kernel as launched inside the timed steps, and it fits into a step
        assert r["kernel_ms"] <= d["ms_per_step"] and "inside the timed steps" in r["kernel"]
        if d["n_gpus"] == 1:
            assert {x["cores"] for x in d["cpu_baseline"]["rows"]} >= {1} and len(d["cpu_baseline"]["rows"]) == 3
            assert set(d["e2e"]["routes"]) == {"packed", "bytes"}
            legs = d["chain"]["legs"]
            assert legs["try_clocks"]["kernel_packets_per_s"] > 10 * 3.5e6          # >= 10x the 3.5 M packets/s of round 1
            assert legs["decode"]["rv_histogram"]["10"] > 100000                    # the real decode leg, not the reject path
            assert 0.9 < d["known_lap"]["roofline"]["frac"] < 1.05
        else:
            s = d["strong"]
            assert s["scaling"] == "strong" and s["symbols_per_gpu"] * d["n_gpus"] == d["config"]["symbols_per_gpu"]


def test_recorded_sweep_matches_the_reference_at_1_and_8_gpus():
    for n in (1, 8):
        rows = [json.loads(l) for l in open(os.path.join(PROFILES, f"r02_sweep_config5_n{n}.json"))]
        cells = [r for r in rows if "k" in r]
        assert len(cells) == 30 and all(c["matches_cpu"] and c["n_gpus"] == n and c["cpu_kind"] == "reference" for c in cells)
        assert {c["k"] for c in cells} == {0, 1, 2, 3, 4} and rows[-1] == {"summary": "all_match", "value": True, "cells": 30}
    # detection-rate curves coincide between 1 and 8 GPUs (same lists)
    a = {(c["k"], c["ber"]): c["detection_rate"] for c in (json.loads(l) for l in open(os.path.join(PROFILES, "r02_sweep_config5_n1.json"))) if "k" in c}
    b = {(c["k"], c["ber"]): c["detection_rate"] for c in (json.loads(l) for l in open(os.path.join(PROFILES, "r02_sweep_config5_n8.json"))) if "k" in c}
    assert a == b


@pytest.mark.parametrize("name", ["r01_bench_reference_arm.json", "r02_bench_reference_arm.json"])
def test_recorded_reference_arm_line(name):
    d = json.load(open(os.path.join(PROFILES, name)))
    for k in BASE_KEYS + ["impl", "cpu_baseline"]:
        assert k in d, k
    assert d["impl"] == "reference" and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"]
    if name.startswith("r02"):
        assert d["product_library_loaded"] is False
