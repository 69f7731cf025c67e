"""The bench.py JSON contract, checked on the lines committed under profiles/ (they were printed by bench.py on the
B200 box): every key the driver and the judge read must be there, with the right shape."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not committed yet")
    for raw in open(path):
        raw = raw.strip()
        if raw.startswith("{"):
            return json.loads(raw)
    raise AssertionError(f"no JSON line in {name}")


@pytest.mark.parametrize("name", ["r01_final_f32_b256.json", "r01_final_i8_b1024.json", "r02_final_default.json",
                                  "r02_final_i8_b1024.json"])
def test_our_arm_line_has_the_contract_keys(name):
    j = _line(name)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in j, key
    assert j["metric"] == "queries/sec" and j["unit"] == "queries/s" and j["higher_is_better"] is True
    assert j["n_gpus"] == 1 and j["warmup"] >= 3 and j["data"] == "synthetic" and j["vs_baseline"] is None
    assert j["dtype"] in ("f32", "i8", "f16") and "workload" in j["config"] and "model" not in j["config"]
    assert j["value"] > 0 and abs(j["ms_per_step"] * 1e-3 * j["value"] - j["config"]["batch"]) < 1e-6 * j["config"]["batch"] + 1e-3
    e = j["e2e"]
    assert e["value"] > 0 and e["unit"] == "queries/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] != j["value"], "e2e must be measured, not copied from the device-resident number"
    assert j["gpu_launches"] > 0
    r = j["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s", "TOP/s")
    assert r["peak"] > 0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] > 0
    c = j["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    k = j["clocks"]
    assert k["sm_mhz"] > 0 and k["sm_max_mhz"] >= k["sm_mhz"] and isinstance(k["reasons"], list)
    assert not set(k["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    p = j["parity"]
    assert p.get("bit_exact", p.get("within_1e-5")) is True


def test_reference_arm_line():
    j = _line("r01_final_reference.json")
    assert j["impl"] == "reference" and j["metric"] == "queries/sec" and j["unit"] == "queries/s"
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_multi_gpu_lines_scale():
    one = _line("r01_final_f32_b256.json")["value"]
    for n in (2, 4, 8):
        j = _line(f"r01_final_multi_f32_b256_n{n}.json")
        assert j["n_gpus"] == n and j["scaling"] == "strong" and j["value"] > one


def test_round2_default_line_carries_the_targets_and_the_measured_peak():
    """C2 / C3 sub-records, the int8 tensor peak measured in the same run, counted traffic,
    a sustained figure, and a parity leg that names the batch shape that was timed."""
    j = _line("r02_final_default.json")
    r = j["roofline"]
    assert r["bound"] == "tensor" and r["int8_peak_measured"]["int8_tops_burst"] > 1000
    assert abs(r["peak"] - r["int8_peak_measured"]["int8_tops_burst"]) < 1e-6 and "measured in this run" in r["peak_source"]
    assert "f32_hbm_roofline_frac" not in r and "counted" in r["traffic_note"]
    assert r["traffic"] > r["algorithmic_bytes_per_step"]          # the re-scorer's reads are in
    s = j["sustained"]
    assert s["seconds"] >= 3.0 and s["value"] > 0 and s["clocks"]["sm_mhz"] > 0
    assert "all 256 queries of the timed batch shape" in j["parity"]["checked"]
    c2, c3 = j["configs"]["C2"], j["configs"]["C3"]
    assert c2["workload"].startswith("1Mx768 f32 cosine") and "batch=256" in c2["workload"] and c2["value"] > 0
    assert c3["workload"].startswith("10Mx768 i8 dot") and "batch=1024" in c3["workload"]
    assert c3["roofline"]["tensor_frac_of_nominal"] >= 0.40          # north_star: >= 40 % of the int8 tensor-pipe peak
    ref = _line("r02_final_reference.json")
    assert ref["config"] == j["config"], "both arms must describe the same configuration"


def test_round2_eight_gpu_line_holds_configs_4_and_5_and_an_eight_rank_parity_leg():
    """BASELINE configs 4 and 5 on 8 GPUs and N = 8 parity, as part of the line
    `bench.py --gpus 8` prints under torchrun (the driver's own scaling run takes the same path)."""
    j = _line("r02_final_multi_default_n8.json")
    one = _line("r02_final_default.json")
    assert j["n_gpus"] == 8 and j["scaling"] == "strong" and j["value"] > one["value"]
    assert j["config"]["parallelism"] == "row-shard x8" and j["e2e"]["value"] > 0 and j["sustained"]["value"] > 0
    p = j["parity"]
    assert p["checked"].startswith("8 ranks x") and p["within_1e-5"] and p["counts_equal"] and p["ids_equal_frac"] == 1.0
    c4, c5 = j["configs"]["C4"], j["configs"]["C5"]
    assert c4["workload"].startswith("50Mx512 f16 cosine") and "batch=4096" in c4["workload"] and c4["n_gpus"] == 8
    assert c4["value"] > 0 and c4["e2e"] > 0
    assert "tag bitmap" in c5["workload"] and c5["n_gpus"] == 8 and c5["value"] > 0
    two = _line("r02_final_multi_default_n2.json")
    assert two["n_gpus"] == 2 and one["value"] < two["value"] < j["value"]
