"""The headline numbers README.md quotes must be the ones in the committed evidence files under profiles/ (the round-1
review found stale figures in the docs): a cheap guard that fails when a bench line is refreshed and the text is not."""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as fh:
        return json.loads(fh.read().strip().splitlines()[-1])


def _readme():
    with open(os.path.join(ROOT, "README.md"), encoding="utf-8") as fh:
        return fh.read()


def test_config2_range_in_readme_is_the_range_of_the_final_bench_lines():
    lines = [_line(n) for n in ("r2_bench_default_final.json", "r2_bench_default_final_box2.json",
                                "r2_bench_default_final_box3.json", "r2_bench_config2_final2.json",
                                "r2_bench_default_head.json")]
    for d in lines:   # every one of them is the headline configuration, measured the way the contract says
        assert d["config"]["workload"].startswith("vit_small/16 DINO multi-crop 2x224+10x96, batch 256/GPU")
        assert "NON-DEFAULT" not in d["config"]["workload"] and d["dtype"] == "bf16" and d["warmup"] >= 3
        assert d["gpu_launches"] > 0 and d["roofline"]["bound"] == "tensor" and d["cpu_baseline"]["kind"] == "port"
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    lo, hi = min(d["value"] for d in lines), max(d["value"] for d in lines)
    ms_hi, ms_lo = max(d["ms_per_step"] for d in lines), min(d["ms_per_step"] for d in lines)
    text = _readme()
    assert f"| 1 GPU, crops resident in HBM | {lo:,.0f} – {hi:,.0f} | {ms_hi:.1f} – {ms_lo:.1f} |" in text
    best = max(lines, key=lambda d: d["value"])
    assert 0.68 < best["roofline"]["frac"] < 0.71 and "0.69 of the measured sustained bf16 peak" in text


def test_other_configurations_in_readme_are_the_final_bench_lines():
    text = _readme()
    for cfg, unit in (("1", "images/s"), ("4", "images/s"), ("5a", "images/s"), ("5b", "images/s")):
        d = _line(f"r2_bench_config{cfg}_final.json")
        assert d["unit"] == unit and f"{d['value']:,.0f}" in text, (cfg, d["value"])
    drop = _line("r2_bench_config2_drop01.json")
    assert "NON-DEFAULT drop_rate 0.1" in drop["config"]["workload"]          # never mistaken for the headline line


def test_gpu_test_count_in_readme_is_the_last_log():
    with open(os.path.join(ROOT, "profiles", "r2_pytest_gpu_final4.log")) as fh:
        m = re.search(r"(\d+) passed", fh.read())
    assert m and f"({m.group(1)} GPU tests" in _readme()
