"""The evidence under profiles/ must stay readable by the tools that produced its summaries (CPU only): the ncu launch
list of one step, the per-kernel shares, the HBM table and the conv DRAM-traffic summary."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args):
    return subprocess.run([sys.executable, *args], cwd=ROOT, capture_output=True, text=True, timeout=120)


def test_launch_summary_reads_the_committed_launch_list():
    r = run("tools/launch_summary.py", "profiles/r02_launches_step.csv", "188")
    assert r.returncode == 0, r.stderr
    head = r.stdout.splitlines()[0]
    assert "step of 188 kernels" in head
    assert "conv_fprop_kernel<64, 0, 2>" in r.stdout and "conv_wgrad_kernel" in r.stdout   # CTA pairs are in the step


def test_hbm_table_matches_the_round2_kernel_sequence():
    r = run("tools/hbm_table.py", "profiles/r02_launches_step.csv")
    assert r.returncode == 0, r.stderr
    rows = [ln for ln in r.stdout.splitlines()[1:] if "|" in ln]
    assert len(rows) >= 30
    roles = " ".join(rows)
    assert "student pool1 backward" in roles and "SE squeeze, stage 2 (of the 3x3 output)" in roles
    for ln in rows:
        gbs = float(ln.split("|")[4])
        assert 100 < gbs < 9000, ln                      # a wrong tensor size would show up as an absurd bandwidth


def test_bench_lines_under_profiles_carry_the_contract_keys():
    for name in ("n1", "n2", "n4", "n8", "c2", "c3", "c5", "ref"):
        line = json.loads(open(os.path.join(ROOT, "profiles", "r02_bench_%s.json" % name)).read().strip().splitlines()[-1])
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                    "dtype", "data", "config", "e2e"):
            assert key in line, (name, key)
        assert "workload" in line["config"]
        if name == "ref":
            assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port"
            continue
        assert line["gpu_launches"] > 0 and "clocks" in line
        roof = line["roofline"]
        assert roof["bound"] == "tensor" and 0.0 < roof["frac"] < 1.0 and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-6
        assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    n1 = json.loads(open(os.path.join(ROOT, "profiles", "r02_bench_n1.json")).read().strip().splitlines()[-1])
    assert n1["cpu_baseline"]["value"] > 0 and n1["e2e"]["h2d_bytes_per_step"] > 0 and n1["kernels_per_step"] == 188
