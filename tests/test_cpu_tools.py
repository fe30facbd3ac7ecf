"""not gpu: the evidence tooling runs on the committed artefacts (launch-list summary,
roofline.traffic lookup), so a broken tool is noticed before GPU time is spent."""
import importlib.util
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_launch_summary_on_committed_csv():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"),
                          os.path.join(ROOT, "profiles", "r01_launches_c4_final.csv"), "cmd"],
                         capture_output=True, text=True, check=True).stdout
    lines = [ln for ln in out.splitlines() if "knn_screen_kernel<1>" in ln]
    assert len(lines) == 1
    share = float(lines[0].split()[-1].rstrip("%"))
    assert 85.0 < share < 95.0          # the dominant kernel's share of the step
    assert "cutlass" in out and out.count("%") > 10


def test_bench_traffic_lookup_reads_the_ncu_summary():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    traffic, source = bench.ncu_traffic("screen-dual")
    assert traffic is not None and 1e10 < traffic < 1e11
    assert "ncu_knn_screen_dual" in source and source.startswith("profiles/")
    assert bench.ncu_traffic("no-such-kind") == (None, None)
    # the CPU-leg sample stays bounded (about 10-30 s of work at C4)
    w = dict(bench.WORKLOADS["c4"])
    assert 1024 <= bench.cpu_sample_rows(w, 0) <= 8192


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) on a tiny workload:
    exactly one JSON line on stdout with the contract's keys."""
    import json

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--workload", "custom", "--n", "600", "--m", "500", "--d", "32",
                          "--c", "10", "--k", "5", "--steps", "1", "--warmup", "1",
                          "--cpu-sample", "128"], capture_output=True, text=True, check=True,
                         timeout=300).stdout
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert len(lines) == 1, out
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["higher_is_better"] is True
    assert line["metric"] == "queries_per_s" and line["unit"] == "queries/s"
    assert line["value"] > 0 and line["n_gpus"] == 1 and line["steps"] == 1
    # the unmodified reference (baseline/_ref or /root/reference) when a copy is present
    from oracle import ref_shim
    assert line["cpu_baseline"]["kind"] == ("reference" if ref_shim.reference_available() else "port")
    assert line["cpu_baseline"]["cores"] >= 1 and line["warmup"] == 1
    assert set(line["config"]) >= {"workload", "precision", "fused", "parallelism"}
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "600x500" in line["config"]["workload"]


def test_bench_reference_arm_port_fallback():
    import json

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--workload", "c1", "--steps", "1", "--warmup", "0", "--cpu-kind", "port"],
                         capture_output=True, text=True, check=True, timeout=300).stdout
    line = json.loads(out.strip())
    assert line["cpu_baseline"]["kind"] == "port" and line["value"] > 0


def test_bench_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--gpus", "2", "--workload", "c1"], capture_output=True, text=True,
                         check=True, timeout=120, env=env).stdout
    assert out.strip() == ""
