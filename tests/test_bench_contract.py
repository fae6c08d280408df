"""The parts of the bench contract that need no GPU: the reference arm (the restated CPU path timed on the host cores,
`bench.py --impl reference`) prints one JSON line with the contract's keys, and the DRAM-traffic record that
`roofline.traffic` is read from has the shape bench.py expects."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--cpu-refine", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "dofs_assembled_per_s" and line["unit"] == "DoFs/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1 and line["value"] > 0
    base = line["cpu_baseline"]
    assert base["kind"] == "port" and base["cores"] >= 1 and base["value"] == line["value"] and "refine=2" in base["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "DoFs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_traffic_record_matches_what_bench_reads():
    sys.path.insert(0, ROOT)
    import bench
    d = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    assert "staged:r6" in d, "the default workload (staged strategy, refine 6) has an ncu capture"
    e = d["staged:r6"]
    assert e["dram_bytes_per_step"] > 0 and "source" in e and any("th_gather_kernel" in k for k in e["kernels"])
    traffic, src = bench.ncu_traffic("staged", 6)
    assert traffic == e["dram_bytes_per_step"] and "ncu" in src
    assert bench.ncu_traffic("staged", 99) == (None, None)


def test_traffic_record_is_reproducible_from_the_committed_launch_list(tmp_path):
    """profiles/traffic.json is a sum over profiles/r02_launches_r6_final.csv (ncu launch list with DRAM counters): redo
    the sum with the committed script and compare."""
    import shutil
    work = tmp_path / "profiles"
    work.mkdir()
    shutil.copy(os.path.join(ROOT, "profiles", "extract_traffic.py"), work / "extract_traffic.py")
    csv_path = os.path.join(ROOT, "profiles", "r02_launches_r6_final.csv")
    r = subprocess.run([sys.executable, str(work / "extract_traffic.py"), csv_path, "staged", "6", "3"], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    redone = json.load(open(work / "traffic.json"))["staged:r6"]
    kept = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["staged:r6"]
    assert abs(redone["dram_bytes_per_step"] - kept["dram_bytes_per_step"]) <= 1e-9 * kept["dram_bytes_per_step"]
    per_step = {k: v["launches_per_step"] for k, v in redone["kernels"].items()}
    assert per_step[[k for k in per_step if k.startswith("th_stage_kernel")][0]] == 24   # 1 572 864 cells in chunks of 65 536
    assert per_step[[k for k in per_step if k.startswith("th_gather_kernel")][0]] == 24
