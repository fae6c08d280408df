#!/usr/bin/env python
"""Sum dram__bytes_read.sum + dram__bytes_write.sum over the launches of one bench step from an ncu CSV
(`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file <csv> python bench.py ...`) and record the
NSE system pass (th_stage_kernel + th_gather_kernel, or th_mma_kernel<1> for the reduction strategy) in profiles/traffic.json, which bench.py reads for `roofline.traffic`.

    python profiles/extract_traffic.py <csv> <strategy> <refine> <steps captured> [<commit>]
"""
import collections
import csv
import json
import os
import sys


def main():
    path, strategy, refine, steps = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    commit = sys.argv[5] if len(sys.argv) > 5 else ""
    rows = list(csv.reader(open(path)))
    hdr = None
    per_kernel = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if not hdr or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") not in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            continue
        v = float(d["Metric Value"].replace(",", ""))
        unit = d["Metric Unit"].lower()
        scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(unit, 1.0)
        name = d["Kernel Name"].replace("<unnamed>::", "").replace("void ", "")
        name = name.split("(thmma")[0].split("(GatherArgs")[0].split("(<unnamed>")[0].split("(MmaArgs")[0].strip()
        if d["Metric Name"] == "dram__bytes_read.sum":
            per_kernel[name][0] += 1
        per_kernel[name][1] += v * scale
    def in_system_pass(n):
        if any(k in n for k in ("th_stage_kernel", "th_gather", "th_pre_gather", "th_fused_kernel")):
            return True
        return "th_mma_kernel" in n and ("<1>" in n or "(bool)1" in n)
    total = sum(b for n, (c, b) in per_kernel.items() if in_system_pass(n))
    out_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json")
    data = json.load(open(out_path)) if os.path.exists(out_path) else {}
    data[f"{strategy}:r{refine}"] = {
        "dram_bytes_per_step": total / steps,
        "kernels": {n: {"launches_per_step": c / steps, "dram_bytes_per_step": b / steps} for n, (c, b) in sorted(per_kernel.items())},
        "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum of `python bench.py --refine {refine}` "
                  f"({os.path.basename(path)}, {steps} step(s) captured" + (f", commit {commit}" if commit else "") + ")",
    }
    json.dump(data, open(out_path, "w"), indent=1, sort_keys=True)
    print(json.dumps(data[f"{strategy}:r{refine}"], indent=1)[:1500])


if __name__ == "__main__":
    main()
