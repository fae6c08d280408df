"""Per-assembler timings of the four named configs and of synthetic refinements of the FEEC shell,
one GPU.  Not the contract benchmark (that is bench.py at the repo root): this script documents the other element
family and the small named cases.  One JSON line per config on stdout.

    python benchmarks/bench_configs.py [--reps 5]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CONFIGS = [
    ("aqua_planet_test_2d.prm", dict(geometry="annulus", dim=2, R0=10.0, R1=30.0, temperature_degree=2, refine=4, renumber="cuthill_mckee"), "annulus_2d"),
    ("aqua_planet_cube_test_3d.prm", dict(geometry="cube", family="feec", refine=4), "cube_3d"),
    ("aqua_planet_shell_test_3d-classic.prm", dict(geometry="shell", refine=2), "shell_3d_classic"),
    ("aqua_planet_shell_test_3d-feec.prm", dict(geometry="shell", family="feec", refine=3), "shell_3d_feec"),
    ("shell FEEC refine 5 (synthetic)", dict(geometry="shell", family="feec", refine=5), "shell_3d_feec"),
    ("shell FEEC refine 6 (synthetic)", dict(geometry="shell", family="feec", refine=6), "shell_3d_feec"),
    ("shell classic refine 4, Q2 temperature (synthetic)", dict(geometry="shell", refine=4, temperature_degree=2), "shell_3d_classic"),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import torch
    import dycore_b200  # noqa: F401
    from dycore_b200 import device, harness, params
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device; the product path has no CPU fallback")
    ctx = device.Context(0)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    for name, spec, pname in CONFIGS:
        t0 = time.perf_counter()
        P = harness.Problem(threads=os.cpu_count() or 1, **spec)
        mp = params.NAMED[pname]
        feec = spec.get("family") == "feec"
        n, nT = P.scalar("nse.n_dofs"), P.scalar("temp.n_dofs")
        rng = np.random.default_rng(20261018)
        model = device.BoussinesqModel.from_problem(ctx, P, mp, device_geometry=True)
        setup = time.perf_counter() - t0
        with torch.cuda.stream(stream):
            u = torch.from_numpy(0.1 * rng.uniform(-1, 1, n) + 0.05).cuda()
            T = torch.from_numpy(2.0 + 0.3 * rng.uniform(-1, 1, nT)).cuda()
            x, y = torch.from_numpy(rng.standard_normal(n)).cuda(), torch.zeros(n, dtype=torch.float64, device="cuda")
            xt, yt = torch.from_numpy(rng.standard_normal(nT)).cuda(), torch.zeros(nT, dtype=torch.float64, device="cuda")
        calls = [("nse_system", lambda: model.assemble_nse_system(u, T)),
                 ("nse_preconditioner", model.assemble_nse_preconditioner),
                 ("temperature_matrix", model.assemble_temperature_matrix),
                 ("temperature_rhs", lambda: model.assemble_temperature_rhs(T, u)),
                 ("spmv_nse", lambda: model.nse_matrix.vmult(y, x)),
                 ("spmv_temperature", lambda: model.temperature_matrix.vmult(yt, xt))]
        ms = {}
        with torch.cuda.stream(stream):
            for _ in range(3):
                for _, f in calls:
                    f()
            for key, f in calls:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(args.reps):
                    f()
                e1.record(stream)
                e1.synchronize()
                ms[key] = e0.elapsed_time(e1) / args.reps
        total = sum(ms.values())
        nb = 3 if feec else 2
        nnz = sum(P.scalar(f"nse.b{i}{j}.nnz") for i in range(nb) for j in range(nb))
        print(json.dumps({"config": name, "family": "feec" if feec else "classic", "dim": P.dim, "cells": P.n_cells,
                          "dofs": n + nT, "nnz_nse": nnz, "ms": {k: round(v, 4) for k, v in ms.items()},
                          "ms_per_pass": round(total, 4), "dofs_per_s": (n + nT) / (total * 1e-3),
                          "setup_s": round(setup, 2)}), flush=True)
        model.close()
        P.close()
    ctx.close()


if __name__ == "__main__":
    main()
