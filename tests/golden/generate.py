"""Generates the frozen fixtures tests/golden/*.npz from the CPU oracle (SURVEY.md 8c pin 4).

The reference has no golden vectors and cannot run here, so these fixtures freeze the PINNED oracle
(identities + independent re-derivation) for the named configs: pattern hashes, dof counts, block norms,
a few hundred sampled entries and SpMV results for seeded vectors.  Regenerate with
    python tests/golden/generate.py
and commit the result only when a change of the oracle is intended.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = {
    "shell_r2_classic": dict(spec=dict(geometry="shell", refine=2), params="shell_3d_classic"),
    "cube_r2_classic": dict(spec=dict(geometry="cube", refine=2), params="cube_3d"),
    "shell_r1_classic_Tq2": dict(spec=dict(geometry="shell", refine=1, temperature_degree=2), params="shell_3d_classic"),
    "annulus_r3_classic_2d": dict(spec=dict(geometry="annulus", dim=2, refine=3, R0=10.0, R1=30.0, temperature_degree=2),
                                  params="annulus_2d"),
    "shell_r2_feec": dict(spec=dict(geometry="shell", refine=2, family="feec"), params="shell_3d_feec"),
    "cube_r2_feec": dict(spec=dict(geometry="cube", refine=2, family="feec"), params="cube_3d"),
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def summarize(name, vals, rng, out):
    out[name + ".sum"] = vals.sum()
    out[name + ".l2"] = np.sqrt((vals ** 2).sum())
    out[name + ".max"] = np.abs(vals).max() if vals.size else 0.0
    idx = rng.choice(vals.size, size=min(300, vals.size), replace=False) if vals.size else np.zeros(0, dtype=np.int64)
    idx.sort()
    out[name + ".sample_idx"] = idx
    out[name + ".sample_val"] = vals[idx]


def build(case):
    import dycore_b200  # noqa: F401
    from dycore_b200 import harness, params
    from oracle import oracle as orc
    from util import split_blocks, synthetic_fields
    P = harness.Problem(**case["spec"])
    mp = params.NAMED[case["params"]]
    prm = orc.params_from(mp)
    if case["spec"].get("family") == "feec":
        return build_feec(P, mp, prm)
    u, T = synthetic_fields(P)
    rng = np.random.default_rng(7)
    out = {"n_cells": P.n_cells, "n_u": P.scalar("nse.n_u"), "n_p": P.scalar("nse.n_p"), "n_T": P.scalar("temp.n_dofs")}
    for pat in ("nse.full", "pre.full", "temp.pat"):
        out[pat + ".nnz"] = P.scalar(pat + ".nnz")
        out[pat + ".sha"] = sha(P[pat + ".rowptr"]) + sha(P[pat + ".col"])
    out["nse.l2g.sha"] = sha(P["nse.l2g"])
    out["nse.cs.sha"] = sha(P["nse.cs.line_dof"]) + sha(P["nse.cs.entry_dof"])
    vals, rhs = orc.assemble_nse_system(P, prm, u, T)
    for (bi, bj), v in split_blocks(P, "nse", vals).items():
        summarize(f"nse.b{bi}{bj}", v, rng, out)
    summarize("nse.rhs", rhs, rng, out)
    for (bi, bj), v in split_blocks(P, "pre", orc.assemble_nse_preconditioner(P, prm)).items():
        summarize(f"pre.b{bi}{bj}", v, rng, out)
    m, k = orc.assemble_temperature_matrix(P, prm)
    summarize("temp.mass", m, rng, out)
    summarize("temp.stiff", k, rng, out)
    tm = orc.temperature_matrix_combine(m, k, mp.time_step / mp.NSE_solver_interval)
    summarize("temp.matrix", tm, rng, out)
    summarize("temp.rhs", orc.assemble_temperature_rhs(P, prm, T, u), rng, out)
    x = np.random.default_rng(1).standard_normal(P.scalar("nse.n_dofs"))
    rp, col, _, _ = P.csr("nse.full")
    summarize("spmv.nse", orc.spmv(rp, col, vals, x), rng, out)
    xt = np.random.default_rng(2).standard_normal(P.scalar("temp.n_dofs"))
    rp, col, _, _ = P.csr("temp.pat")
    summarize("spmv.temp", orc.spmv(rp, col, tm, xt), rng, out)
    return out


def feec_fields(P):
    rng = np.random.default_rng(20261018)
    u = np.ascontiguousarray(0.1 * rng.uniform(-1, 1, P.scalar("nse.n_dofs")) + 0.05)
    T = np.ascontiguousarray(2.0 + 0.3 * rng.uniform(-1, 1, P.scalar("temp.n_dofs")))
    return u, T


def build_feec(P, mp, prm):
    from oracle import oracle as orc
    u, T = feec_fields(P)
    rng = np.random.default_rng(7)
    out = {"n_cells": P.n_cells, "n_w": P.scalar("nse.n_w"), "n_u": P.scalar("nse.n_u"), "n_p": P.scalar("nse.n_p"),
           "n_T": P.scalar("temp.n_dofs")}
    for pat in ("nse.full", "pre.full", "temp.pat"):
        out[pat + ".nnz"] = P.scalar(pat + ".nnz")
        out[pat + ".sha"] = sha(P[pat + ".rowptr"]) + sha(P[pat + ".col"])
    out["nse.l2g.sha"] = sha(P["nse.l2g"])
    out["nse.sign.sha"] = sha(P["nse.sign"])
    vals, rhs = orc.feec_assemble_nse_system(P, prm, u, T)
    summarize("nse.full", vals, rng, out)
    summarize("nse.rhs", rhs, rng, out)
    summarize("pre.full", orc.feec_assemble_nse_preconditioner(P, prm), rng, out)
    m, k = orc.feec_assemble_temperature_matrix(P, prm)
    summarize("temp.mass", m, rng, out)
    summarize("temp.stiff", k, rng, out)
    summarize("temp.rhs", orc.feec_assemble_temperature_rhs(P, prm, T, u), rng, out)
    x = np.random.default_rng(1).standard_normal(P.scalar("nse.n_dofs"))
    rp, col, _, _ = P.csr("nse.full")
    summarize("spmv.nse", orc.spmv(rp, col, vals, x), rng, out)
    return out


if __name__ == "__main__":
    for name, case in CASES.items():
        out = build(case)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: out[k] for k in ("n_cells", "n_u", "n_p", "n_T", "nse.full.nnz")})
