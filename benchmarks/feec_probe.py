"""Assemble the FEEC shell a few times (profiling target for ncu)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dycore_b200  # noqa: E402,F401
from dycore_b200 import device, harness, params  # noqa: E402

refine = int(sys.argv[1]) if len(sys.argv) > 1 else 4
P = harness.Problem(geometry="shell", family="feec", refine=refine)
ctx = device.Context(0)
model = device.BoussinesqModel.from_problem(ctx, P, params.NAMED["shell_3d_feec"])
rng = np.random.default_rng(1)
u = np.ascontiguousarray(0.1 * rng.uniform(-1, 1, P.scalar("nse.n_dofs")))
T = np.ascontiguousarray(2 + 0.3 * rng.uniform(-1, 1, P.scalar("temp.n_dofs")))
for _ in range(4):
    model.assemble_nse_system(u, T)
    model.assemble_nse_preconditioner()
ctx.synchronize()
print("done")
model.close()
ctx.close()
