#!/usr/bin/env python
"""bench.py -- DoFs assembled/s + SpMV GB/s (FP64) on the hypershell, classic Taylor-Hood config.

One "step" = one pass of the hot path over one mesh: the four assemblers of the reference's time loop
(assemble_nse_system, assemble_nse_preconditioner, assemble_temperature_matrix, assemble_temperature_rhs;
/root/reference/include/core/boussinesq_model.tpp:1867-1884) followed by the SpMVs one outer Krylov iteration
makes (full nse_matrix block vmult + temperature_matrix vmult).  `value` = DoFs (n_u+n_p+n_T) / step time with
all inputs resident in HBM; `e2e` = the same step through the C ABI with pinned HOST buffers (H2D of the solution
vectors and SpMV sources, D2H of the right-hand sides and SpMV results inside the timed region, moved with the
library's asynchronous copies behind the kernels; DCP_NO_ASYNC_E2E=1: the blocking DCP_HOST path of the entry points).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--refine R] [--temperature-degree D]
  python bench.py --impl reference ...   # the restated CPU path (oracle, OpenMP) on the host cores
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


FP64_DMMA_PEAK_TFLOPS = 37.0


def ncu_traffic(strategy, refine):
    """dram__bytes_read.sum + dram__bytes_write.sum of the NSE system pass (all its launches of one step, single GPU)
    from the `ncu --set full` capture recorded in profiles/traffic.json (written from the .ncu-rep by
    profiles/extract_traffic.py together with the commit it was taken at); None when no capture matches."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        e = d.get(f"{strategy}:r{refine}")
        return (float(e["dram_bytes_per_step"]), e.get("source")) if e else (None, None)
    except Exception:
        return None, None


def key_field(keys, seed):
    """Deterministic field as a function of the global dof identity: every partition of the same mesh sees the same
    values, so the checksums of a run at N ranks can be compared with the run at 1 rank."""
    k = np.asarray(keys, dtype=np.int64)
    return np.ascontiguousarray(np.sin(0.37 * (k % 1000003) + seed) + 0.1 * np.cos(0.011 * (k % 7919)))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 7 and r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def algorithmic_bytes(P, T_deg_rule_shared=True, nb=2):
    """Per-kernel algorithmic bytes (DESIGN.md 'roofline accounting').  nb = 3: FEEC family (the mapping
    record carries J, J^-1, det J and x_q: 23 doubles per point; the preconditioner runs on its own 8-point rule)."""
    nc, nq, dim = P.n_cells, P.scalar("q_nse.nq"), P.dim
    nqt = P.scalar("q_temp.nq")
    nd, ndt = P.scalar("nse.n_local"), P.scalar("temp.n_local")
    n_nse, n_t = P.scalar("nse.n_dofs"), P.scalar("temp.n_dofs")
    geom_n = nc * nq * (1 + dim * dim + dim) * 8
    geom_n_noxq = nc * nq * (1 + dim * dim) * 8
    if nb == 3:
        geom_n = nc * nq * 23 * 8
        geom_n_noxq = nc * 8 * 23 * 8
    geom_t = nc * nqt * (1 + dim * dim) * 8
    nnz_nse = sum(P.scalar(f"nse.b{i}{j}.nnz") for i in range(nb) for j in range(nb))
    nnz_pre = sum(P.scalar(f"pre.b{i}{j}.nnz") for i in range(nb) for j in range(nb))
    nnz_t = P.scalar("temp.pat.nnz")
    out = {
        "nse_system": geom_n + nc * nd * 4 + nc * ndt * 4 + n_nse * 8 + n_t * 8 + nnz_nse * 8 + n_nse * 8,
        "nse_preconditioner": geom_n_noxq + nc * nd * 4 + nnz_pre * 8,
        "temperature_matrix": geom_t + nc * ndt * 4 + 2 * nnz_t * 8,
        "temperature_rhs": geom_t + nc * ndt * 4 + nc * nd * 4 + n_nse * 8 + n_t * 8 + 3 * nnz_t * 8 + n_t * 8,
    }
    spmv = 0
    for i in range(nb):
        first = True
        for j in range(nb):
            z = P.scalar(f"nse.b{i}{j}.nnz")
            if z:
                spmv += z * 12 + P.scalar(f"nse.b{i}{j}.n_rows") * 16 + P.scalar(f"nse.b{i}{j}.n_cols") * 8
                if not first:
                    spmv += P.scalar(f"nse.b{i}{j}.n_rows") * 8   # further blocks of a block row are applied as vmult_add
                first = False
    out["spmv_nse"] = spmv
    out["spmv_temperature"] = nnz_t * 12 + n_t * 16 + n_t * 8
    return out


def cpu_leg(refine, temperature_degree, steps, warmup, mp, family="classic"):
    """The restated CPU path (oracle, OpenMP over cells with atomic adds) on the host cores."""
    import dycore_b200  # noqa: F401
    from dycore_b200 import harness
    from oracle import oracle as orc
    feec = family == "feec"
    if feec:
        P = harness.Problem(geometry="shell", refine=refine, family="feec", threads=os.cpu_count() or 1)
    else:
        P = harness.Problem(geometry="shell", refine=refine, temperature_degree=temperature_degree,
                            threads=os.cpu_count() or 1)
    u = 0.1 * key_field(P["nse.dof_key"], 1.0)
    T = 2.0 + 0.2 * key_field(P["temp.dof_key"], 2.0)
    asm_sys = orc.feec_assemble_nse_system if feec else orc.assemble_nse_system
    asm_pre = orc.feec_assemble_nse_preconditioner if feec else orc.assemble_nse_preconditioner
    asm_tm = orc.feec_assemble_temperature_matrix if feec else orc.assemble_temperature_matrix
    asm_tr = orc.feec_assemble_temperature_rhs if feec else orc.assemble_temperature_rhs
    prm = orc.params_from(mp)
    n_dofs = P.scalar("nse.n_dofs") + P.scalar("temp.n_dofs")
    rp, col, _, _ = P.csr("nse.full")
    rpt, colt, _, _ = P.csr("temp.pat")
    x = np.random.default_rng(1).standard_normal(P.scalar("nse.n_dofs"))
    xt = np.random.default_rng(2).standard_normal(P.scalar("temp.n_dofs"))
    times, t_spmv = [], []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        vals, _ = asm_sys(P, prm, u, T, use_omp=True)
        asm_pre(P, prm, use_omp=True)
        m, k = asm_tm(P, prm, use_omp=True)
        tm = orc.temperature_matrix_combine(m, k, mp.time_step / mp.NSE_solver_interval)
        asm_tr(P, prm, T, u, use_omp=True)
        t1 = time.perf_counter()
        orc.spmv(rp, col, vals, x, use_omp=True)
        orc.spmv(rpt, colt, tm, xt, use_omp=True)
        t2 = time.perf_counter()
        if it >= warmup:
            times.append(t2 - t0)
            t_spmv.append(t2 - t1)
    ab = algorithmic_bytes(P, nb=3 if feec else 2)
    dt = float(np.mean(times))
    return {"value": n_dofs / dt, "unit": "DoFs/s", "cores": orc.max_threads(), "kind": "port",
            "sample": f"hypershell {family} refine={refine} ({P.n_cells} cells, {n_dofs} DoFs), {steps} step(s) "
                      f"after {warmup} warm-up; restated CPU path (oracle), OpenMP, not the deal.II/Trilinos MPI binary",
            "ms_per_step": dt * 1e3,
            "spmv_gbs": (ab["spmv_nse"] + ab["spmv_temperature"]) / float(np.mean(t_spmv)) / 1e9}, n_dofs, dt


def emit(line):
    """The one JSON line goes to the real stdout; everything else (NCCL banners, warnings) was sent to stderr."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # libraries that print to fd 1 (e.g. "NCCL version ...") must not pollute the JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--refine", type=int, default=0,
                    help="global refinements of the 6-tree shell; 0 = 6 (41.2 M DoFs, the largest that fits one B200) "
                         "when the box has the host and device memory for it, else 5")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N>1: strong = the same shell cut into N chunks; weak = N x the radial layers")
    ap.add_argument("--temperature-degree", type=int, default=0,
                    help="0 = 1 on one GPU (41.2 M DoFs at refine 6, the largest that fits 180 GB) and 2 on several "
                         "(52.3 M DoFs at refine 6: the >= 50 M-DoF point of the 6-tree shell)")
    ap.add_argument("--family", default="classic", choices=["classic", "feec"],
                    help="classic = Taylor-Hood Q2/Q1 (the headline metric); feec = the FEEC element family of "
                         "aqua_planet_shell_test_3d-feec.prm on the same shell (Nedelec/Raviart-Thomas/DG0, 19 dofs per cell)")
    ap.add_argument("--strategy", default="auto", choices=["auto", "search", "positions", "owner", "staged"])
    ap.add_argument("--cpu-refine", type=int, default=4, help="refinement of the CPU sample (4: ~3 s per pass on 16 cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--overlap-halo", default="nse", choices=["none", "nse", "all"],
                    help="N>1: hide the ghost exchange behind the rows without ghost columns (dcp_halo_block_vmult with "
                         "overlap = 1).  Default: for nse_matrix only -- measured at 8 GPUs the overlapped product gains 12 %% "
                         "on the 8 G-nonzero Stokes matrix and loses on the small temperature matrix, whose product is "
                         "shorter than the extra launches")
    ap.add_argument("--torch-halo", action="store_true",
                    help="N>1: exchange through torch.distributed p2p (round-1 path) instead of the library's NCCL halo")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.temperature_degree == 0:
        args.temperature_degree = 2 if (world > 1 and args.scaling == "strong" and args.refine in (0, 6)) else 1

    import dycore_b200  # noqa: F401
    from dycore_b200 import params
    mp = params.NAMED["shell_3d_classic"]

    if args.impl == "reference":
        if rank != 0:
            return
        warm = min(args.warmup, 1)
        steps = min(args.steps, 3)
        if args.family == "feec":
            mp = params.NAMED["shell_3d_feec"]
        base, n_dofs, dt = cpu_leg(args.cpu_refine, max(args.temperature_degree, 1), steps, warm, mp, args.family)
        line = {"impl": "reference", "metric": "dofs_assembled_per_s", "value": base["value"], "unit": "DoFs/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3,
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": ("hypershell FEEC (Nedelec/Raviart-Thomas/DG0 + Q1 temperature)" if args.family == "feec" else
                                        "hypershell classic (Taylor-Hood Q2/Q1 + Q1 temperature)") + ": full Boussinesq "
                                       "assembly pass + nse_matrix/temperature_matrix SpMV",
                           "note": "bounded sample: " + base["sample"]},
                "cpu_baseline": base,
                "e2e": {"value": base["value"], "unit": "DoFs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return

    import torch
    import torch.distributed as dist
    from dycore_b200 import device, harness
    from util import synthetic_fields
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- problem.  N ranks share ONE shell, partitioned along the (tree, Morton) curve like the reference's p4est
    # partition; every rank builds only its own subdomain (owned cells + one ghost-cell layer).  strong: the shell
    # of the N=1 run; weak: N x the radial layers (synthetic refinement), i.e. fixed cells per GPU.
    refine = args.refine
    if refine == 0:
        host_gb = os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES") / 1e9
        dev_gb = torch.cuda.mem_get_info(local_rank)[1] / 1e9
        # refine 6 needs ~105 GB of host memory for the sparsity patterns and ~150 GB of HBM in total; under weak
        # scaling every rank would need that much host memory, so N > 1 weak runs use refine 5 per GPU
        shared = world == 1 or args.scaling == "strong"
        refine = 6 if (shared and host_gb >= 150 and dev_gb * world >= 170) else 5
    t_setup = time.perf_counter()
    feec = args.family == "feec"
    nb = 3 if feec else 2
    if feec:
        mp = params.NAMED["shell_3d_feec"]
    spec = dict(geometry="shell", refine=refine, temperature_degree=args.temperature_degree, geometry_data=0,
                threads=max(1, (os.cpu_count() or 1) // world))  # torchrun pins OMP_NUM_THREADS=1
    if feec:
        spec.update(family="feec", temperature_degree=1)
    if world > 1:
        spec.update(n_ranks=world, rank=rank)
        if args.scaling == "weak":
            spec.update(radial_factor=world)
    P = harness.Problem(**spec)
    # state and SpMV sources as functions of the global dof identity (the same numbers on every partition)
    u = 0.1 * key_field(P["nse.dof_key"], 1.0)
    T = 2.0 + 0.2 * key_field(P["temp.dof_key"], 2.0)
    n_nse, n_t = P.scalar("nse.n_dofs"), P.scalar("temp.n_dofs")
    block_names = ("n_w", "n_u", "n_p") if feec else ("n_u", "n_p")
    block_sizes = [P.scalar("nse." + b) for b in block_names]
    owned = [P.scalar("nse." + b + "_owned") for b in block_names]
    n_u = block_sizes[0]   # first block (classic: velocity)
    n_dofs = sum(owned) + P.scalar("temp.n_owned")
    ctx = device.Context(local_rank)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    names = {"search": device.STRATEGY_SEARCH, "positions": device.STRATEGY_POSITIONS, "owner": device.STRATEGY_OWNER,
             "staged": device.STRATEGY_STAGED}
    model = device.BoussinesqModel.from_problem(ctx, P, mp, owner_plan=(args.strategy == "owner"), device_geometry=True)
    if args.strategy != "auto":
        model.set_strategy(names[args.strategy])
    strategy = {v: k for k, v in names.items()}[model.strategy]   # auto: the library's default (staged when the model qualifies)
    halo_nse = halo_t = comm = None
    if world > 1:
        from dycore_b200 import distributed
        model.set_owned(owned, P.scalar("temp.n_owned"))
        torch.cuda.set_stream(stream)
        halo_nse = distributed.HaloPlan(P["nse.dof_key"], P["nse.dof_owner"], rank, world, device="cuda")
        halo_t = distributed.HaloPlan(P["temp.dof_key"], P["temp.dof_owner"], rank, world, device="cuda")
        comm = distributed.Communicator(ctx, rank, world)
        if args.torch_halo:     # round-1 path: pack kernel, torch.distributed p2p, unpack kernel
            if args.overlap_halo == "all":
                op_nse = distributed.OverlappedMatrix(model.nse_matrix, halo_nse, local_rank, stream)
                op_t = distributed.OverlappedMatrix(model.temperature_matrix, halo_t, local_rank, stream)
            else:
                op_nse = distributed.DistributedMatrix(model.nse_matrix, halo_nse, ctx)
                op_t = distributed.DistributedMatrix(model.temperature_matrix, halo_t, ctx)
        else:                   # the library's data plane: Epetra_Import + Multiply as one C-ABI call
            dh_nse, dh_t = distributed.DeviceHalo(halo_nse, comm), distributed.DeviceHalo(halo_t, comm)
            op_nse = distributed.HaloMatrix(model, device.MAT_NSE, dh_nse, overlap=args.overlap_halo in ("nse", "all"))
            op_t = distributed.HaloMatrix(model, device.MAT_TEMP, dh_t, overlap=args.overlap_halo == "all")
    t_setup = time.perf_counter() - t_setup

    with torch.cuda.stream(stream):
        d_u = torch.from_numpy(u).cuda()
        d_T = torch.from_numpy(T).cuda()
        x_np, xt_np = key_field(P["nse.dof_key"], 3.0), key_field(P["temp.dof_key"], 5.0)
        if world > 1:          # ghost slots start empty: the halo exchange has to fill them
            x_np[P["nse.dof_owner"] != rank] = 0.0
            xt_np[P["temp.dof_owner"] != rank] = 0.0
        d_x = torch.from_numpy(x_np).cuda()
        d_y = torch.zeros(n_nse, dtype=torch.float64, device="cuda")
        d_xt = torch.from_numpy(xt_np).cuda()
        d_yt = torch.zeros(n_t, dtype=torch.float64, device="cuda")
    # pinned host buffers for the e2e leg
    h_u, h_T = torch.from_numpy(u).pin_memory(), torch.from_numpy(T).pin_memory()
    h_x, h_y = d_x.cpu().pin_memory(), torch.zeros(n_nse, dtype=torch.float64).pin_memory()
    h_xt, h_yt = d_xt.cpu().pin_memory(), torch.zeros(n_t, dtype=torch.float64).pin_memory()
    h_rhs, h_trhs = torch.zeros(n_nse, dtype=torch.float64).pin_memory(), torch.zeros(n_t, dtype=torch.float64).pin_memory()

    phases = ["nse_system", "nse_preconditioner", "temperature_matrix", "temperature_rhs", "spmv_nse", "spmv_temperature"]

    def step_device(ev=None):
        def mark(i):
            if ev is not None:
                ev[i].record(stream)
        with torch.cuda.stream(stream):
            mark(0)
            model.assemble_nse_system(d_u, d_T)
            mark(1)
            model.assemble_nse_preconditioner()
            mark(2)
            model.assemble_temperature_matrix()
            mark(3)
            model.assemble_temperature_rhs(d_T, d_u)
            mark(4)
            if halo_nse is not None:
                op_nse.vmult(d_y, d_x)          # Epetra_Import of the ghost columns (NCCL p2p over NVLink) + product
            else:
                model.nse_matrix.vmult(d_y, d_x)
            mark(5)
            if halo_t is not None:
                op_t.vmult(d_yt, d_xt)
            else:
                model.temperature_matrix.vmult(d_yt, d_xt)
            mark(6)

    def hp(t):
        return ctypes.c_void_p(t.data_ptr())

    L = device.lib()

    COMPUTE_WAITS, COPIES_WAIT = 0, 1
    rhs_ptr, trhs_ptr = ctypes.c_void_p(), ctypes.c_void_p()
    _n = ctypes.c_int64()
    device.check(L.dcp_vector_device(model._h, device.VEC_NSE_RHS, ctypes.byref(rhs_ptr), ctypes.byref(_n)), "dcp_vector_device")
    device.check(L.dcp_vector_device(model._h, device.VEC_TEMP_RHS, ctypes.byref(trhs_ptr), ctypes.byref(_n)), "dcp_vector_device")

    def step_host():
        # the same step for a caller whose vectors live in (pinned) HOST memory, through the C ABI: every input comes up and
        # every result goes down inside the timed region, with the library's asynchronous copies (dcp_memcpy_*_async on
        # the context's copy stream, ordered by dcp_copy_fence) so that the SpMV sources travel while the assemblers run
        # and the right-hand sides while the products run -- what a deal.II host does with the vectors of the next /
        # previous operator call.  (DCP_NO_ASYNC_E2E=1: the blocking host-pointer path of the entry points, DCP_HOST.)
        if os.environ.get("DCP_NO_ASYNC_E2E") and halo_nse is None:
            model.assemble_nse_system(h_u, h_T)
            model.assemble_nse_preconditioner()
            model.assemble_temperature_matrix()
            model.assemble_temperature_rhs(h_T, h_u)
            device.check(L.dcp_vector_download(model._h, device.VEC_NSE_RHS, hp(h_rhs)), "dcp_vector_download")
            device.check(L.dcp_vector_download(model._h, device.VEC_TEMP_RHS, hp(h_trhs)), "dcp_vector_download")
            model.nse_matrix.vmult(h_y, h_x)
            model.temperature_matrix.vmult(h_yt, h_xt)
            return
        c = ctx._h
        device.check(L.dcp_copy_fence(c, COPIES_WAIT))                         # the previous step is done with d_u, d_T, d_x
        device.check(L.dcp_memcpy_h2d_async(c, hp(d_u), hp(h_u), 8 * n_nse))   # old solution
        device.check(L.dcp_memcpy_h2d_async(c, hp(d_T), hp(h_T), 8 * n_t))
        device.check(L.dcp_copy_fence(c, COMPUTE_WAITS))
        device.check(L.dcp_memcpy_h2d_async(c, hp(d_x), hp(h_x), 8 * n_nse))   # SpMV sources: behind the assemblers
        device.check(L.dcp_memcpy_h2d_async(c, hp(d_xt), hp(h_xt), 8 * n_t))
        model.assemble_nse_system(d_u, d_T)
        model.assemble_nse_preconditioner()
        model.assemble_temperature_matrix()
        model.assemble_temperature_rhs(d_T, d_u)
        device.check(L.dcp_copy_fence(c, COMPUTE_WAITS))                       # sources are up
        device.check(L.dcp_copy_fence(c, COPIES_WAIT))                         # right-hand sides are assembled
        device.check(L.dcp_memcpy_d2h_async(c, hp(h_rhs), rhs_ptr, 8 * n_nse))  # ... and go down behind the products
        device.check(L.dcp_memcpy_d2h_async(c, hp(h_trhs), trhs_ptr, 8 * n_t))
        if halo_nse is not None:
            op_nse.vmult(d_y, d_x)
            op_t.vmult(d_yt, d_xt)
        else:
            model.nse_matrix.vmult(d_y, d_x)
            model.temperature_matrix.vmult(d_yt, d_xt)
        device.check(L.dcp_copy_fence(c, COPIES_WAIT))
        device.check(L.dcp_memcpy_d2h_async(c, hp(h_y), hp(d_y), 8 * n_nse))
        device.check(L.dcp_memcpy_d2h_async(c, hp(h_yt), hp(d_yt), 8 * n_t))
        device.check(L.dcp_copy_fence(c, COMPUTE_WAITS))                       # the step ends when the results are on the host

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, with_events):
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(7)] for _ in range(steps)] if with_events else None
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for s in range(steps):
            fn(evs[s]) if with_events else fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, evs

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()   # runs through warm-up + timed region (a 100 ms sampler needs more than a few steps)
    t_w = time.perf_counter()
    n_warm = 0
    while n_warm < max(args.warmup, 3) or (time.perf_counter() - t_w < 0.5 and world == 1):
        step_device()
        torch.cuda.synchronize()
        n_warm += 1
    l0 = ctx.launch_count()
    ms, evs = timed(step_device, args.steps, True)
    launches = (ctx.launch_count() - l0) // args.steps
    clocks = sampler.stop() if rank == 0 else None
    phase_ms = {p: float(np.mean([evs[s][i].elapsed_time(evs[s][i + 1]) for s in range(args.steps)]))
                for i, p in enumerate(phases)}
    if rank == 0:
        print("[bench] device-resident: %.3f ms/step %s" % (ms / args.steps, json.dumps(phase_ms)), file=sys.stderr, flush=True)
    for _ in range(2):
        step_host()
    ms_e2e, _ = timed(step_host, args.steps, False)
    ctx.synchronize()

    # (M2) the operators the Krylov solves apply one by one: block(0,0), block(0,1), block(1,0) of nse_matrix and
    # temperature_matrix, each timed alone (single rank only: block products need no extra exchange pattern here)
    spmv_blocks = {}
    if world == 1 and not feec:
        n_p = n_nse - n_u
        views = {"b00": (d_y[:n_u], d_x[:n_u]), "b01": (d_y[:n_u], d_x[n_u:]), "b10": (d_y[n_u:], d_x[:n_u])}
        reps = 5
        for name, (dst, src) in views.items():
            A = model.nse_matrix.block(int(name[1]), int(name[2]))
            nnz = P.scalar(f"nse.{name}.nnz")
            nbytes = nnz * 12 + A.m() * 16 + A.n() * 8
            with torch.cuda.stream(stream):
                for _ in range(2):
                    A.vmult(dst, src)
                b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                b0.record(stream)
                for _ in range(reps):
                    A.vmult(dst, src)
                b1.record(stream)
            torch.cuda.synchronize()
            t_ms = b0.elapsed_time(b1) / reps
            spmv_blocks[name] = {"ms": t_ms, "gbs": nbytes / (t_ms * 1e-3) / 1e9, "nnz": int(nnz),
                                 "l2_resident": bool(nbytes < 126e6)}
        assert n_p == model.nse_matrix.block(1, 0).m()

    # correctness signal carried by every line: l2 norms over the owned entries of A x, the NSE right-hand side, the
    # temperature right-hand side and T x_T for inputs that are functions of the global dof identity -- the same
    # numbers (to summation order) for every number of ranks
    def owned_norm(t, ranges):
        if comm is not None:
            return float(np.sqrt(comm.dot(t, t, ranges)))
        return float(np.sqrt(sum(float(torch.dot(t[b:e], t[b:e])) for b, e in ranges)))
    with torch.cuda.stream(stream):
        step_device()
        nse_ranges, off = [], 0
        for bsz, own in zip(block_sizes, owned):
            nse_ranges.append((off, off + own))
            off += bsz
        t_ranges = [(0, P.scalar("temp.n_owned"))]
        d_rhs = torch.from_numpy(model.nse_rhs).cuda()
        d_trhs = torch.from_numpy(model.temperature_rhs).cuda()
        checksum = {"norm_A_x": owned_norm(d_y, nse_ranges), "norm_nse_rhs": owned_norm(d_rhs, nse_ranges),
                    "norm_temperature_rhs": owned_norm(d_trhs, t_ranges), "norm_T_x": owned_norm(d_yt, t_ranges)}

    total_dofs = n_dofs
    if world > 1:
        t = torch.tensor([n_dofs], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        total_dofs = int(t.item())
    if rank == 0:
        ab = algorithmic_bytes(P, nb=nb)
        peak, peak_src = peaks()
        ms_step = ms / args.steps
        # staged strategy: the system pass also writes the preconditioner matrix (its second gather), the preconditioner
        # call only refreshes the Jacobi diagonals -- the roofline of the fused pass counts both matrices' bytes
        fused = strategy == "staged" and phase_ms["nse_preconditioner"] < 0.05 * phase_ms["nse_system"]
        if fused:
            ab["nse_system"] += ab["nse_preconditioner"]
            ab["nse_preconditioner"] = sum(P.scalar(f"pre.b{i}{i}.nnz") * 8 + P.scalar(f"pre.b{i}{i}.n_rows") * 12 for i in range(nb))
        dom = max(("nse_system", "nse_preconditioner"), key=lambda k: phase_ms[k])
        ach = ab[dom] / (phase_ms[dom] * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic(strategy, refine) if (world == 1 and dom == "nse_system") else (None, None)
        asm_ms = sum(phase_ms[p] for p in phases[:4])
        spmv_ms = phase_ms["spmv_nse"] + phase_ms["spmv_temperature"]
        line = {
            "metric": "dofs_assembled_per_s", "value": total_dofs / (ms_step * 1e-3), "unit": "DoFs/s",
            "n_gpus": world, "steps": args.steps, "warmup": n_warm, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": (f"hypershell FEEC refine={refine} (Nedelec/Raviart-Thomas/DG0 + Q1 temperature, "
                                    "aqua_planet_shell_test_3d-feec.prm)" if feec else
                                    f"hypershell classic refine={refine} (Taylor-Hood Q2/Q1 + Q{args.temperature_degree} temperature)")
                                   + ": full Boussinesq assembly pass + nse_matrix/temperature_matrix SpMV",
                       "cells_per_gpu": P.scalar("n_owned_cells"), "dofs_per_gpu": n_dofs, "total_dofs": total_dofs,
                       "ghost_cells_rank0": P.n_cells - P.scalar("n_owned_cells"),
                       "nnz_nse": sum(P.scalar(f"nse.b{i}{j}.nnz") for i in range(nb) for j in range(nb)),
                       "strategy": strategy, "fused_preconditioner": bool(fused), "l2": "inputs larger than L2" if ab["nse_system"] > 4 * 126e6 else "inputs fit L2",
                       "partition": ((f"shell with {world}x radial layers" if args.scaling == "weak" else "the same shell")
                                     + f" cut into {world} contiguous (tree, Morton) chunks, one ghost-cell layer, "
                                     "ghost-dof halo over NCCL p2p") if world > 1 else "single GPU",
                       "mapping_data": "evaluated on the device from support points (dcp_geometry_create)",
                       "device_mem_gb_rank0": round((torch.cuda.mem_get_info(local_rank)[1] - torch.cuda.mem_get_info(local_rank)[0]) / 1e9, 1),
                       "setup_s": round(t_setup, 2)},
            "assembly_dofs_per_s": total_dofs / (asm_ms * 1e-3),
            "spmv_gbs": world * (ab["spmv_nse"] + ab["spmv_temperature"]) / (spmv_ms * 1e-3) / 1e9,
            "phase_ms": phase_ms,
            "phase_gbs": {p: ab[p] / (phase_ms[p] * 1e-3) / 1e9 for p in phases},
            "spmv_blocks": spmv_blocks,
            # FP64 roofline of the local-matrix contraction (0.60 MFLOP/cell, unsymmetrised structure-
            # exploiting count) against the DMMA peak measured by benchmarks/fp64_peaks.cu on this pool
            "fp64_roofline": None if feec else
                             {"kernel": "nse_system", "flop_per_cell": 0.60e6, "peak_tflops": FP64_DMMA_PEAK_TFLOPS,
                              "achieved_tflops": 0.60e6 * P.n_cells / (phase_ms["nse_system"] * 1e-3) / 1e12,
                              "frac": 0.60e6 * P.n_cells / (phase_ms["nse_system"] * 1e-3) / 1e12 / FP64_DMMA_PEAK_TFLOPS,
                              "peak_source": "profiles/r01_fp64_peaks.json (mma.sync m8n8k4 f64, measured)"},
            "roofline": {"bound": "hbm", "kernel": dom + (" (+ nse_preconditioner, fused pass)" if fused and dom == "nse_system" else ""), "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": ach / peak,
                         "traffic": traffic,
                         "traffic_note": ("dram__bytes_read.sum + dram__bytes_write.sum of the NSE system pass (all its launches "
                                          "of one step), " + (traffic_src or "no ncu capture for this strategy / refinement")
                                          + "; algorithmic bytes of the pass: %.3e" % ab[dom]),
                         "peak_source": peak_src,
                         "spmv_frac": ab["spmv_nse"] / (phase_ms["spmv_nse"] * 1e-3) / 1e9 / peak},
            "e2e": {"value": total_dofs / (ms_e2e / args.steps * 1e-3), "unit": "DoFs/s",
                    "h2d_bytes_per_step": int(8 * (2 * (n_nse + n_t))),
                    "d2h_bytes_per_step": int(8 * (2 * (n_nse + n_t)))},
            "gpu_launches": int(launches), "clocks": clocks,
            "parity_checksum": checksum,
        }
        if not args.no_cpu_baseline and world == 1:
            base, _, _ = cpu_leg(args.cpu_refine, args.temperature_degree, 3, 1, mp, args.family)
            line["cpu_baseline"] = base
        emit(line)
    model.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
