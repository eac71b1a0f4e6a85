#!/usr/bin/env python
"""bench.py -- the round contract benchmark.

Metric (BASELINE.json): elements assembled per second for the FV1 Jacobian + defect of the incompressible
Navier-Stokes system, workload = config 3 (3-D cavity, hexahedra, LinearProfileSkewedUpwind + FIELDS/RAW,
nu = 1e-2) at the single-GPU share of the 368^3 mesh: 184^3 = 6.23 M elements per GPU (weak scaling:
N GPUs assemble N such blocks of one (2x2x2-blocked) global mesh and sum interface rows over NCCL).

A "step" = one pass of the hot path over the mesh: Jacobian AND defect of the stiffness part for every
element, scattered into the global CSR matrix / defect vector.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--cells 184] [--mode gather|colored|atomic]
  python bench.py --impl reference ...     # the reference arm: CPU oracle on the host cores

One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fv1_jacobian_defect_elements_assembled_per_s"
UNIT = "elements/s"
VISC = 1e-2
UPWIND, STAB = "lps", "fields"


def algorithmic_bytes(n_elem, n_node, nsh, dim, n_dof, nnz, n_timepoints=1):
    """connectivity + coordinates + state + every Jacobian nonzero written once + defect"""
    return 4 * nsh * n_elem + 8 * dim * n_node + 8 * n_dof * n_timepoints + 8 * nnz + 8 * n_dof


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """samples SM clocks / throttle reasons while the timed region runs: an NVML polling thread (5 ms period, starts
    instantly, so even a 100 ms timed region gets samples); falls back to `nvidia-smi -lms` when NVML is unavailable"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index=0):
        self.index, self.samples, self.proc = index, [], None
        self.nvml, self.handle, self.run, self.thread, self.sm, self.mx, self.reasons = None, None, False, None, [], None, set()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(self.handle, pynvml.NVML_CLOCK_SM)      # priming query (discarded)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll(self):
        nv = self.nvml
        while self.run:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for bit, name in self.BITS:
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nvml is not None:
            self.run = True
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.run = False
            self.thread.join(timeout=1.0)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi"}


def build_problem(n, rank=0, world=1):
    """the rank's block of the global hex mesh + state (config 3). Returns dict."""
    from plugin_navierstokes_b200 import meshgen
    if world == 1:
        coords, conn = meshgen.hex_grid(n, n, n)
        u = meshgen.state_vortex3d(coords, seed=3)
        return dict(coords=coords, conn=conn, u=u, iface=None)
    from plugin_navierstokes_b200 import partition
    return partition.block_problem(n, rank, world)


def cpu_baseline(n_sample, threads, sweeps=1):
    """the CPU oracle ("port" of the reference's element routines, UG4-like loop: separate Jacobian and
    defect sweeps, each rebuilding geometry + stabilisation) on a bounded sample of the same workload"""
    from oracle import oracle as ora
    from plugin_navierstokes_b200 import meshgen
    coords, conn = meshgen.hex_grid(n_sample, n_sample, n_sample)
    u = meshgen.state_vortex3d(coords, seed=3)
    p = ora.make_params(elem="hex", upwind=UPWIND, stab=STAB, kin_visc=VISC)
    rowptr, colind = ora.fv1_csr(ora.HEX, conn, coords.shape[0])
    vals, dfc = np.zeros(colind.size), np.zeros(rowptr.size - 1)
    best = None
    for _ in range(sweeps):
        vals[:] = 0
        dfc[:] = 0
        t0 = time.perf_counter()
        ora.assemble(p, conn, coords, u, rowptr, colind, ora.JAC_A, nthreads=threads, values=vals, defect=dfc)
        ora.assemble(p, conn, coords, u, rowptr, colind, ora.DEF_A, nthreads=threads, values=vals, defect=dfc)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return conn.shape[0] / best, conn.shape[0], best


def run_reference(args):
    """--impl reference: the reference's CPU path. The UG4 plugin cannot be compiled here (ugcore absent),
    so this is the oracle port, all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_sample = args.ref_n
    rates = []
    for i in range(args.warmup + args.steps):
        r, ne, dt = cpu_baseline(n_sample, threads)
        if i >= args.warmup:
            rates.append((r, dt))
    value = float(np.mean([r for r, _ in rates]))
    ms = float(np.mean([dt for _, dt in rates])) * 1e3
    sample = "hex %d^3 (%d elements) of config 3, Jacobian sweep + defect sweep, %d OpenMP threads" % (n_sample, n_sample ** 3, threads)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "config3: FV1 hex, LPS upwind + FIELDS/RAW, nu=1e-2, Jacobian+defect (A part); each step = a bounded "
                               "sample of %d^3 = %d elements of the %d^3-per-GPU workload (a per-element rate)" % (n_sample, n_sample ** 3, args.n),
                   "sample_cells": n_sample, "sample_elements": n_sample ** 3,
                   "reference_kind": "oracle port of the UG4 element routines (UG4 itself not buildable: ugcore absent)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cells", dest="n", type=int, default=184, help="hex cells per direction per GPU (config 3: 184)")
    ap.add_argument("--mode", default="gather", choices=["gather", "colored", "atomic"])
    ap.add_argument("--ref-n", type=int, default=96, help="cells per direction of the CPU sample (96^3 = 0.88 M elements, ~2 s per sweep pair on 16 cores)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-nccl-priority", action="store_true", help="N > 1: default-priority NCCL stream")
    ap.add_argument("--overlap", action="store_true", help="N > 1: force the interface-first split (default from 4 ranks on)")
    ap.add_argument("--no-overlap", action="store_true", help="N > 1: exchange after the whole assembly instead of behind the interior rows")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e-full", action="store_true", help="skip the full-matrix D2H variant of the end-to-end leg")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the owner-row check against the single-domain oracle")
    ap.add_argument("--parity-n", type=int, default=6, help="N > 1: cells per direction per rank of the parity problem")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import plugin_navierstokes_b200 as pkg
    from plugin_navierstokes_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        pgo = None
        if not args.no_nccl_priority:
            # the NCCL stream gets a high priority: its few CTAs are scheduled ahead of the pending blocks of the interior rows
            # kernel, so the transfers really run behind the assembly (the persistent rows kernel otherwise fills every SM first)
            try:
                pgo = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            except Exception:
                pgo = None
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=pgo)
    dev = torch.device("cuda", local)

    prob = build_problem(args.n, rank, world)
    coords, conn, u = prob["coords"], prob["conn"], prob["u"]
    disc = pkg.NavierStokes("u,v,w,p", "Inner", "fv1", device=local)
    disc.set_kinematic_viscosity(VISC)
    disc.set_upwind(UPWIND)
    disc.set_stabilization(STAB)
    disc.set_grid("hex", conn, coords)
    disc.prep_elem_loop()
    mode = {"gather": capi.SCATTER_GATHER, "colored": capi.SCATTER_COLORED, "atomic": capi.SCATTER_ATOMIC}[args.mode]
    what = capi.JAC_A | capi.DEF_A
    n_elem, n_node = conn.shape[0], coords.shape[0]
    nnz, n_dof = disc.nnz, disc.num_dofs
    abytes = algorithmic_bytes(n_elem, n_node, 8, 3, n_dof, nnz)

    disc.use_stream(torch.cuda.current_stream().cuda_stream)
    ud = torch.from_numpy(np.ascontiguousarray(u.reshape(-1))).to(dev)
    vals = torch.empty(nnz, dtype=torch.float64, device=dev)
    dfc = torch.empty(n_dof, dtype=torch.float64, device=dev)
    exch = None
    overlap = False
    if world > 1:
        from plugin_navierstokes_b200 import partition
        exch = partition.InterfaceExchange(disc, prob["iface"], dev)
        # measured on 8 x B200 (profiles/r2_bench_8gpu*.json): 20.57 ms overlapped vs 21.00 ms; at 2 GPUs (one face per rank) the extra
        # launch costs more than the hidden transfer (20.41 vs 20.26 ms), so the split is used from 4 ranks on
        overlap = (not args.no_overlap) and (world >= 4 or args.overlap)
        if overlap:
            exch.enable_overlap()          # interface rows first; the exchange then runs behind the interior rows

    def step():
        if exch is None:
            disc.assemble(what, ud, values=vals, defect=dfc, scatter_mode=mode)
        elif overlap:
            disc.assemble(what | capi.PHASE_PRIORITY, ud, values=vals, defect=dfc, scatter_mode=mode)
            works = exch.start_sum_to_owner(vals, dfc)
            disc.assemble(what | capi.PHASE_REST, ud, values=vals, defect=dfc, scatter_mode=mode)
            exch.finish_sum_to_owner(works, vals, dfc)
        else:
            disc.assemble(what, ud, values=vals, defect=dfc, scatter_mode=mode)
            exch.sum_to_owner(vals, dfc)

    sampler = ClockSampler(local)           # NVML initialised and primed before the warm-up: the first query of a process is slow
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    disc.check_errors()
    l0 = disc.launch_count + (exch.launches if exch else 0)
    if world > 1:
        dist.barrier()
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 2)]
    kernel_ms = []
    ev[0].record()
    for i in range(args.steps):
        ev[2 * i + 1].record()
        if exch is not None and overlap:
            step()                          # assembly and exchange interleaved: the pair of events brackets both
            ev[2 * i + 2].record()
        else:
            disc.assemble(what, ud, values=vals, defect=dfc, scatter_mode=mode)
            ev[2 * i + 2].record()
            if exch is not None:
                exch.sum_to_owner(vals, dfc)
    ev[-1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    disc.check_errors()
    total_ms = ev[0].elapsed_time(ev[-1])
    kernel_ms = [ev[2 * i + 1].elapsed_time(ev[2 * i + 2]) for i in range(args.steps)]
    launches = disc.launch_count + (exch.launches if exch else 0) - l0
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    ne = torch.tensor([float(n_elem)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(ne, op=dist.ReduceOp.SUM)
    total_ms = float(t.item())
    total_elems = float(ne.item())
    ms_per_step = total_ms / args.steps
    value = total_elems / (ms_per_step * 1e-3)

    setup_s = disc.query(capi.Q_SETUP_SECONDS)
    dev_bytes = disc.query(capi.Q_DEVICE_BYTES)

    # ---- N > 1: numerics of the multi-GPU path, checked after the timed region on a small block problem: owner rows (matrix AND
    # defect, interface rows included) after the NCCL interface summation against the single-domain CPU oracle ----
    parity_maxrel = None
    if world > 1 and not args.no_parity:
        from plugin_navierstokes_b200 import partition
        from oracle import oracle as ora            # the checker, outside every timed region
        pn = args.parity_n
        pp = partition.block_problem(pn, rank, world)
        pd = pkg.NavierStokes("u,v,w,p", "Inner", "fv1", device=local)
        pd.set_kinematic_viscosity(VISC); pd.set_upwind(UPWIND); pd.set_stabilization(STAB)
        pd.set_grid("hex", pp["conn"], pp["coords"])
        pu = torch.from_numpy(np.ascontiguousarray(pp["u"].reshape(-1))).to(dev)
        pv, pdf = pd.assemble(what, pu, scatter_mode=mode)
        pex = partition.InterfaceExchange(pd, pp["iface"], dev)
        pex.sum_to_owner(pv, pdf)
        torch.cuda.synchronize()
        pd.check_errors()
        gc, gconn, gu = partition.block_problem_global(pn, world)
        prm = ora.make_params(elem="hex", upwind=UPWIND, stab=STAB, kin_visc=VISC)
        grp, gci = ora.fv1_csr(ora.HEX, gconn, gc.shape[0])
        gv, gd = ora.assemble(prm, gconn, gc, gu.reshape(-1), grp, gci, ora.JAC_A | ora.DEF_A, nthreads=max(1, (os.cpu_count() or 1) // world))
        lrp, lci = pd.csr()
        em, ed = partition.owner_rows_error(lrp, lci, pv.cpu().numpy(), pdf.cpu().numpy(), pp["iface"]["l2g"], pex.owner, rank, grp, gci, gv, gd, 4)
        pt = torch.tensor([em, ed], dtype=torch.float64, device=dev)
        dist.all_reduce(pt, op=dist.ReduceOp.MAX)
        parity_maxrel = {"matrix_rows": float(pt[0].item()), "defect": float(pt[1].item()),
                         "problem": "hex %d^3 per rank, %d ranks, owner rows after sum_to_owner vs the single-domain oracle, relative to the largest entry" % (pn, world)}
        pd.close()
        del pv, pdf, pu

    # ---- end-to-end through the public API with HOST buffers (pinned), copies inside the timed region ----
    # e2e: the GPU-resident hand-off (nsb_assemble_resident + nsb_apply_jacobian): u goes up, the Jacobian stays on the device
    #      where its consumer runs (matrix-vector product of the solver), defect and J*x come back.
    # e2e_full_matrix: nsb_assemble(NSB_HOST), every CSR value returned to the host (a CPU solver's hand-off).
    e2e, e2e_full = None, None
    if not args.no_e2e:
        try:
            hu = torch.from_numpy(np.ascontiguousarray(u.reshape(-1))).pin_memory()
            hd = torch.empty(n_dof, dtype=torch.float64).pin_memory()
            hx = torch.from_numpy(np.random.default_rng(1).uniform(-1, 1, n_dof)).pin_memory()
            hy = torch.empty(n_dof, dtype=torch.float64).pin_memory()
            del vals
            torch.cuda.empty_cache()

            def timed(fn, steps):
                fn()                                                   # warm-up, allocates staging
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(steps):
                    fn()
                torch.cuda.synchronize()
                tt = torch.tensor([(time.perf_counter() - t0) / steps], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                return float(tt.item())

            hu_n, hd_n, hx_n, hy_n = hu.numpy(), hd.numpy(), hx.numpy(), hy.numpy()

            def resident_step():
                # NSB_HOST_ASYNC: x goes up while the assembly runs, the defect comes back while the product runs
                disc.assemble_resident(what, hu_n, defect=hd_n, scatter_mode=mode, asynchronous=True)
                disc.apply_jacobian(hx_n, y=hy_n, asynchronous=True)
                disc.synchronize()
                disc.check_errors()

            dt = timed(resident_step, max(args.e2e_steps, 3))
            e2e = {"value": total_elems / dt, "unit": UNIT, "h2d_bytes_per_step": int(8 * 2 * n_dof), "d2h_bytes_per_step": int(8 * 2 * n_dof),
                   "ms_per_step": dt * 1e3,
                   "note": "nsb_assemble_resident(NSB_HOST_ASYNC) + nsb_apply_jacobian(NSB_HOST_ASYNC) + nsb_synchronize: u and x from pinned "
                           "host memory, the CSR values stay on the device (GPU-resident hand-off, SURVEY 8f-2), defect and J*x returned to "
                           "pinned host memory; the copies run on the context's copy streams beside the kernels"}
            if not args.no_e2e_full:
                hv = torch.empty(nnz, dtype=torch.float64).pin_memory()
                dtf = timed(lambda: disc.assemble(what, hu.numpy(), values=hv.numpy(), defect=hd.numpy(), scatter_mode=mode), args.e2e_steps)
                e2e_full = {"value": total_elems / dtf, "unit": UNIT, "h2d_bytes_per_step": int(8 * n_dof), "d2h_bytes_per_step": int(8 * (nnz + n_dof)),
                            "ms_per_step": dtf * 1e3,
                            "note": "nsb_assemble(NSB_HOST): u from pinned host memory, CSR values + defect returned to pinned host memory (host-solver hand-off; PCIe-bound)"}
        except Exception as ex:       # noqa: BLE001
            e2e = e2e or {"value": None, "unit": UNIT, "error": str(ex)[:200]}

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        kms = float(np.mean(kernel_ms))
        achieved = abytes / (kms * 1e-3) / 1e9
        cpu = None
        if not args.no_cpu:
            threads = os.cpu_count() or 1
            r, ne_s, dt = cpu_baseline(args.ref_n, threads, sweeps=5)
            n1 = max(8, args.ref_n // 3)
            r1, ne_1, dt1 = cpu_baseline(n1, 1, sweeps=2)             # the UG4-like single-thread loop
            cpu = {"value": r, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": "hex %d^3 (%d elements) of the same workload, Jacobian sweep + defect sweep, best of 5 (%.1f s each)" % (args.ref_n, ne_s, dt),
                   "single_core": {"value": r1, "unit": UNIT, "cores": 1, "sample": "hex %d^3 (%d elements), best of 2 (%.1f s each)" % (n1, ne_1, dt1)}}
        # DRAM traffic of one pass: from the ncu capture of THIS kernel code (profiles/traffic.json carries the digest of the
        # CUDA sources it was taken with, written by tools/ncu_summary.py); a stale capture is not reported
        traffic, share, traffic_note = None, None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                import build as _build
                tj = json.load(open(tp))
                if tj.get("sources_digest") == _build._sources_digest():
                    bpe = tj["bytes_per_element"].get(args.mode)
                    traffic = bpe * n_elem if bpe else None          # per launch (= pass), scaled from the ncu capture
                    share = tj.get("kernel_share_ncu") if args.mode == "gather" else None
                    traffic_note = tj.get("note")
                else:
                    traffic_note = "profiles/traffic.json was captured with other kernel sources (stale): not reported"
            except Exception:
                traffic = None
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "config3: FV1 hex %d^3 per GPU (%d elements/GPU), LPS upwind + FIELDS/RAW, nu=1e-2, "
                                   "Jacobian+defect (A part) into global CSR" % (args.n, n_elem),
                       "scatter": args.mode, "l2": "inputs+outputs (%.1f GB) exceed L2, no flush needed" % (abytes / 1e9),
                       "algorithmic_bytes_per_element": abytes / n_elem, "nnz": int(nnz), "colors": disc.num_colors,
                       "setup_s": setup_s, "device_bytes_resident": int(dev_bytes),
                       "setup_note": "host preprocessing + table upload of nsb_upload_mesh, outside the timed region; the static Jacobian part J0 is built by the first pass"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel_ms": kms,
                         "kernel": "fv1 %s (all launches of one assembly pass: fv1_flux_kernel + fv1_rows_owner_kernel)" % args.mode,
                         "kernel_share_ncu": share, "traffic_note": traffic_note},
            "cpu_baseline": cpu, "e2e": e2e, "e2e_full_matrix": e2e_full, "gpu_launches": int(launches), "clocks": clocks,
        }
        if parity_maxrel is not None:
            out["parity_maxrel"] = parity_maxrel
        if world > 1:
            out["config"]["interface_exchange"] = ("overlapped: interface rows first (NSB_PHASE_PRIORITY), NCCL p2p behind the interior rows"
                                                   if overlap else "after the assembly pass")
            out["config"]["exchange_bytes_per_rank"] = exch.bytes_per_exchange()
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
