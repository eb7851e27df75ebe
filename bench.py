#!/usr/bin/env python
"""bench.py -- MLUPS of the B200 D2Q9-BGK solver on the reference's headline workload.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (CUDA, through the C ABI)
    python bench.py --impl reference [--gpus N] --steps K --warmup W  # the reference's CPU path

Workload (BASELINE.json configs[3]): synthetic 16384 x 16384 channel -- solid rows 0 and
ny-1, 1 % random solids (splitmix64, seed 1234), density 0.1, accel 0.005, omega 1.85 -- STRONG scaling
over N GPUs (row slabs, halo rows stored straight into the neighbour GPU's memory over NVLink).
A "step" is one lattice timestep = one launch of the fused sweep kernel per GPU.
`--workload weak32768` runs BASELINE.json configs[4] instead (32768 x 32768 per GPU, weak scaling).

One JSON line on stdout (rank 0):
  value       MLUPS = nx*ny*K / t / 1e6, t = CUDA-event time of the K launches, max over ranks, lattice
              resident in HBM (9.66 GB per buffer >> 126 MB L2, so no flush is needed between steps)
  e2e         the same metric for one whole job through the C ABI with pinned HOST buffers:
              lbm_upload (AoS cells + obstacle bytes) -> lbm_run(K) -> lbm_download (cells + av_vels),
              wall clock between barriers; h2d/d2h bytes are per step (job bytes / K)
  roofline    achieved = 72 B/cell * cells per launch / average launch time, against the measured copy
              bandwidth in MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (a port of the reference's main/LastChance.cpp loop, OpenMP over rows)
              timed on this box's host cores on a bounded sample
Under torchrun (N > 1) every rank runs this file; rank r drives GPU LOCAL_RANK.
"""
import argparse
import ctypes
import importlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_CELL = 72  # 9 fp32 read + 9 fp32 written per lattice update
DENSITY, ACCEL, OMEGA, REYNOLDS_DIM = 0.1, 0.005, 1.85, 10
SEED, SOLID_PROB = 1234, 0.01


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--workload", choices=["channel16384", "weak32768"], default="channel16384")
    ap.add_argument("--nx", type=int, default=0, help="override the grid width (development only)")
    ap.add_argument("--ny", type=int, default=0, help="override the grid height (development only)")
    ap.add_argument("--local-slabs", type=int, default=1,
                    help="development only: cut the grid into this many row slabs that all live on GPU 0")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU time budget of the cpu_baseline sample")
    return ap.parse_args()


def grid_for(args, world):
    if args.workload == "weak32768":
        nx, ny, scaling, name = 32768, 32768 * world, "weak", "synthetic channel 32768x32768 per GPU"
    else:
        nx, ny, scaling, name = 16384, 16384, "strong", "synthetic channel 16384x16384"
    if args.nx:
        nx = args.nx
    if args.ny:
        ny = args.ny
    if args.nx or args.ny:
        name = f"synthetic channel {nx}x{ny} (override)"
    return nx, ny, scaling, name


# ---------------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi during the timed region
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), reasons=sorted(reasons),
                       samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port (OpenMP over rows) on a bounded sample of the same workload
# ---------------------------------------------------------------------------------------------------
def cpu_baseline(nx, ny, budget_s, lbm_mod):
    from oracle import oracle

    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        pass
    # same generator and parameters; a full-width horizontal band keeps the sample's memory bounded
    rows = min(ny, 2048)
    op = oracle.make_params(nx, rows, 1, REYNOLDS_DIM, DENSITY, ACCEL, OMEGA)
    obst = lbm_mod.synthetic_obstacles(nx, rows, SEED, SOLID_PROB)
    a = oracle.init_cells(op)
    b = np.empty_like(a)
    oracle.accelerate_flow(op, a, obst)
    oracle.timestep_rows(op, a, b, obst, 0, rows, oracle.ACCUM_DOUBLE, cores)  # warm-up / page faults
    a, b = b, a
    t0 = time.perf_counter()
    steps = 0
    while True:
        oracle.timestep_rows(op, a, b, obst, 0, rows, oracle.ACCUM_DOUBLE, cores)
        a, b = b, a
        steps += 1
        dt = time.perf_counter() - t0
        if dt > budget_s or steps >= 200:
            break
    mlups = nx * rows * steps / dt / 1e6
    return {"value": round(mlups, 2), "unit": "MLUPS", "cores": cores, "kind": "port",
            "sample": f"{steps} timesteps of a {nx}x{rows} band of the same synthetic channel "
                      f"(oracle/lbm_oracle.c = main/LastChance.cpp:185-267 restated, gcc -O3 -march=x86-64-v3 "
                      f"-ffp-contract=off, OpenMP over rows, fp64 tot_u), {dt:.1f} s"}


def reference_program_baseline():
    """The UNMODIFIED reference program (oracle/_ref/compareLbm_fma = main/LastChance.cpp built with the reference's
    own flags, -O3 and FMA contraction allowed) on the reference's own CPU-runnable case, BASELINE.json configs[0]:
    input_128x128, 40 000 timesteps, single-threaded as shipped.  None when oracle/_ref was never built."""
    exe = os.path.join(ROOT, "oracle", "_ref", "compareLbm_fma")
    gold = os.path.join(ROOT, "tests", "golden", "params")
    if not os.path.exists(exe):
        return None
    with tempfile.TemporaryDirectory() as tmp:
        t0 = time.perf_counter()
        r = subprocess.run([exe, os.path.join(gold, "input_128x128.params"), os.path.join(gold, "obstacles_128x128.dat")],
                           cwd=tmp, capture_output=True, text=True)
        dt = time.perf_counter() - t0
        if r.returncode != 0 or not os.path.exists(os.path.join(tmp, "av_vels.dat")):
            return None
    return {"value": round(128 * 128 * 40000 / dt / 1e6, 2), "unit": "MLUPS", "cores": 1, "kind": "reference",
            "sample": f"unmodified main/LastChance.cpp (g++ -O3 -march=x86-64-v3), params/input_128x128.params, 40000 timesteps "
                      f"incl. its file output, {dt:.1f} s wall"}


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  The unmodified
    program (oracle/_ref/compareLbm_*) always writes final_state.dat (25 GB of text at 16384^2) and is
    single-threaded, so the headline grid is timed with the oracle port of the same loop on all cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # only the pure-numpy problem generator is shared with our arm; none of the CUDA product is loaded here
    import importlib.util
    spec = importlib.util.spec_from_file_location("lbm_problem", os.path.join(ROOT, "lbm-graphcore_b200", "problem.py"))
    lbm_mod = importlib.util.module_from_spec(spec)
    sys.modules["lbm_problem"] = lbm_mod
    spec.loader.exec_module(lbm_mod)
    nx, ny, scaling, name = grid_for(args, max(args.gpus, 1))
    # budget: K+W "steps" of a bounded band, a few minutes at most
    base = cpu_baseline(nx, ny, min(120.0, max(10.0, 0.5 * (args.steps + args.warmup))), lbm_mod)
    line = {
        "impl": "reference", "metric": "MLUPS", "value": base["value"], "unit": "MLUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(nx * ny / (base["value"] * 1e6) * 1e3, 3), "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": name, "nx": nx, "ny": ny, "density": DENSITY, "accel": ACCEL, "omega": OMEGA, "seed": SEED,
                   "solid_prob": SOLID_PROB},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != max(args.gpus, 1):
        if world == 1 and args.gpus > 1:
            # convenience: re-launch ourselves under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
            sys.exit(subprocess.call(cmd))
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")

    lbm = importlib.import_module("lbm-graphcore_b200")  # raises if liblbm_b200.so is missing: no fallback
    slabs = importlib.import_module("lbm-graphcore_b200.slabs")
    if lbm.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device visible; the solver has no CPU path")

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="cpu:gloo,cuda:nccl", init_method="env://")

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    nx, ny, scaling, name = grid_for(args, world)
    K, W = args.steps, args.warmup
    p = lbm.Params(nx, ny, W + K, REYNOLDS_DIM, DENSITY, ACCEL, OMEGA)

    if world == 1:
        solver = lbm.Solver(p, devices=[0] * args.local_slabs) if args.local_slabs > 1 else lbm.Solver(p, num_gpus=1)
    else:
        solver = lbm.Solver(p, rank=rank, world=world, device=local_rank)
        slabs.connect_ring(solver, dist)  # CUDA-IPC handles all-gathered, ring neighbours mapped

    # ---- device-resident timing ----------------------------------------------------------------
    solver.init_synthetic(SEED, SOLID_PROB)
    barrier()
    # nvidia-smi needs ~0.1 s to deliver its first sample: start it before the warm-up steps (same kernel, same load)
    # so that short timed regions are covered too
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.2)
    solver.run(W)
    barrier()
    launches0 = solver.launch_count
    t0 = time.perf_counter()
    solver.run(K)  # blocks; CUDA events bracket the K launches on the launching stream
    wall = time.perf_counter() - t0
    ms = max_over_ranks(solver.elapsed_ms)
    wall = max_over_ranks(wall)
    clocks = sampler.stop() if rank == 0 else None
    launches = solver.launch_count - launches0
    barrier()
    av_last = float(slabs.combine_av_vels(dist, solver.download_tot_u(), solver.free_cells)[-1])  # rank order
    if not np.isfinite(av_last):
        raise SystemExit("bench.py: av_vels is not finite")
    cells = nx * ny
    mlups = cells * K / (ms * 1e-3) / 1e6
    variant = solver.kernel_variant
    solver.close()

    # ---- end-to-end through host buffers -------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        rows0, rows = slabs.slab_rows(rank, world, ny)
        p2 = lbm.Params(nx, ny, K, REYNOLDS_DIM, DENSITY, ACCEL, OMEGA)
        nbytes_cells = rows * nx * 9 * 4
        nbytes_obst = rows * nx
        hc = lbm.lib.lbm_host_alloc(nbytes_cells)
        ho = lbm.lib.lbm_host_alloc(nbytes_obst)
        if not hc or not ho:
            raise SystemExit("bench.py: pinned host allocation failed")
        cells_h = np.ctypeslib.as_array(ctypes.cast(hc, ctypes.POINTER(ctypes.c_float)), shape=(rows, nx, 9))
        obst_h = np.ctypeslib.as_array(ctypes.cast(ho, ctypes.POINTER(ctypes.c_uint8)), shape=(rows, nx))
        init = lbm.initial_cells(lbm.Params(1, 1, 1, 1, DENSITY, ACCEL, OMEGA))[0, 0]
        cells_h[...] = init
        band = 1024
        for r in range(0, rows, band):
            n = min(band, rows - r)
            obst_h[r:r + n] = lbm.synthetic_obstacles(nx, ny, SEED, SOLID_PROB, rows0 + r, n)
        av_h = np.zeros(K, dtype=np.float32)
        if world == 1:
            s2 = lbm.Solver(p2, num_gpus=1)
        else:
            s2 = lbm.Solver(p2, rank=rank, world=world, device=local_rank)
            slabs.connect_ring(s2, dist)
        barrier()
        t0 = time.perf_counter()
        s2.upload_raw(hc, ho)
        barrier()  # every rank's lattice is in place before anyone pushes halo rows
        s2.run(K)
        if world == 1:
            s2.download_raw(hc, av_h.ctypes.data)
        else:
            s2.download_raw(hc, None)
            part = s2.download_tot_u()
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        if world == 1 and not np.isfinite(av_h[-1]):
            raise SystemExit("bench.py: e2e av_vels is not finite")
        e2e = {"value": round(cells * K / dt / 1e6, 1), "unit": "MLUPS",
               "h2d_bytes_per_step": int((nbytes_cells + nbytes_obst) * world / K),
               "d2h_bytes_per_step": int((nbytes_cells * world + 8 * K) / K),
               "job": f"lbm_upload + lbm_run({K}) + lbm_download, pinned host AoS buffers, {dt:.3f} s wall"}
        s2.close()
        lbm.lib.lbm_host_free(hc)
        lbm.lib.lbm_host_free(ho)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline ------------------------------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    launch_ms = ms / K  # one sweep launch per step per GPU
    achieved = (cells / world) * BYTES_PER_CELL / (launch_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "sweep_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            if tj.get("cells_per_launch") == cells // world:
                traffic = tj.get("dram_bytes_per_launch")
        except (OSError, ValueError):
            pass
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "kernel": f"lbm::sweep_kernel<SWEEP_{variant.upper()},{'true' if world > 1 or args.local_slabs > 1 else 'false'}>",
                "bytes_per_cell": BYTES_PER_CELL, "cells_per_launch": cells // world,
                "avg_launch_ms": round(launch_ms, 4), "frac_of_nominal_8TBs": round(achieved / 8000.0, 4),
                "context": "best plain copy measured on this pool with tools/probes/stream_probe.cu: 6920 GB/s "
                           "(profiles/r01c_stream_probe_4.txt); ncu reports 85 % DRAM throughput for this kernel"}

    line = {
        "metric": "MLUPS", "value": round(mlups, 1), "unit": "MLUPS", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": round(ms / K, 4), "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": name, "nx": nx, "ny": ny, "density": DENSITY, "accel": ACCEL, "omega": OMEGA,
                   "seed": SEED, "solid_prob": SOLID_PROB, "partition": f"{world} row slab(s)",
                   "l2": (f"working set (2 x {cells * 36 / 1e9 / world:.2f} GB lattices per GPU) >> 126 MB L2: no flush needed "
                          "between steps" if cells * 72 / world > 4 * 126e6 else
                          "development override: the lattices fit in L2 (latency-bound small grid, not the benchmark)"),
                   "kernel_variant": variant, "av_vels_last": av_last,
                   "wall_ms_per_step": round(wall * 1e3 / K, 4)},
        "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches * world), "clocks": clocks,
    }
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(nx, ny, args.cpu_seconds, lbm)
        ref = reference_program_baseline()
        if ref is not None:
            line["cpu_baseline"]["reference_program"] = ref
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
