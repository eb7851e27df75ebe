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
              timed on this box's host cores on a bounded sample (N = 1 only)
  parity      (outside the timed region) a perturbed random problem run on the N GPUs of this launch, compared
              with the same problem as one slab on rank 0's GPU and with the CPU oracle: lattice SHA-256 per
              rank, av_vels in ulps.  A mismatch makes the run exit non-zero.
  weak        BASELINE.json configs[4] -- 32768 x 32768 cells PER GPU -- measured in the same launch
  shipped     (N = 1) BASELINE.json configs[0-2]: the reference's 128^2 / 256^2 / 1024^2 inputs for their full
              maxIters, microseconds per step, MLUPS, check.py's verdict and the lattice hash against the oracle
  detail      free-form facts about this run (partition, kernel variant, last av_vels, wall clock per step)
The reference arm times the oracle port (OpenMP over rows, all host cores) on the SAME grid for W + K real timesteps.
Under torchrun (N > 1) every rank runs this file; rank r drives GPU LOCAL_RANK.
"""
import argparse
import ctypes
import gzip
import hashlib
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
GOLDEN = os.path.join(ROOT, "tests", "golden")


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
    ap.add_argument("--in-process", type=int, default=0, metavar="N",
                    help="comparison runs: ONE process drives N GPUs (lbm_create / lbm_create_grid) instead of one rank per GPU")
    ap.add_argument("--grid", default="", metavar="PXxPY",
                    help="with --in-process: cut the grid into PX x PY blocks with ghost columns instead of N row slabs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-weak", action="store_true")
    ap.add_argument("--no-shipped", action="store_true")
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


def workload_config(name, nx, ny):
    """The `config` object: the workload only, so that both arms print the same one."""
    lattice_gb = nx * ny * 36 / 1e9
    if nx * ny * 72 > 8 * 4 * 126e6:
        l2 = (f"working set (2 lattices of {lattice_gb:.2f} GB, split evenly over the GPUs of the run) >> 126 MB L2 per "
              "GPU: no flush needed between steps")
    else:
        l2 = "development override: the lattices fit in L2 (latency-bound small grid, not the benchmark)"
    return {"workload": name, "nx": nx, "ny": ny, "density": DENSITY, "accel": ACCEL, "omega": OMEGA, "seed": SEED,
            "solid_prob": SOLID_PROB, "l2": l2}


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
# CPU side: the oracle port (OpenMP over rows).  cpu_baseline = a bounded band of the workload (our arm's reported
# baseline); the reference arm = the whole grid for W + K real timesteps.
# ---------------------------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def load_problem_module():
    """lbm-graphcore_b200/problem.py on its own (pure numpy): the reference arm loads none of the CUDA product."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("lbm_problem", os.path.join(ROOT, "lbm-graphcore_b200", "problem.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["lbm_problem"] = mod
    spec.loader.exec_module(mod)
    return mod


ORACLE_BUILD = ("oracle/lbm_oracle.c = main/LastChance.cpp:185-267 restated, gcc -O3 -march=x86-64-v3 -ffp-contract=off, "
                "OpenMP over rows, fp64 tot_u")


def cpu_channel(nx, ny, rows, problem_mod):
    """Initial state of the first `rows` rows of the nx x ny synthetic channel (all of it when rows == ny)."""
    from oracle import oracle
    op = oracle.make_params(nx, rows, 1, REYNOLDS_DIM, DENSITY, ACCEL, OMEGA)
    obst = np.empty((rows, nx), dtype=np.uint8)
    for r in range(0, rows, 1024):
        n = min(1024, rows - r)
        obst[r:r + n] = problem_mod.synthetic_obstacles(nx, ny if rows == ny else rows, SEED, SOLID_PROB, r, n)
    a = oracle.init_cells(op)
    b = np.empty_like(a)
    oracle.accelerate_flow(op, a, obst)
    return oracle, op, obst, a, b


def cpu_baseline(nx, ny, budget_s, problem_mod):
    cores = host_cores()
    rows = min(ny, 2048)  # a full-width horizontal band keeps the sample's memory and time bounded
    oracle, op, obst, a, b = cpu_channel(nx, ny, rows, problem_mod)
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
            "sample": f"{steps} timesteps of a {nx}x{rows} band of the same synthetic channel ({ORACLE_BUILD}), {dt:.1f} s"}


def reference_program_baseline():
    """The UNMODIFIED reference program (oracle/_ref/compareLbm_fma = main/LastChance.cpp built with the reference's
    own flags, -O3 and FMA contraction allowed) on the reference's own CPU-runnable case, BASELINE.json configs[0]:
    input_128x128, 40 000 timesteps, single-threaded as shipped.  None when oracle/_ref was never built."""
    exe = os.path.join(ROOT, "oracle", "_ref", "compareLbm_fma")
    gold = os.path.join(GOLDEN, "params")
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


def mem_available_bytes():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) * 1024
    except OSError:
        pass
    return None


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path (main/LastChance.cpp:185-267, restated in
    oracle/lbm_oracle.c) on all host cores, on the SAME grid as our arm, for W untimed + K timed real timesteps.
    The unmodified program (oracle/_ref/compareLbm_*) is single-threaded and always writes final_state.dat (25 GB of text
    at 16384^2), so the headline grid is timed with the port of the same loop; the unmodified program's own timing on
    its 128x128 case rides along as cpu_baseline.reference_program."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    problem_mod = load_problem_module()
    nx, ny, scaling, name = grid_for(args, max(args.gpus, 1))
    K, W = args.steps, args.warmup
    cores = host_cores()
    need = nx * ny * (2 * 36 + 1) * 1.05
    avail = mem_available_bytes()
    rows = ny
    if avail is not None and need > 0.9 * avail:  # a host too small for the whole grid: the tallest band that fits
        rows = max(64, int(ny * 0.9 * avail / need) // 64 * 64)
    oracle, op, obst, a, b = cpu_channel(nx, ny, rows, problem_mod)
    tot = []
    for _ in range(W):
        oracle.timestep_rows(op, a, b, obst, 0, rows, oracle.ACCUM_DOUBLE, cores)
        a, b = b, a
    t0 = time.perf_counter()
    for _ in range(K):
        tot.append(oracle.timestep_rows(op, a, b, obst, 0, rows, oracle.ACCUM_DOUBLE, cores))
        a, b = b, a
    dt = time.perf_counter() - t0
    if not np.all(np.isfinite(tot)):
        raise SystemExit("bench.py: the reference arm produced a non-finite av_vels")
    mlups = nx * rows * K / dt / 1e6
    whole = rows == ny
    sample = (f"{W} + {K} timesteps of the whole {nx}x{ny} grid" if whole else
              f"{W} + {K} timesteps of a {nx}x{rows} band (the host has {avail / 1e9:.0f} GB free; the whole grid needs "
              f"{need / 1e9:.0f} GB)") + f" ({ORACLE_BUILD}), {dt:.1f} s timed"
    base = {"value": round(mlups, 2), "unit": "MLUPS", "cores": cores, "kind": "port", "sample": sample}
    ref = reference_program_baseline() if not (args.nx or args.ny) else None
    if ref is not None:
        base["reference_program"] = ref
    line = {
        "impl": "reference", "metric": "MLUPS", "value": base["value"], "unit": "MLUPS", "n_gpus": args.gpus,
        "steps": K, "warmup": W, "ms_per_step": round(dt * 1e3 / K * (ny / rows), 3), "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(name, nx, ny),
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "detail": {"rows_timed": rows, "timed_s": round(dt, 6), "tot_u_last": float(tot[-1]) if tot else None},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# driver-visible correctness and the other BASELINE.json configurations (all outside the timed region)
# ---------------------------------------------------------------------------------------------------
def ulp_distance(a, b):
    ai = np.ascontiguousarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    bi = np.ascontiguousarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    return int(np.abs(ai - bi).max()) if ai.size else 0


def parity_record(lbm, slabs, dist, rank, world, local_rank):
    """A problem whose halo rows matter from the first step (5 % random solids, every population perturbed) on the N
    GPUs of this launch -- one rank per GPU, halo rows stored into the neighbour's memory by the sweep kernel -- against
    (a) the same problem as ONE slab on rank 0's GPU and (b) the CPU oracle (the checker, never the thing measured).
    With one GPU the multi-slab kernel is exercised by two slabs that share the GPU."""
    nx, ny, iters, seed = 2048, 96 * max(world, 2), 40, 4242
    p, cells, obst = lbm.perturbed_problem(nx, ny, iters, seed)
    if world > 1:
        s = lbm.Solver(p, rank=rank, world=world, device=local_rank)
        slabs.connect_ring(s, dist)
        r0, n = s.row_begin, s.row_count
        s.upload(cells[r0:r0 + n], obst[r0:r0 + n])
        dist.barrier()
        s.run(iters)
        got, _ = s.download(av_vels=False)
        av = slabs.combine_av_vels(dist, s.download_tot_u(), s.free_cells)
        variant = s.kernel_variant
        slabs.close_ring(s, dist)
        mine = (r0, n, hashlib.sha256(got.tobytes()).hexdigest())
        parts = [None] * world
        dist.all_gather_object(parts, mine)
    else:
        with lbm.Solver(p, devices=[0, 0]) as s:
            s.upload(cells, obst)
            s.run(iters)
            got, av = s.download()
            variant = s.kernel_variant
            parts = [(sl["row_begin"], sl["row_count"],
                      hashlib.sha256(got[sl["row_begin"]:sl["row_begin"] + sl["row_count"]].tobytes()).hexdigest())
                     for sl in s.slabs]
    if rank != 0:
        return None
    from oracle import oracle
    op = oracle.make_params(nx, ny, iters, 10, 0.1, 0.005, 1.85)
    want, want_av = oracle.run(op, cells, obst, iters, accum=oracle.ACCUM_DOUBLE, threads=host_cores())
    halo_equal = all(hashlib.sha256(want[r0:r0 + n].tobytes()).hexdigest() == h for r0, n, h in parts)
    with lbm.Solver(p, num_gpus=1) as s1:
        s1.upload(cells, obst)
        s1.run(iters)
        one, one_av = s1.download()
    rec = {"problem": f"perturbed {nx}x{ny}, 5 % random solids, {iters} steps, seed {seed}",
           "partition": f"{world} rank(s), one row slab per GPU" if world > 1 else "2 row slabs sharing the one GPU",
           "kernel_variant": variant,
           "halo_lattice_sha_equal": bool(halo_equal),
           "one_slab_lattice_sha_equal": hashlib.sha256(one.tobytes()).hexdigest() == hashlib.sha256(want.tobytes()).hexdigest(),
           "lattice_sha256": hashlib.sha256(want.tobytes()).hexdigest()[:16],
           "av_vels_max_ulp": ulp_distance(av, want_av), "one_slab_av_vels_max_ulp": ulp_distance(one_av, want_av),
           "vs": "oracle"}
    rec["ok"] = bool(rec["halo_lattice_sha_equal"] and rec["one_slab_lattice_sha_equal"] and rec["av_vels_max_ulp"] <= 1 and
                     rec["one_slab_av_vels_max_ulp"] <= 1)
    return rec


def shipped_records(lbm):
    """BASELINE.json configs[0-2]: the reference's shipped inputs for their full maxIters on one B200, validated the way
    the reference validates itself (check/check.py:84-100,134-147: av_vels and final-state pressure within 1 % of its
    goldens; 256^2 and 1024^2 ship no final_state golden, there the oracle fixture's pressure field stands in) and,
    more strictly, by the SHA-256 of the final lattice against the oracle fixture."""
    from tools import check as check_tool
    out = []
    for case in ("128x128", "256x256", "1024x1024"):
        pfile = os.path.join(GOLDEN, "params", f"input_{case}.params")
        p = lbm.Params.from_file(pfile)
        obst = lbm.load_obstacles(p.nx, p.ny, os.path.join(GOLDEN, "params", f"obstacles_{case}.dat"))
        fx = np.load(os.path.join(GOLDEN, f"oracle_{case}.npz"))
        with lbm.Solver(p) as s:
            s.upload(lbm.initial_cells(p), obst)
            s.run(p.max_iters)
            ms = s.elapsed_ms
            cells, av = s.download()
            fields = s.download_fields()
            variant = s.kernel_variant
        with gzip.open(os.path.join(GOLDEN, "check", f"{case}.av_vels.dat.gz"), "rt") as f:
            av_ref = np.loadtxt(f, usecols=[1])
        d_av = check_tool.get_diff_values(av_ref, av.astype(np.float64))
        ok = bool(np.isfinite(d_av["max_diff_pcnt"]) and abs(d_av["max_diff_pcnt"]) <= 1.0)
        fs_path = os.path.join(GOLDEN, "check", f"{case}.final_state.dat.gz")
        if os.path.exists(fs_path):
            with gzip.open(fs_path, "rt") as f:
                p_ref = np.loadtxt(f, usecols=[5])
            p_src = "reference golden check/%s.final_state.dat" % case
            p_sim = fields[..., 3].astype(np.float64).reshape(-1)
        else:
            st = int(fx["stride"])
            p_ref = fx["pressure"].astype(np.float64).reshape(-1)
            p_src = "oracle fixture (the reference ships no final_state golden for this size)"
            p_sim = fields[::st, ::st, 3].astype(np.float64).reshape(-1)
        d_fs = check_tool.get_diff_values(p_ref, p_sim)
        ok = ok and bool(np.isfinite(d_fs["max_diff_pcnt"]) and abs(d_fs["max_diff_pcnt"]) <= 1.0)
        sha_ok = hashlib.sha256(cells.tobytes()).hexdigest() == str(fx["cells_sha256"])
        out.append({"config": f"params/input_{case}.params + obstacles_{case}.dat", "steps": p.max_iters,
                    "us_per_step": round(ms * 1e3 / p.max_iters, 3),
                    "mlups": round(p.nx * p.ny * p.max_iters / (ms * 1e-3) / 1e6, 1), "device_s": round(ms / 1e3, 4),
                    "kernel_variant": variant, "check_py": "pass" if ok else "FAIL",
                    "av_vels_max_diff_pcnt": float(abs(d_av["max_diff_pcnt"])),
                    "pressure_max_diff_pcnt": float(abs(d_fs["max_diff_pcnt"])), "pressure_reference": p_src,
                    "lattice_sha256_equals_oracle": bool(sha_ok),
                    "av_vels_max_ulp_vs_oracle": ulp_distance(av, fx["av_f64sum"])})
    return out


def weak_record(lbm, slabs, dist, rank, world, local_rank, peak, max_over_ranks):
    """BASELINE.json configs[4]: 32768 x 32768 cells per GPU (77 GB of lattices each), row slabs, weak scaling."""
    nx, nyl, K, W = 32768, 32768, 10, 3
    p = lbm.Params(nx, nyl * world, W + K, REYNOLDS_DIM, DENSITY, ACCEL, OMEGA)
    if world == 1:
        s = lbm.Solver(p, num_gpus=1)
    else:
        s = lbm.Solver(p, rank=rank, world=world, device=local_rank)
        slabs.connect_ring(s, dist)
    s.init_synthetic(SEED, SOLID_PROB)
    if dist is not None:
        dist.barrier()
    s.run(W)
    if dist is not None:
        dist.barrier()
    s.run(K)
    ms = max_over_ranks(s.elapsed_ms)
    av_last = float(slabs.combine_av_vels(dist, s.download_tot_u(), s.free_cells)[-1])
    slabs.close_ring(s, dist)
    mlups = nx * nyl * world * K / (ms * 1e-3) / 1e6
    gbs = nx * nyl * BYTES_PER_CELL / (ms / K * 1e-3) / 1e9
    return {"workload": "synthetic channel 32768x32768 per GPU (BASELINE.json configs[4])", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": round(ms / K, 4), "mlups": round(mlups, 1), "per_gpu_gbs": round(gbs, 1),
            "frac_of_measured_peak": round(gbs / peak, 4), "av_vels_last": av_last, "finite": bool(np.isfinite(av_last))}


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

    inproc = args.in_process if world == 1 else 0
    if inproc:  # comparison of partitions: one process, N GPUs, row slabs or PX x PY blocks
        if args.grid:
            gx, gy = (int(v) for v in args.grid.lower().split("x"))
            if gx * gy != inproc:
                raise SystemExit(f"--grid {args.grid} needs --in-process {gx * gy}")
            solver = lbm.Solver(p, grid=(gx, gy))
        else:
            solver = lbm.Solver(p, num_gpus=inproc)
        args.no_e2e = args.no_parity = args.no_weak = args.no_shipped = args.no_cpu_baseline = True
    elif world == 1:
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
    slabs.close_ring(solver, dist)

    # ---- end-to-end through host buffers -------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        rows0, rows = slabs.slab_rows(rank, world, ny)
        p2 = lbm.Params(nx, ny, K, REYNOLDS_DIM, DENSITY, ACCEL, OMEGA)
        nbytes_cells = rows * nx * 9 * 4
        nbytes_obst = rows * nx
        hc = lbm.host_alloc(nbytes_cells, local_rank)
        ho = lbm.host_alloc(nbytes_obst, local_rank)
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
        t_up = time.perf_counter() - t0
        barrier()  # every rank's lattice is in place before anyone pushes halo rows (protocol of include/lbm_b200.h)
        s2.run(K)
        t1 = time.perf_counter()
        if world == 1:
            s2.download_raw(hc, av_h.ctypes.data)
        else:
            s2.download_raw(hc, None)
            part = s2.download_tot_u()
        s2_free = s2.free_cells
        t_down = time.perf_counter() - t1
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        t_up, t_down = max_over_ranks(t_up), max_over_ranks(t_down)
        if world == 1 and not np.isfinite(av_h[-1]):
            raise SystemExit("bench.py: e2e av_vels is not finite")
        e2e = {"value": round(cells * K / dt / 1e6, 1), "unit": "MLUPS",
               "h2d_bytes_per_step": int((nbytes_cells + nbytes_obst) * world / K),
               "d2h_bytes_per_step": int((nbytes_cells * world + 8 * K) / K),
               "job": f"lbm_upload + lbm_run({K}) + lbm_download, pinned host AoS buffers, {dt:.3f} s wall",
               "upload_s": round(t_up, 4), "download_s": round(t_down, 4),
               "h2d_gbs_per_gpu": round((nbytes_cells + nbytes_obst) / t_up / 1e9, 1),
               "d2h_gbs_per_gpu": round(nbytes_cells / t_down / 1e9, 1)}
        slabs.close_ring(s2, dist)
        if world == 1 and nbytes_cells <= 12e9:  # (a second pinned lattice-sized host buffer: not for the 38 GB weak-scaling grid)
            # The same job as ONE call, lbm_pipeline_solve: the grid is solved band by band (a band plus K + a few margin rows
            # is all that K steps of it depend on), so uploads, sweeps and downloads of different bands overlap and PCIe
            # carries data both ways at once.  Same result (checked below against the sequential job's), fewer seconds.
            seq = dict(e2e)
            out_ptr = lbm.host_alloc(nbytes_cells, local_rank)
            if not out_ptr:
                raise SystemExit("bench.py: pinned host allocation failed")
            out_h = np.ctypeslib.as_array(ctypes.cast(out_ptr, ctypes.POINTER(ctypes.c_float)), shape=(rows, nx, 9))
            final_h = np.ctypeslib.as_array(ctypes.cast(hc, ctypes.POINTER(ctypes.c_float)), shape=(rows, nx, 9))
            probe = [slice(0, 64), slice(ny // 2 - 32, ny // 2 + 32), slice(ny - 64, ny)]
            want = [final_h[sl].copy() for sl in probe]  # the sequential job's final state (it was downloaded into hc)
            cells_h[...] = init                          # ... which becomes the input again
            av2 = np.zeros(K, dtype=np.float32)
            pipe = lbm.Pipeline(p2, K, local_rank)  # like the Engine / the handle above: prepared outside the timed region
            info = pipe.info
            t0 = time.perf_counter()
            pipe.solve_raw(hc, ho, out_ptr, av2.ctypes.data)  # untimed warm-up job (first use of the output buffer by the DMA engines)
            dt_first = time.perf_counter() - t0
            t0 = time.perf_counter()
            pipe.solve_raw(hc, ho, out_ptr, av2.ctypes.data)
            dt2 = time.perf_counter() - t0
            pipe.close()
            same = all(np.array_equal(out_h[sl].view(np.uint32), w.view(np.uint32)) for sl, w in zip(probe, want))
            av_ulp = ulp_distance(av2, av_h)
            nbands, band_rows, margin = info["bands"], info["band_rows"], info["margin_rows"]
            banded = nbands > 1
            up_rows = ny + 2 * margin * nbands
            e2e = {"value": round(cells * K / dt2 / 1e6, 1), "unit": "MLUPS",
                   "h2d_bytes_per_step": int(up_rows * nx * 37 / K), "d2h_bytes_per_step": int((nbytes_cells + 8 * K * nbands) / K),
                   "job": (f"lbm_pipeline_solve: {nbands} bands of {band_rows} rows + 2 x {margin} margin rows, uploads / {K} steps / "
                           f"downloads of different bands overlapped, pinned host AoS buffers, {dt2:.3f} s wall") if banded else
                          f"lbm_pipeline_solve (one band: upload + run + download inside; the job is too long for bands to pay), {dt2:.3f} s wall",
                   "pcie_gbs_both_ways": round((up_rows * nx * 37 + nbytes_cells) / dt2 / 1e9, 1), "first_job_s": round(dt_first, 4),
                   "equals_sequential_job": {"lattice_rows_probed_bit_equal": bool(same), "av_vels_max_ulp": av_ulp},
                   "sequential_job": seq}
            if not same or av_ulp > 1:
                raise SystemExit("bench.py: lbm_pipeline_solve differs from upload + run + download: " + json.dumps(e2e))
            lbm.lib.lbm_host_free(out_ptr)
        if world > 1 and nbytes_cells <= 12e9 and rows % 8 == 0:
            # N ranks: every rank solves ITS slab as a window pipeline (lbm_pipeline_solve_window): its input holds the slab's
            # initial rows plus `margin` rows of its neighbours' either side (here generated; in an application exchanged once,
            # host to host), so the K steps of the slab depend on nothing else -- no halo traffic, no rank waits for another,
            # and each rank's PCIe link carries both directions at once.  Checked against the sequential job above (whose
            # result is in hc) before it replaces its number; any failure on any rank keeps the sequential number.
            seq = dict(e2e)
            ok, note, dt2, info, pipe, av_ulp = True, "", None, None, None, None
            bufs = []

            def everyone(flag):  # collective: did every rank get this far?  (no collective ever sits inside a try block)
                return max_over_ranks(0.0 if flag else 1.0) == 0.0

            try:  # phase 1, local: prepare the pipeline, the window buffers and one untimed warm-up job
                pipe = lbm.Pipeline(p2, K, local_rank, window=(rows0, rows))
                info = pipe.info
                M = info["margin_rows"]
                wrows = rows + 2 * M
                wc = lbm.host_alloc(wrows * nx * 36, local_rank)
                wo = lbm.host_alloc(wrows * nx, local_rank)
                out_ptr = lbm.host_alloc(nbytes_cells, local_rank)
                bufs = [q for q in (wc, wo, out_ptr) if q]
                if len(bufs) != 3:
                    raise MemoryError("pinned host allocation failed")
                wcells = np.ctypeslib.as_array(ctypes.cast(wc, ctypes.POINTER(ctypes.c_float)), shape=(wrows, nx, 9))
                wobst = np.ctypeslib.as_array(ctypes.cast(wo, ctypes.POINTER(ctypes.c_uint8)), shape=(wrows, nx))
                out_h = np.ctypeslib.as_array(ctypes.cast(out_ptr, ctypes.POINTER(ctypes.c_float)), shape=(rows, nx, 9))
                final_h = np.ctypeslib.as_array(ctypes.cast(hc, ctypes.POINTER(ctypes.c_float)), shape=(rows, nx, 9))
                wcells[...] = init
                for r in range(0, wrows, band):
                    n = min(band, wrows - r)
                    g0 = (rows0 - M + r) % ny
                    n1 = min(n, ny - g0)  # (a chunk may wrap around the top of the grid)
                    wobst[r:r + n1] = lbm.synthetic_obstacles(nx, ny, SEED, SOLID_PROB, g0, n1)
                    if n1 < n:
                        wobst[r + n1:r + n] = lbm.synthetic_obstacles(nx, ny, SEED, SOLID_PROB, 0, n - n1)
                probe = [slice(0, 64), slice(rows // 2 - 32, rows // 2 + 32), slice(rows - 64, rows)]
                want = [final_h[sl].copy() for sl in probe]
                pipe.solve_window_raw(wc, wo, out_ptr)  # untimed warm-up job
            except Exception as exc:  # noqa: BLE001 -- never lose the bench line over the optional path
                ok, note = False, f"{type(exc).__name__}: {exc}"
            all_ok = everyone(ok)
            if all_ok:  # phase 2: the timed job, between barriers
                tot2, free2 = None, 0
                barrier()
                t0 = time.perf_counter()
                try:
                    tot2, free2 = pipe.solve_window_raw(wc, wo, out_ptr)
                except Exception as exc:  # noqa: BLE001
                    ok, note = False, f"{type(exc).__name__}: {exc}"
                barrier()
                dt2 = max_over_ranks(time.perf_counter() - t0)
                all_ok = everyone(ok)
            if all_ok:  # phase 3: the same answer as the sequential job?
                same = all(np.array_equal(out_h[sl].view(np.uint32), w.view(np.uint32)) for sl, w in zip(probe, want))
                av2 = slabs.combine_av_vels(dist, tot2, free2)
                av1 = slabs.combine_av_vels(dist, part, s2_free)
                av_ulp = ulp_distance(av2, av1)
                if not same or av_ulp > 1:
                    ok, note = False, f"window pipeline differs from the sequential job on rank {rank} (rows equal: {same}, av_vels ulp {av_ulp})"
                all_ok = everyone(ok)
            if pipe is not None:
                pipe.close()
            for ptr in bufs:
                lbm.lib.lbm_host_free(ptr)
            if all_ok:
                up_rows = (rows + 2 * info["margin_rows"] * info["bands"]) * world
                e2e = {"value": round(cells * K / dt2 / 1e6, 1), "unit": "MLUPS",
                       "h2d_bytes_per_step": int(up_rows * nx * 37 / K), "d2h_bytes_per_step": int((nbytes_cells * world + 8 * K * world) / K),
                       "job": (f"lbm_pipeline_solve_window on every rank: {info['bands']} bands of {info['band_rows']} rows + 2 x "
                               f"{info['margin_rows']} margin rows per slab, no halo traffic, uploads / {K} steps / downloads overlapped, "
                               f"pinned host AoS buffers, {dt2:.3f} s wall"),
                       "pcie_gbs_both_ways_per_gpu": round((up_rows * nx * 37 / world + nbytes_cells) / dt2 / 1e9, 1),
                       "equals_sequential_job": {"lattice_rows_probed_bit_equal": True, "av_vels_max_ulp": av_ulp},
                       "sequential_job": seq}
            else:
                e2e["window_pipeline_skipped"] = note or "another rank could not run it"
        lbm.lib.lbm_host_free(hc)
        lbm.lib.lbm_host_free(ho)

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    # ---- driver-visible parity, weak scaling and the shipped inputs (outside the timed region) ----
    parity = None if args.no_parity else parity_record(lbm, slabs, dist, rank, world, local_rank)
    weak = None
    if not args.no_weak and args.workload != "weak32768" and not (args.nx or args.ny):
        weak = weak_record(lbm, slabs, dist, rank, world, local_rank, peak, max_over_ranks)
    shipped = None
    if not args.no_shipped and world == 1 and rank == 0 and not (args.nx or args.ny):
        shipped = shipped_records(lbm)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline ------------------------------------------------------------------------------
    launch_ms = ms / K  # one sweep launch per step per GPU
    gpus_used = inproc or world
    achieved = (cells / gpus_used) * BYTES_PER_CELL / (launch_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "sweep_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            for entry in tj.get("entries", [tj]):
                if entry.get("cells_per_launch") == cells // gpus_used and bool(entry.get("multi", False)) == (gpus_used > 1 or args.local_slabs > 1):
                    traffic = entry.get("dram_bytes_per_launch")
        except (OSError, ValueError):
            pass
    multi = gpus_used > 1 or args.local_slabs > 1 or bool(args.grid)
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "kernel": f"lbm::sweep_kernel<SWEEP_{variant.upper()},{'true' if multi else 'false'}>",
                "bytes_per_cell": BYTES_PER_CELL, "cells_per_launch": cells // gpus_used,
                "avg_launch_ms": round(launch_ms, 4), "frac_of_nominal_8TBs": round(achieved / 8000.0, 4),
                "context": "best plain copy measured on this pool with tools/probes/stream_probe.cu: 6920 GB/s "
                           "(profiles/r01c_stream_probe_4.txt); ncu reports 85 % DRAM throughput for this kernel"}

    line = {
        "metric": "MLUPS", "value": round(mlups, 1), "unit": "MLUPS", "n_gpus": gpus_used, "steps": K, "warmup": W,
        "ms_per_step": round(ms / K, 4), "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(name, nx, ny),
        "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches * world), "clocks": clocks,
        "detail": {"partition": (f"{args.grid} blocks with ghost columns, one per GPU, one process" if (inproc and args.grid) else
                                 f"{inproc} row slabs, one per GPU, one process" if inproc else
                                 f"{world} row slab(s), one per GPU" if args.local_slabs <= 1 else f"{args.local_slabs} row slabs on one GPU"),
                   "kernel_variant": variant, "av_vels_last": av_last, "wall_ms_per_step": round(wall * 1e3 / K, 4)},
    }
    if parity is not None:
        line["parity"] = parity
    if weak is not None:
        line["weak"] = weak
    if shipped is not None:
        line["shipped"] = shipped
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(nx, ny, args.cpu_seconds, lbm)
        ref = reference_program_baseline()
        if ref is not None:
            line["cpu_baseline"]["reference_program"] = ref
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    bad = []
    if parity is not None and not parity["ok"]:
        bad.append("parity")
    if shipped is not None and any(r["check_py"] != "pass" or not r["lattice_sha256_equals_oracle"] for r in shipped):
        bad.append("shipped")
    if weak is not None and not weak["finite"]:
        bad.append("weak")
    if bad:
        raise SystemExit(f"bench.py: FAILED checks: {', '.join(bad)} (see the JSON line)")


if __name__ == "__main__":
    main()
