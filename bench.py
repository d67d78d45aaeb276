#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json): CG iters/s and steps/s on the ~1M-tet
corotational cantilever (config C2: RegularGridTopology 33x33x161 -> 175 329 nodes, 983 040 tetrahedra,
method=large, Vec3f, EulerImplicitSolver rayleigh 0.1/0.1, CGLinearSolver 25 iterations forced by tol 1e-9).

    python bench.py --gpus N --steps K --warmup W            our arm (sm_100a CUDA through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...   the reference's CPU algorithm (oracle) on the host cores

A "step" is one EulerImplicitSolver::solve: addForce + right-hand side + 25 CG iterations + integration.
value   = CG iterations per second with x, v resident in HBM (CUDA events around K steps, max over ranks)
e2e     = the same metric through the host-buffer entry point (H2D of x,v and D2H of x,v inside every step)
roofline= A*p element-pass kernel: algorithmic bytes per launch / event-timed launch duration vs measured HBM peak
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

CG_ITERS = 25
WORKLOADS = {
    "C2": dict(n=(33, 33, 161), mn=(0, 0, 0), mx=(4, 4, 20), box=(-1, -1, -1, 5, 5, 1e-6)),
    "C5": dict(n=(129, 129, 161), mn=(0, 0, 0), mx=(16, 16, 20), box=(-1, -1, -1, 17, 17, 1e-6)),
    "C1": dict(n=(5, 5, 20), mn=(-5, -5, 0), mx=(5, 5, 40), box=(-6, -6, -1, 50, 6, 0.1)),     # the reference's example scene (1824 tets): pure latency
    "SMALL": dict(n=(17, 17, 41), mn=(0, 0, 0), mx=(4, 4, 10), box=(-1, -1, -1, 5, 5, 1e-6)),
    "L2FIT": dict(n=(33, 33, 41), mn=(0, 0, 0), mx=(4, 4, 5), box=(-1, -1, -1, 5, 5, 1e-6)),   # 245 760 tets: the element records fit in L2
}
SCENE = dict(young=1000.0, poisson=0.3, density=1.0, gravity=(0.0, -9.0, 0.0), dt=0.01, rK=0.1, rM=0.1,
             iterations=CG_ITERS, tolerance=1e-9, threshold=1e-9, method="large")


def build_mesh(name):
    from sofa_b200 import topology as T
    w = WORKLOADS[name]
    pos, hexas = T.regular_grid(w["n"], w["mn"], w["mx"])
    tets = T.hexas_to_tetras(hexas, w["n"], "mapping_swapping")
    fixed = T.box_roi(pos, w["box"])
    return pos, tets, fixed


def algorithmic_bytes(T, N, s):
    """compulsory traffic. Per CG iteration: T*(16+20s) + N*34s; the A*p element-pass launch alone:
    element stream T*(16+20s) + nodal p read 3s + q write 3s + mass read s per node."""
    return dict(cg_iteration=T * (16 + 20 * s) + N * 34 * s, element_pass=T * (16 + 20 * s) + N * 7 * s)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled from a thread every few milliseconds
    (nvidia-smi in loop mode as the fallback)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device, self.proc, self.path, self.thread, self.stop_flag = device, None, None, None, False
        self.sm, self.mx, self.reasons = [], [], set()

    def _poll(self):
        import pynvml as N
        bits = {"hw_slowdown": N.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": N.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": N.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": N.nvmlClocksThrottleReasonSwPowerCap}
        while True:
            try:
                self.sm.append(float(N.nvmlDeviceGetClockInfo(self.h, N.NVML_CLOCK_SM)))
                r = N.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, b in bits.items():
                    if r & b:
                        self.reasons.add(k)
            except Exception:
                pass
            if self.stop_flag:
                return
            time.sleep(0.002)

    def start(self):
        try:
            import threading
            import pynvml as N
            N.nvmlInit()
            idx = self.device
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[self.device])
                except Exception:
                    pass
            self.h = N.nvmlDeviceGetHandleByIndex(idx)
            self.mx = [float(N.nvmlDeviceGetMaxClockInfo(self.h, N.NVML_CLOCK_SM))]
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv"); os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.thread:
            self.stop_flag = True
            self.thread.join(timeout=2)
            if self.sm:
                out.update(sm_mhz=float(np.median(self.sm)), sm_max_mhz=float(max(self.mx)), reasons=sorted(self.reasons), samples=len(self.sm), source="nvml")
            return out
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm), source="nvidia-smi")
        return out


def dist_env():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def cpu_baseline(workload, steps, threads):
    """The oracle (CPU restatement of the reference path) timed on the host cores on a bounded sample of the workload."""
    import oracle_lib as O
    pos, tets, fixed = build_mesh(workload)
    s = O.OracleScene(np.float32, pos)
    s.set_params(gravity=SCENE["gravity"], dt=SCENE["dt"], rayleighStiffness=SCENE["rK"], rayleighMass=SCENE["rM"],
                 iterations=SCENE["iterations"], tolerance=SCENE["tolerance"], threshold=SCENE["threshold"])
    s.set_mass_density(SCENE["density"], tets); s.set_tets(tets, SCENE["method"], SCENE["young"], SCENE["poisson"]); s.set_fixed(fixed)
    s.set_threads(threads)
    s.step()  # warm-up (also first-touch of the buffers)
    t0 = time.perf_counter(); iters = 0
    for _ in range(steps):
        iters += min(s.step(), CG_ITERS)
    dt = time.perf_counter() - t0
    return dict(value=iters / dt, unit="cg_iters/s", cores=threads, kind="port", steps_per_s=steps / dt,
                sample=f"{steps} EulerImplicit steps ({iters} CG iterations) of workload {workload} "
                       f"({tets.shape[0]} tets, Vec3f), oracle -O3 no-fma, {'sequential' if threads == 1 else 'ParallelTetrahedronFEMForceField-style addDForce'}")


def run_reference(args):
    """--impl reference: the reference's CPU algorithm for the path on this box's host cores.  SOFA itself cannot be
    built offline (Boost/Eigen/TinyXML2 absent), so this is the oracle port with the MultiThreading plugin's parallel
    addDForce on all cores, as DESIGN.md explains."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    steps = max(1, min(args.steps, 3))
    b = cpu_baseline(args.workload, steps, cores)
    seq = cpu_baseline(args.workload, 1, 1)
    pos, tets, fixed = build_mesh(args.workload)
    line = {"impl": "reference", "metric": "cg_iters_per_s", "value": b["value"], "unit": "cg_iters/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": 1, "ms_per_step": 1000.0 / b["steps_per_s"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "steps_per_s": b["steps_per_s"],
            "config": {"workload": f"{args.workload}: RegularGridTopology {WORKLOADS[args.workload]['n']} cantilever, {tets.shape[0]} tetrahedra, "
                                   f"{pos.shape[0]} nodes, method=large, {CG_ITERS} CG it/step", "cpu": _cpu_model()},
            "cpu_baseline": dict(b, sequential_1core=seq["value"]),
            "e2e": {"value": b["value"], "unit": "cg_iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def _cpu_model():
    try:
        for l in open("/proc/cpuinfo"):
            if l.startswith("model name"):
                return l.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_ours(args):
    import torch
    import sofa_b200 as sb
    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- sofa_b200 has no CPU path (use --impl reference for the CPU arm)")
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dtype, template, s = (np.float32, "B200Vec3f", 4) if args.dtype == "f32" else (np.float64, "B200Vec3d", 8)

    if world > 1:
        try:
            return run_ours_distributed(args, rank, world, local, dtype, template, s)
        finally:
            import sys
            sys.stdout.flush()
            dist.destroy_process_group()
    # ---- synthetic input of BASELINE's shape
    pos, tets, fixed = build_mesh(args.workload)
    ctx = sb.Context(local)
    mo = sb.MechanicalObject(ctx, template, position=pos)
    ff = sb.TetrahedronFEMForceField(mo, tets, youngModulus=SCENE["young"], poissonRatio=SCENE["poisson"], method=SCENE["method"], tileElems=args.tile)
    mass = sb.DiagonalMass(mo, tets, massDensity=SCENE["density"])
    node = sb.SolverNode(mo, ff, mass, sb.FixedProjectiveConstraint(mo, fixed), dt=SCENE["dt"], gravity=SCENE["gravity"],
                         rayleighStiffness=SCENE["rK"], rayleighMass=SCENE["rM"], iterations=SCENE["iterations"],
                         tolerance=SCENE["tolerance"], threshold=SCENE["threshold"])
    T, N = tets.shape[0], pos.shape[0]
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm: K steps replayed from the step's CUDA graph, CUDA events on the launching stream
    for _ in range(args.warmup):
        node.step()
    barrier()
    sampler = ClockSampler(local); sampler.start()
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        node.step()
    e1.record(stream)
    barrier()
    launches = ctx.launch_count - launches0
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    info = node.last_solve()
    iters_per_step = min(info["iterations"], CG_ITERS)
    # ---- per-kernel durations: the same K steps again with an event pair recorded around every launch (plain launches,
    # the graph is bypassed while the profiler is on); used for roofline.achieved of the dominant kernel
    ctx.profile_begin()
    for _ in range(args.steps):
        node.step()
    prof = ctx.profile_end()

    # ---- end-to-end arm: host (pinned) state vectors through sofab200_node_step_host
    xh = torch.from_numpy(pos.astype(dtype)).pin_memory(); vh = torch.zeros_like(xh).pin_memory()
    for _ in range(min(args.warmup, 3)):
        node.step_host(xh, vh)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 20))
    for _ in range(e2e_steps):
        node.step_host(xh, vh)
    barrier()
    e2e_s = time.perf_counter() - t0

    if world > 1:
        t = torch.tensor([ms, e2e_s], dtype=torch.float64, device=ctx.device)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1])
    if rank != 0:
        return

    total_iters = iters_per_step * args.steps * world
    value = total_iters / (ms * 1e-3)
    peak, peak_src = measured_peaks()
    ab = algorithmic_bytes(T, N, s)
    # dominant kernel: the persistent CG kernel (one launch = the whole CGLinearSolver loop of a step); when the mesh does not fit
    # it, the A*p element pass of the multi-kernel loop
    persistent = prof.get("cg_persistent", {}).get("launches", 0) > 0
    ep = prof["cg_persistent"] if persistent else prof["element_pass_dforce"]
    k_ms = ep["ms"] / max(ep["launches"], 1)
    k_bytes = ab["cg_iteration"] * iters_per_step if persistent else ab["element_pass"]
    k_name = (f"tet_cg_persistent_kernel (CGLinearSolver loop, {iters_per_step} iterations per launch: A*p element pass, shared-node sums, "
              "x/r/p updates, both dot products)") if persistent else "tet_tile_kernel<DF_COROT> (A*p element pass)"
    achieved = k_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(f"{args.workload}_{args.dtype}_{'cg_persistent' if persistent else 'element_pass'}_bytes")
    cg_gbs = ab["cg_iteration"] * total_iters / world / (ms * 1e-3) / 1e9
    line = {
        "metric": "cg_iters_per_s", "value": value, "unit": "cg_iters/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "steps_per_s": args.steps * world / (ms * 1e-3), "cg_iters_per_step": iters_per_step,
        "config": {"workload": f"{args.workload}: RegularGridTopology {WORKLOADS[args.workload]['n']} cantilever, {T} tetrahedra, {N} nodes per GPU, "
                               f"TetrahedronFEMForceField method=large E=1000 nu=0.3, EulerImplicit rayleigh 0.1/0.1, CG {CG_ITERS} it (tol 1e-9)",
                   "partition": "one beam per GPU" if world > 1 else "single GPU", "l2": "working set per CG iteration exceeds the 126 MB L2 "
                   f"({(T * 120 + N * 34 * s) / 1e6:.0f} MB streamed)", "layout": ff.stats()},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "kernel": k_name, "algorithmic_bytes_per_launch": k_bytes,
                     "avg_launch_ms": k_ms, "launches_timed": ep["launches"], "peak_source": peak_src,
                     "cg_loop": {"algorithmic_bytes_per_iteration": ab["cg_iteration"], "achieved": cg_gbs, "frac": cg_gbs / peak,
                                 "note": "whole step time attributed to the CG iterations (includes addForce, RHS, integration)"},
                     "kernel_ms": {k: v for k, v in prof.items()}},
        "e2e": {"value": iters_per_step * e2e_steps * world / e2e_s, "unit": "cg_iters/s", "h2d_bytes_per_step": 2 * N * 3 * s, "d2h_bytes_per_step": 2 * N * 3 * s,
                "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps},
        "gpu_launches": launches, "clocks": clocks,
    }
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(args.workload, args.cpu_steps, 1)
    print(json.dumps(line))


def run_ours_distributed(args, rank, world, local, dtype, template, s):
    """N > 1: ONE beam N times as long as the workload's, cut into N z-slabs of the workload's size (weak scaling), with the
    per-iteration halo exchange and the two scalar allreduces over NCCL (sofa_b200/parallel.py)."""
    import torch
    import torch.distributed as dist
    import sofa_b200 as sb
    import sofa_b200.parallel as PAR
    from sofa_b200 import topology as T
    w = WORKLOADS[args.workload]
    nz = (w["n"][2] - 1) * world + 1
    n = (w["n"][0], w["n"][1], nz)
    mx = (w["mx"][0], w["mx"][1], w["mn"][2] + (w["mx"][2] - w["mn"][2]) * world)
    pos, hexas = T.regular_grid(n, w["mn"], mx)
    tets = T.hexas_to_tetras(hexas, n, "mapping_swapping")
    fixed = T.box_roi(pos, (w["box"][0], w["box"][1], w["box"][2], mx[0] + 1, mx[1] + 1, w["box"][5]))
    ctx = sb.Context(local)
    node = PAR.DistributedSolverNode(pos, tets, fixed, SCENE["density"], SCENE["young"], SCENE["poisson"], SCENE["method"], ctx=ctx, template=template,
                                     dt=SCENE["dt"], gravity=SCENE["gravity"], rayleighStiffness=SCENE["rK"], rayleighMass=SCENE["rM"],
                                     iterations=SCENE["iterations"], tolerance=SCENE["tolerance"], threshold=SCENE["threshold"])
    stream = torch.cuda.current_stream()

    def barrier():
        dist.barrier(); torch.cuda.synchronize()
    for _ in range(args.warmup):
        node.step()
    barrier()
    sampler = ClockSampler(local); sampler.start()
    launches0 = ctx.launch_count
    node.cg_iterations_total = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        node.step()
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    # ---- end-to-end arm: every rank's local (x, v) in pinned host memory, copied in and out every step (sofab200_node_step_host)
    xh = node.be.x.detach().cpu().pin_memory(); vh = node.be.v.detach().cpu().pin_memory()
    for _ in range(3):
        node.be.node.step_host(xh, vh)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 20))
    for _ in range(e2e_steps):
        node.be.node.step_host(xh, vh)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e0.elapsed_time(e1), e2e_s], dtype=torch.float64, device=ctx.device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_s = float(t[0]), float(t[1])
    if rank != 0:
        return
    T_loc, N_loc = node.rm.elems.shape[0], node.rm.n_local
    info = node.be.node.last_solve()
    if info["end_condition"] == 99:
        raise RuntimeError("a cross-GPU wait timed out inside the CG kernel (end_condition 99): the numbers of this run are void")
    iters = min(info["iterations"], CG_ITERS) * args.steps     # every step of this workload runs the same forced iteration count
    peer = bool(getattr(node.be, "peer", False))
    exchange = ("ONE persistent CG kernel per GPU; interface partial sums stored into the neighbours' mailboxes over NVLink (CUDA IPC peer memory), "
                "flag-sequenced halo sync + rank-ordered all-reduce inside the kernel, no NCCL call in the loop") if peer else \
               "multi-kernel loop, NCCL send/recv + allreduce enqueued by the library, device-resident CG scalars"
    peak, peak_src = measured_peaks()
    ab = algorithmic_bytes(T_loc, N_loc, s)
    value = iters * world / (ms * 1e-3)    # every CG iteration processes `world` partitions of the workload's size
    cg_gbs = ab["cg_iteration"] * iters / (ms * 1e-3) / 1e9
    line = {"metric": "cg_iters_per_s", "value": value, "unit": "cg_iters/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "steps_per_s": args.steps / (ms * 1e-3), "cg_iters_per_step": iters / args.steps,
            "config": {"workload": f"{args.workload} x {world}: ONE RegularGridTopology {n} cantilever, {tets.shape[0]} tetrahedra, {pos.shape[0]} nodes, "
                                   f"z-slab partition ({T_loc} tets, {N_loc} nodes per GPU), method=large, CG {CG_ITERS} it", "partition": f"{world} slabs, halo "
                       f"{len(node.rm.interface)} nodes/rank; {exchange}", "l2": "working set per CG iteration exceeds L2",
                       "value_counts": f"weak scaling: one CG iteration of the {world}x longer beam = {world} partitions of the N=1 workload's size, counted as {world} units"},
            "roofline": {"bound": "hbm", "achieved": cg_gbs, "peak": peak, "unit": "GB/s", "frac": cg_gbs / peak, "traffic": None,
                         "kernel": "whole distributed CG iteration per GPU (algorithmic bytes of one partition)", "peak_source": peak_src},
            "e2e": {"value": min(info["iterations"], CG_ITERS) * e2e_steps * world / e2e_s, "unit": "cg_iters/s", "h2d_bytes_per_step": 2 * N_loc * 3 * s * world,
                    "d2h_bytes_per_step": 2 * N_loc * 3 * s * world, "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                    "note": "every rank copies its partition's x, v from pinned host memory and back each step (sofab200_node_step_host); wall clock, max over ranks"},
            "gpu_launches": launches, "clocks": clocks}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=list(WORKLOADS))
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--tile", type=int, default=0, help="elements per CTA tile (0 = library default)")
    ap.add_argument("--cpu-steps", type=int, default=6, help="oracle steps timed for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
