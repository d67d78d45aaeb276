#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json): CG iters/s and steps/s on the ~1M-tet corotational
cantilever (config C2: RegularGridTopology 33x33x161 -> 175 329 nodes, 983 040 tetrahedra, method=large, Vec3f,
EulerImplicitSolver rayleigh 0.1/0.1, CGLinearSolver 25 iterations forced by tol 1e-9).

    python bench.py --gpus N --steps K --warmup W            our arm (sm_100a CUDA through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...   the reference's CPU algorithm (oracle) on the host cores
    python bench.py --workload C5 --gpus N                    config 5: the 15.7 M-tet beam cut into N slabs (strong scaling)

A "step" is one EulerImplicitSolver::solve: addForce + right-hand side + 25 CG iterations + integration.
value    = CG iterations per second with x, v resident in HBM (CUDA events around K steps, max over ranks)
e2e      = the same metric through the host-buffer entry points, pinned host memory, H2D and D2H inside every step.  N=1: the coupling of a device-resident
           state with a host loop (sofab200_node_step_pipelined: that step's external forces up, the new positions down on a copy stream under the next
           step) with the synchronous round trip of x (sofab200_node_step_host_x) beside it; N>1: the synchronous round trip on every rank
roofline = the CG kernel: algorithmic bytes per launch / event-timed launch duration vs the measured HBM peak
configs  = (N=1, default workload) the other single-GPU configurations of BASELINE.json, each with its own numbers: C3 (hexahedra, polar),
           C2 in Vec3d, C4 (liver-sized mesh: latency)
Prints ONE JSON line on rank 0.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

CG_ITERS = 25
GRID = dict(young=1000.0, poisson=0.3, density=1.0, gravity=(0.0, -9.0, 0.0), dt=0.01, rK=0.1, rM=0.1, iterations=CG_ITERS, tolerance=1e-9, threshold=1e-9)
WORKLOADS = {
    "C2": dict(kind="tet", n=(33, 33, 161), mn=(0, 0, 0), mx=(4, 4, 20), box=(-1, -1, -1, 5, 5, 1e-6), method="large", **GRID),
    "C3": dict(kind="hex", n=(65, 65, 121), mn=(0, 0, 0), mx=(8, 8, 15), box=(-1, -1, -1, 9, 9, 1e-6), method="polar", **GRID),
    # liver.scn as written (Demos/liver.scn: 181 nodes, 596 tetrahedra, TetrahedralCorotationalFEMForceField), rotations by polar decomposition
    "C4": dict(kind="liver", method="polar", young=3000.0, poisson=0.3, density=1.0, gravity=(0.0, -9.81, 0.0), dt=0.02, rK=0.1, rM=0.1,
               iterations=CG_ITERS, tolerance=1e-9, threshold=1e-9),
    # C2's mesh under FastTetrahedralCorotationalForceField (method "qr", the class's default): the CG loop runs over the EDGES
    "C2_FAST": dict(kind="fast", n=(33, 33, 161), mn=(0, 0, 0), mx=(4, 4, 20), box=(-1, -1, -1, 5, 5, 1e-6), method="qr", **GRID),
    "C5": dict(kind="tet", n=(129, 129, 161), mn=(0, 0, 0), mx=(16, 16, 20), box=(-1, -1, -1, 17, 17, 1e-6), method="large", **GRID),
    "C1": dict(kind="tet", n=(5, 5, 20), mn=(-5, -5, 0), mx=(5, 5, 40), box=(-6, -6, -1, 50, 6, 0.1), method="large", **GRID),   # the reference's example scene (1824 tets): pure latency
    "SMALL": dict(kind="tet", n=(17, 17, 41), mn=(0, 0, 0), mx=(4, 4, 10), box=(-1, -1, -1, 5, 5, 1e-6), method="large", **GRID),
    "C2_SMALL": dict(kind="tet", n=(9, 9, 41), mn=(0, 0, 0), mx=(4, 4, 20), box=(-1, -1, -1, 5, 5, 1e-6), method="large", **GRID),
    "L2FIT": dict(kind="tet", n=(33, 33, 41), mn=(0, 0, 0), mx=(4, 4, 5), box=(-1, -1, -1, 5, 5, 1e-6), method="large", **GRID),   # 245 760 tets: the element records fit in L2
}
SCENE = dict(GRID, method="large")     # (tools/ import this)


def build_mesh(name, stretch=1):
    """positions, elements, fixed indices of a workload; stretch = N makes the beam N times as long (weak scaling)."""
    from sofa_b200 import topology as T
    w = WORKLOADS[name]
    if w["kind"] == "liver":
        z = np.load(os.path.join(ROOT, "tests", "golden", "liver_mesh.npz"))
        return z["positions"], z["tetrahedra"], np.array([3, 39, 64], np.uint32)
    n = (w["n"][0], w["n"][1], (w["n"][2] - 1) * stretch + 1)
    mx = (w["mx"][0], w["mx"][1], w["mn"][2] + (w["mx"][2] - w["mn"][2]) * stretch)
    pos, hexas = T.regular_grid(n, w["mn"], mx)
    fixed = T.box_roi(pos, w["box"])
    if w["kind"] == "hex":
        return pos, hexas, fixed
    return pos, T.hexas_to_tetras(hexas, n, "mapping_swapping"), fixed


def workload_string(name, n_elems, n_nodes, dtype):
    """The same string in both arms (the driver compares them)."""
    w = WORKLOADS[name]
    what = {"tet": "tetrahedra, TetrahedronFEMForceField", "hex": "hexahedra, HexahedronFEMForceField", "liver": "tetrahedra (Demos/liver.scn mesh), TetrahedralCorotationalFEMForceField",
            "fast": "tetrahedra, FastTetrahedralCorotationalForceField"}[w["kind"]]
    grid = f"RegularGridTopology {w['n']} cantilever, " if "n" in w else ""
    return (f"{name}: {grid}{n_elems} {what} method={w['method']} E={w['young']:g} nu={w['poisson']:g}, {n_nodes} nodes, Vec3{'f' if dtype == 'f32' else 'd'}, "
            f"DiagonalMass, FixedProjectiveConstraint, EulerImplicit rayleigh {w['rK']:g}/{w['rM']:g} dt={w['dt']:g}, CG {w['iterations']} it (tol {w['tolerance']:g})")


def algorithmic_bytes(kind, E, N, s):
    """compulsory traffic per CG iteration.  Tetrahedra: E*(16+20s) + N*34s.  Hexahedra: SURVEY counts the 24x24 K_e per element
    (E*(32+309s)); on a regular grid every element has the SAME K_e, which the element record references instead of carrying, so what has to move
    is E*(32+9s) (indices + rotation) + N*34s -- both are reported, the roofline uses the second."""
    if kind == "hex":
        return dict(cg_iteration=E * (32 + 9 * s) + N * 34 * s, cg_iteration_survey=E * (32 + 309 * s) + N * 34 * s)
    if kind == "fast":      # E = number of EDGES: two indices + the edge's 3x3 matrix, plus the node streams
        return dict(cg_iteration=E * (8 + 9 * s) + N * 34 * s)
    return dict(cg_iteration=E * (16 + 20 * s) + N * 34 * s)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_source_sha():
    """Hash of the CUDA sources: profiles/traffic.json is only trusted for the code it was measured on."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "sofa_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".h")):
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def measured_traffic(key):
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None, "no ncu capture"
    d = json.load(open(p))
    if d.get("source_sha") != kernel_source_sha():
        return None, f"stale: profiles/traffic.json was measured on source {d.get('source_sha')}, this is {kernel_source_sha()} (tools/update_traffic.py)"
    return d.get(key), d.get("how", "")


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
            return self
        except Exception:
            self.thread = None
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv"); os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

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


# ---------------------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (restatement of the reference's CPU classes) on the host cores
# ---------------------------------------------------------------------------------------------------------------------------------
def oracle_scene(workload, dtype, threads):
    import oracle_lib as O
    w = WORKLOADS[workload]
    pos, elems, fixed = build_mesh(workload)
    s = O.OracleScene(np.float32 if dtype == "f32" else np.float64, pos)
    s.set_params(gravity=w["gravity"], dt=w["dt"], rayleighStiffness=w["rK"], rayleighMass=w["rM"], iterations=w["iterations"], tolerance=w["tolerance"], threshold=w["threshold"])
    s.set_mass_density(w["density"], elems)
    if w["kind"] == "hex":
        s.set_hexas(elems, w["method"], w["young"], w["poisson"])
    elif w["kind"] == "fast":
        s.set_fast_tets(elems, w["method"], w["young"], w["poisson"])
    else:
        s.set_tets(elems, w["method"], w["young"], w["poisson"])
        if w["kind"] == "liver":
            s.set_tetrahedral_corotational(True)
    s.set_fixed(fixed)
    s.set_threads(threads)
    return s, elems.shape[0], pos.shape[0]


def cpu_baseline(workload, dtype, steps, warmup, threads):
    """The oracle timed on a bounded sample of the workload: `warmup` untimed steps, then exactly `steps` timed ones."""
    s, E, N = oracle_scene(workload, dtype, threads)
    for _ in range(max(warmup, 0)):
        s.step()
    t0 = time.perf_counter(); iters = 0
    for _ in range(steps):
        iters += min(s.step(), WORKLOADS[workload]["iterations"])
    dt = time.perf_counter() - t0
    par = {"tet": "ParallelTetrahedronFEMForceField-style addDForce", "hex": "ParallelHexahedronFEMForceField-style addDForce", "liver": "ParallelTetrahedronFEMForceField-style addDForce",
           "fast": "sequential addDForce (the MultiThreading plugin has no parallel FastTetrahedralCorotationalForceField)"}[WORKLOADS[workload]["kind"]]
    return dict(value=iters / dt, unit="cg_iters/s", cores=threads, kind="port", steps_per_s=steps / dt,
                sample=f"{steps} EulerImplicit steps ({iters} CG iterations) of workload {workload} ({E} elements, Vec3{'f' if dtype == 'f32' else 'd'}) after {warmup} warm-up steps, "
                       f"oracle -O3 no-fma, {'sequential (the reference classes as they are)' if threads == 1 else par + f' on {threads} threads'}")


def unit_for(workload):
    """C5 is strong-scaled (one mesh, value = its CG iterations/s); every other workload counts partitions of the workload's size."""
    return "cg_iters/s" if workload == "C5" else "partition_cg_iters/s"


def run_reference(args):
    """--impl reference: the reference's CPU algorithm for the path on this box's host cores.  SOFA itself cannot be built offline
    (cmake project with Boost/Eigen/TinyXML2, none present), so this is the oracle port; the timed variant is the MultiThreading plugin's
    parallel addDForce (applications/plugins/MultiThreading/.../Parallel{Tetrahedron,Hexahedron}FEMForceField.inl) on every host core, the others
    are reported beside it: the sequential classes in Vec3f and Vec3d (SOFA's default SReal is double) and SOFA's default task-scheduler size
    hardware_concurrency()/2 (Sofa/framework/Simulation/Core/src/sofa/simulation/task/TaskScheduler.cpp:31-33)."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    wl = args.workload
    b = cpu_baseline(wl, args.dtype, args.steps, args.warmup, cores)
    variants = {"all_cores": dict(value=b["value"], cores=cores)}
    if not args.no_cpu_variants:
        half = max(1, cores // 2)
        v = cpu_baseline(wl, args.dtype, 2, 1, half); variants["sofa_default_task_scheduler_threads"] = dict(value=v["value"], cores=half)
        v = cpu_baseline(wl, "f32", 1, 1, 1); variants["sequential_Vec3f"] = dict(value=v["value"], cores=1)
        v = cpu_baseline(wl, "f64", 1, 1, 1); variants["sequential_Vec3d"] = dict(value=v["value"], cores=1)
    pos, elems, fixed = build_mesh(wl)
    line = {"impl": "reference", "metric": "cg_iters_per_s", "value": b["value"], "unit": unit_for(wl), "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 / b["steps_per_s"], "higher_is_better": True, "scaling": "strong" if wl == "C5" else "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic", "steps_per_s": b["steps_per_s"],
            "config": {"workload": workload_string(wl, elems.shape[0], pos.shape[0], args.dtype), "cpu": _cpu_model(),
                       "partition": "the CPU arm always runs ONE partition of the workload's size on the host cores of rank 0 (at N GPUs our arm runs N of them)"},
            "cpu_baseline": dict(b, variants=variants),
            "e2e": {"value": b["value"], "unit": unit_for(wl), "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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


# ---------------------------------------------------------------------------------------------------------------------------------
# our arm, one GPU
# ---------------------------------------------------------------------------------------------------------------------------------
def build_node(ctx, workload, dtype, tile=0):
    """The workload's solver node on the device; returns (node, ff, meta)."""
    import sofa_b200 as sb
    w = WORKLOADS[workload]
    template = "B200Vec3f" if dtype == "f32" else "B200Vec3d"
    pos, elems, fixed = build_mesh(workload)
    t0 = time.perf_counter()
    mo = sb.MechanicalObject(ctx, template, position=pos)
    if w["kind"] == "hex":
        ff = sb.HexahedronFEMForceField(mo, elems, youngModulus=w["young"], poissonRatio=w["poisson"], method=w["method"], tileElems=tile)
    elif w["kind"] == "liver":
        ff = sb.TetrahedralCorotationalFEMForceField(mo, elems, youngModulus=w["young"], poissonRatio=w["poisson"], method=w["method"])
    elif w["kind"] == "fast":
        ff = sb.FastTetrahedralCorotationalForceField(mo, elems, youngModulus=w["young"], poissonRatio=w["poisson"], method=w["method"])
    else:
        ff = sb.TetrahedronFEMForceField(mo, elems, youngModulus=w["young"], poissonRatio=w["poisson"], method=w["method"], tileElems=tile)
    mass = sb.DiagonalMass(mo, elems, massDensity=w["density"])
    node = sb.SolverNode(mo, ff, mass, sb.FixedProjectiveConstraint(mo, fixed), dt=w["dt"], gravity=w["gravity"], rayleighStiffness=w["rK"], rayleighMass=w["rM"],
                         iterations=w["iterations"], tolerance=w["tolerance"], threshold=w["threshold"])
    import torch
    torch.cuda.synchronize()
    create_s = time.perf_counter() - t0
    if w["kind"] == "fast":
        return node, ff, dict(pos=pos, E=ff.get("n_edges"), N=pos.shape[0], kind="fast", create_s=create_s, mo=mo, n_tets=elems.shape[0])
    return node, ff, dict(pos=pos, E=elems.shape[0], N=pos.shape[0], kind="hex" if w["kind"] == "hex" else "tet", create_s=create_s, mo=mo)


def time_steps(ctx, node, steps, warmup, local=0, sample_clocks=True):
    """K steps replayed from the step's CUDA graph, CUDA events on the launching stream, sync on both sides."""
    import torch
    stream = torch.cuda.current_stream()
    for _ in range(warmup):
        node.step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local).start() if sample_clocks else None
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        node.step()
    e1.record(stream)
    torch.cuda.synchronize()
    launches = ctx.launch_count - launches0
    clocks = sampler.stop() if sampler else None
    return e0.elapsed_time(e1), launches, clocks


def profile_steps(ctx, node, steps):
    """The same steps with an event pair around every launch (plain launches; the graph is bypassed while the profiler is on)."""
    ctx.profile_begin()
    for _ in range(steps):
        node.step()
    return ctx.profile_end()


def kernel_roofline(prof, ab, iters_per_step, kind, peak, peak_src):
    fused = prof.get("cg_persistent", {}).get("launches", 0) > 0
    ep = prof["cg_persistent"] if fused else prof["element_pass_dforce"]
    k_ms = ep["ms"] / max(ep["launches"], 1)
    k_bytes = ab["cg_iteration"] * iters_per_step if fused else ab["cg_iteration"]
    k_name = (f"fused_cg_kernel<{'Hex' if kind == 'hex' else ('Edge' if kind == 'fast' else 'Tet')}Pass> (the whole CGLinearSolver loop, {iters_per_step} iterations per launch: A*p element pass, "
              "shared-node sums, x/r/p updates, one reduction of four dot products per iteration)") if fused else \
             ("fast_edge_kernel (A*p over the edges of the multi-kernel loop)" if kind == "fast" else f"{'hex' if kind == 'hex' else 'tet'}_tile_kernel (A*p element pass of the multi-kernel loop)")
    achieved = k_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    return dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, kernel=k_name, algorithmic_bytes_per_launch=k_bytes,
                avg_launch_ms=k_ms, launches_timed=ep["launches"], peak_source=peak_src), fused


def run_config(ctx, workload, dtype, steps, warmup, peak, peak_src, local=0):
    """One of the secondary configurations: device-resident steps + the CG kernel's roofline."""
    import torch
    node, ff, meta = build_node(ctx, workload, dtype)
    s = 4 if dtype == "f32" else 8
    ms, launches, clocks = time_steps(ctx, node, steps, warmup, local)
    info = node.last_solve()
    it = min(info["iterations"], WORKLOADS[workload]["iterations"])
    prof = profile_steps(ctx, node, min(steps, 10))
    ab = algorithmic_bytes(meta["kind"], meta["E"], meta["N"], s)
    roof, fused = kernel_roofline(prof, ab, it, meta["kind"], peak, peak_src)
    out = {"workload": workload_string(workload, meta.get("n_tets", meta["E"]), meta["N"], dtype), "dtype": dtype, "value": it * steps / (ms * 1e-3), "unit": "cg_iters/s",
           "ms_per_step": ms / steps, "us_per_step": 1e3 * ms / steps, "steps_per_s": steps / (ms * 1e-3), "cg_iters_per_step": it, "steps": steps, "warmup": warmup,
           "create_s": meta["create_s"], "roofline": roof, "gpu_launches": launches, "layout": ff.stats(), "clocks": clocks,
           "algorithmic_bytes_per_cg_iteration": ab}
    if meta["kind"] == "fast":
        out["edges"] = int(meta["E"])
        out["roofline"]["note"] = "addDForce of this class runs over the topology's edges (one 3x3 matrix per edge): algorithmic bytes = edges x (8 + 9s) + nodes x 34s"
    if meta["kind"] == "hex":
        out["roofline"]["note"] = ("issue-bound, not HBM-bound: 1152 FMUL + 1128 FADD per hexahedron (24x24 K_e row sums without FMA, the reference's rounding) "
                                   "= 2280 fp32 instructions per 68 bytes; the HBM fraction is reported for completeness")
    del node, ff, meta
    torch.cuda.synchronize()
    return out


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
        try:
            return run_ours_distributed(args, rank, world, local)
        finally:
            sys.stdout.flush()
            dist.destroy_process_group()
    wl, dtype = args.workload, args.dtype
    s = 4 if dtype == "f32" else 8
    ndtype = np.float32 if dtype == "f32" else np.float64
    peak, peak_src = measured_peaks()
    ctx = sb.Context(local)
    node, ff, meta = build_node(ctx, wl, dtype, args.tile)
    E, N = meta["E"], meta["N"]
    # ---- device-resident arm
    ms, launches, clocks = time_steps(ctx, node, args.steps, args.warmup, local)
    info = node.last_solve()
    iters_per_step = min(info["iterations"], WORKLOADS[wl]["iterations"])
    # ---- per-kernel durations (roofline.achieved of the dominant kernel)
    prof = profile_steps(ctx, node, args.steps)
    # ---- sustained arm: >= 3 s of back-to-back steps, clocks sampled throughout
    sustained = None
    if args.sustain_s > 0:
        n_sus = max(args.steps, int(args.sustain_s * 1e3 / (ms / args.steps)) + 1)
        ms_sus, _, clocks_sus = time_steps(ctx, node, n_sus, 1, local)
        sustained = {"seconds": ms_sus * 1e-3, "steps": n_sus, "value": iters_per_step * n_sus / (ms_sus * 1e-3), "unit": "cg_iters/s", "ms_per_step": ms_sus / n_sus, "clocks": clocks_sus}
    # ---- end-to-end arms through the host-buffer entry points, pinned host memory, H2D and D2H inside every step:
    # (a) synchronous round trip of the positions (sofab200_node_step_host_x): the host owns x;
    # (b) the coupling of a device-resident state with a host loop (sofab200_node_step_pipelined): this step's external forces up, the new positions
    #     down on a copy stream while the next step is already running (two alternating output buffers, flush inside the timed region).
    xh = torch.from_numpy(meta["pos"].astype(ndtype)).pin_memory()
    vh0 = torch.zeros_like(xh).pin_memory()
    node.step_host_x(xh, vh0)          # (initial velocities uploaded once)
    for _ in range(3):
        node.step_host_x(xh)
    torch.cuda.synchronize()
    e2e_steps = max(1, min(args.steps, 50))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        node.step_host_x(xh)
    torch.cuda.synchronize()
    e2e_sync_s = time.perf_counter() - t0
    ext_h = torch.zeros_like(xh).pin_memory()
    xo = [torch.zeros_like(xh).pin_memory() for _ in range(2)]
    for k in range(4):                 # (the external-force path captures its own step graph)
        node.step_pipelined(ext_h, xo[k & 1])
    node.flush()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        node.step_pipelined(ext_h, xo[k & 1])
    node.flush()
    e2e_s = time.perf_counter() - t0
    node.set_external_force(None)

    value = iters_per_step * args.steps / (ms * 1e-3)
    ab = algorithmic_bytes(meta["kind"], E, N, s)
    roof, fused = kernel_roofline(prof, ab, iters_per_step, meta["kind"], peak, peak_src)
    traffic, traffic_how = measured_traffic(f"{wl}_{dtype}_{'cg_fused' if fused else 'element_pass'}_bytes")
    cg_gbs = ab["cg_iteration"] * iters_per_step * args.steps / (ms * 1e-3) / 1e9
    roof.update(traffic=traffic, traffic_source=traffic_how, fused_kernel=node.fused_info() if hasattr(node, "fused_info") else None,
                cg_loop={"algorithmic_bytes_per_iteration": ab["cg_iteration"], "achieved": cg_gbs, "frac": cg_gbs / peak,
                         "note": "whole step time attributed to the CG iterations (includes addForce, RHS, integration)"},
                kernel_ms={k: v for k, v in prof.items()})
    line = {
        "metric": "cg_iters_per_s", "value": value, "unit": unit_for(wl), "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if wl == "C5" else "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "steps_per_s": args.steps / (ms * 1e-3), "cg_iters_per_step": iters_per_step, "true_cg_iters_per_s": value, "create_s": meta["create_s"],
        "config": {"workload": workload_string(wl, meta.get("n_tets", E), N, dtype), "partition": "single GPU: one partition = the whole mesh, so partition_cg_iters/s == CG iterations/s",
                   "l2": f"working set per CG iteration exceeds the 126 MB L2 ({(E * 120 + N * 34 * s) / 1e6:.0f} MB streamed)" if meta["kind"] == "tet" and E * 120 > 126e6 else
                         "working set per CG iteration fits the 126 MB L2 (inputs are not larger than L2; no flush between steps: the steady state of the simulation loop is what is measured)",
                   "layout": ff.stats()},
        "roofline": roof,
        "e2e": {"value": iters_per_step * e2e_steps / e2e_s, "unit": unit_for(wl), "h2d_bytes_per_step": N * 3 * s, "d2h_bytes_per_step": N * 3 * s,
                "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                "note": "sofab200_node_step_pipelined: device-resident x, v; every step uploads that step's external forces (n Vec3, pinned host memory) and "
                        "downloads the new positions into one of two pinned buffers on a copy stream while the next step runs; the last download is inside the timed region",
                "synchronous_round_trip": {"value": iters_per_step * e2e_steps / e2e_sync_s, "ms_per_step": 1e3 * e2e_sync_s / e2e_steps,
                                           "note": "sofab200_node_step_host_x: the host owns x -- H2D of x, step, D2H of x, host sync, every step (velocities resident)"}},
        "sustained": sustained, "gpu_launches": launches, "clocks": clocks,
    }
    del node, ff
    torch.cuda.synchronize()
    if wl == "C2" and dtype == "f32" and not args.no_configs:
        cfgs = {}
        for name, (w2, d2, st, wu) in {"C3_f32": ("C3", "f32", 20, 3), "C2_f64": ("C2", "f64", 30, 3), "C4_f32": ("C4", "f32", 200, 5), "C2_FAST_f32": ("C2_FAST", "f32", 20, 3)}.items():
            try:
                cfgs[name] = run_config(ctx, w2, d2, st, wu, peak, peak_src, local)
            except Exception as e:  # a secondary configuration must not take the headline line down
                cfgs[name] = {"error": repr(e)}
        line["configs"] = cfgs
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(wl, dtype, args.cpu_steps, 1, 1)
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------------------
# our arm, N GPUs of one node: one process per GPU
# ---------------------------------------------------------------------------------------------------------------------------------
def distributed_parity_check(ctx, world, rank, dtype):
    """Before anything is timed: one small beam (C2_SMALL stretched N times), two steps through the N-GPU path and through the single-GPU
    library path on rank 0, positions compared."""
    import torch
    import sofa_b200 as sb
    import sofa_b200.parallel as PAR
    w = WORKLOADS["C2_SMALL"]
    template = "B200Vec3f" if dtype == "f32" else "B200Vec3d"
    pos, tets, fixed = build_mesh("C2_SMALL", stretch=max(1, world // 2))
    node = PAR.DistributedSolverNode(pos, tets, fixed, w["density"], w["young"], w["poisson"], w["method"], ctx=ctx, template=template, dt=w["dt"], gravity=w["gravity"],
                                     rayleighStiffness=w["rK"], rayleighMass=w["rM"], iterations=w["iterations"], tolerance=w["tolerance"], threshold=w["threshold"])
    for _ in range(2):
        node.step()
    x_dist = node.gather_global(node.be.x, pos.shape[0])
    end = node.be.node.last_solve()["end_condition"]
    out = None
    if rank == 0:
        mo = sb.MechanicalObject(ctx, template, position=pos)
        ff = sb.TetrahedronFEMForceField(mo, tets, youngModulus=w["young"], poissonRatio=w["poisson"], method=w["method"])
        single = sb.SolverNode(mo, ff, sb.DiagonalMass(mo, tets, massDensity=w["density"]), sb.FixedProjectiveConstraint(mo, fixed), dt=w["dt"], gravity=w["gravity"],
                               rayleighStiffness=w["rK"], rayleighMass=w["rM"], iterations=w["iterations"], tolerance=w["tolerance"], threshold=w["threshold"])
        for _ in range(2):
            single.step()
        x1 = mo.x.detach().cpu().numpy().astype(np.float64)
        err = float(np.abs(x_dist - x1).max())
        moved = float(np.abs(x1 - pos).max())
        out = {"mesh": f"C2_SMALL x {max(1, world // 2)}: {tets.shape[0]} tetrahedra over {world} GPUs vs the single-GPU library path, 2 steps", "max_abs_dx": err,
               "displacement": moved, "tolerance": 1e-4 if dtype == "f32" else 1e-9, "end_condition": end, "peer_memory": bool(getattr(node.be, "peer", False))}
        if not (err <= out["tolerance"]) or end == 99:
            raise RuntimeError(f"multi-GPU parity check failed: {out}")
    del node
    torch.cuda.synchronize()
    return out


def run_ours_distributed(args, rank, world, local):
    """N > 1.  Default workload: ONE beam N times as long as C2, cut into N slabs of C2's size (weak scaling; `value` counts one CG iteration of
    the long beam as N partition-iterations).  --workload C5: the 15.7 M-tet beam itself cut into N slabs (strong scaling; value = CG iterations/s).
    Interface partial sums and the dot products go through NVLink peer memory inside the persistent CG kernel (sofa_b200/parallel.py)."""
    import torch
    import torch.distributed as dist
    import sofa_b200 as sb
    import sofa_b200.parallel as PAR
    wl, dtype = args.workload, args.dtype
    w = WORKLOADS[wl]
    s = 4 if dtype == "f32" else 8
    template = "B200Vec3f" if dtype == "f32" else "B200Vec3d"
    strong = wl == "C5"
    ctx = sb.Context(local)
    parity = distributed_parity_check(ctx, world, rank, dtype)
    pos, tets, fixed = build_mesh(wl, stretch=1 if strong else world)
    t0c = time.perf_counter()
    node = PAR.DistributedSolverNode(pos, tets, fixed, w["density"], w["young"], w["poisson"], w["method"], ctx=ctx, template=template, partition=args.partition,
                                     dt=w["dt"], gravity=w["gravity"], rayleighStiffness=w["rK"], rayleighMass=w["rM"],
                                     iterations=w["iterations"], tolerance=w["tolerance"], threshold=w["threshold"])
    torch.cuda.synchronize()
    create_s = time.perf_counter() - t0c
    stream = torch.cuda.current_stream()

    def barrier():
        dist.barrier(); torch.cuda.synchronize()
    for _ in range(args.warmup):
        node.step()
    barrier()
    sampler = ClockSampler(local).start()
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        node.step()
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    # ---- per-kernel durations on every rank (plain launches with an event pair around each; all ranks must run the same steps)
    ctx.profile_begin()
    for _ in range(min(args.steps, 20)):
        node.step()
    prof = ctx.profile_end()
    barrier()
    # ---- end-to-end arm: every rank couples its partition's device-resident state to a host loop (sofab200_node_step_pipelined): that step's external
    # forces of the rank's nodes up from pinned host memory, the new positions down into one of two pinned buffers under the next step
    ext_h = torch.zeros_like(node.be.x).cpu().pin_memory()
    xo = [torch.zeros_like(node.be.x).cpu().pin_memory() for _ in range(2)]
    for k in range(4):
        node.be.node.step_pipelined(ext_h, xo[k & 1], node.be.x, node.be.v)
    node.be.node.flush()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 50))
    for k in range(e2e_steps):
        node.be.node.step_pipelined(ext_h, xo[k & 1], node.be.x, node.be.v)
    node.be.node.flush()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e0.elapsed_time(e1), e2e_s], dtype=torch.float64, device=ctx.device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_s = float(t[0]), float(t[1])
    sizes = torch.tensor([node.rm.elems.shape[0], node.rm.n_local, len(node.rm.interface)], dtype=torch.int64, device=ctx.device)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    if rank != 0:
        return
    T_loc, N_loc = node.rm.elems.shape[0], node.rm.n_local
    info = node.be.node.last_solve()
    if info["end_condition"] == 99:
        raise RuntimeError("a cross-GPU wait timed out inside the CG kernel (end_condition 99): the numbers of this run are void")
    it_step = min(info["iterations"], w["iterations"])
    iters = it_step * args.steps     # every step of this workload runs the same forced iteration count
    peer = bool(getattr(node.be, "peer", False))
    exchange = ("ONE persistent CG kernel per GPU (cg_fused.cuh); interface partial sums stored into the neighbours' mailboxes over NVLink (CUDA IPC peer memory), "
                "one rank-ordered all-reduce of four dot products per iteration inside the kernel, no NCCL call in the loop") if peer else \
               "multi-kernel loop, NCCL send/recv + allreduce enqueued by the library, device-resident CG scalars"
    peak, peak_src = measured_peaks()
    true_rate = iters / (ms * 1e-3)
    units = 1 if strong else world       # partitions of the N=1 workload's size that one CG iteration processes
    value = true_rate * units
    max_T = max(int(a[0]) for a in all_sizes); max_N = max(int(a[1]) for a in all_sizes)
    ab = algorithmic_bytes("tet", max_T, max_N, s)
    cg_gbs = ab["cg_iteration"] * iters / (ms * 1e-3) / 1e9
    line = {"metric": "cg_iters_per_s", "value": value, "unit": "cg_iters/s" if strong else "partition_cg_iters/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "steps_per_s": args.steps / (ms * 1e-3), "cg_iters_per_step": it_step, "true_cg_iters_per_s": true_rate, "create_s": create_s,
            "config": {"workload": workload_string(wl, tets.shape[0] if strong else tets.shape[0] // world, pos.shape[0] if strong else N_loc, dtype),
                       "mesh": f"ONE beam {'(the workload itself)' if strong else f'{world} times as long as the workload'}: {tets.shape[0]} tetrahedra, {pos.shape[0]} nodes, "
                               f"{args.partition} partition into {world} parts",
                       "partition": f"per GPU (tets, nodes, interface nodes): {[[int(v) for v in a] for a in all_sizes]}; {exchange}",
                       "value_counts": ("strong scaling: value = CG iterations per second of the whole mesh" if strong else
                                        f"weak scaling: one CG iteration of the {world}x longer beam processes {world} partitions of the N=1 workload's size and is counted as {world} "
                                        "partition-iterations (unit partition_cg_iters/s, the same unit at N=1 where a partition is the whole mesh); true_cg_iters_per_s is the plain rate"),
                       "l2": "working set per CG iteration and GPU exceeds L2" if max_T * 120 > 126e6 else "per-GPU working set fits the 126 MB L2"},
            "parity_check": parity,
            "roofline": {"bound": "hbm", "achieved": cg_gbs, "peak": peak, "unit": "GB/s", "frac": cg_gbs / peak, "traffic": None,
                         "kernel": "whole distributed step per GPU attributed to the CG iterations (algorithmic bytes of the largest partition)", "peak_source": peak_src,
                         "kernel_ms_rank0": {k: v for k, v in prof.items()}},
            "e2e": {"value": it_step * e2e_steps * units / e2e_s, "unit": "cg_iters/s" if strong else "partition_cg_iters/s", "h2d_bytes_per_step": sum(int(a[1]) for a in all_sizes) * 3 * s,
                    "d2h_bytes_per_step": sum(int(a[1]) for a in all_sizes) * 3 * s, "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                    "note": "sofab200_node_step_pipelined on every rank: the partition's state is device-resident, every step uploads that step's external forces of the rank's nodes (pinned host memory) and downloads the new positions into one of two pinned buffers under the next step; wall clock, max over ranks, last download inside the timed region"},
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
    ap.add_argument("--partition", default="slab", choices=["slab", "rcb"], help="N > 1: contiguous element ranges (z-slabs of a grid beam) or recursive coordinate bisection")
    ap.add_argument("--cpu-steps", type=int, default=6, help="oracle steps timed for cpu_baseline")
    ap.add_argument("--sustain-s", type=float, default=3.0, help="seconds of the sustained arm (0 = off)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cpu-variants", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the secondary configurations (C3, C2 in Vec3d, C4)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_ours(args)


if __name__ == "__main__":
    main()
