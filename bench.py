#!/usr/bin/env python
"""bench.py — headline benchmark of the KD-tree path-tracing hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload bunny|dragon5|sphere]

Workload (BASELINE.json configs[1]): scenes/cornell8.txt + stanford_bunny.obj, 800x800, KD full-stack traversal.
One STEP = one batch of `--spp-per-step` samples (iterations) of the whole render loop over the 640 000-pixel frame:
generate rays -> [intersect(+scatter) -> shade -> gather -> compact (-> sort)] x bounces -> accumulate.
metric = Mrays/s = ray-bounces traversed+intersected per second, whole job (all ranks). spp/s is reported next to it.

N > 1: one process per GPU (torchrun), samples sharded by iteration index (rank r renders iterations r+1, r+1+N, ...;
the tree is replicated), and every step ends with one NCCL all-reduce of the 800x800x3 float radiance buffer.
Scaling is weak: every rank renders `spp-per-step` samples per step.

The product path is libkdpt.so (hand-written sm_100a kernels behind the C ABI of include/kdpt.h). oracle/ is only
used for the reported cpu_baseline, the algorithmic-bytes bookkeeping of the roofline, and `--impl reference`.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: (fixture, W, H, kdpt mode name, extra render params, description)
    "bunny": ("cornell8_stanford_bunny.npz", 800, 800, "fullstack", {},
              "scenes/cornell8.txt + stanford_bunny.obj (309 tris, 265 nodes), 800x800, DEPTH 4, KD full-stack traversal"),
    "dragon5": ("cornell_dragon_5.npz", 800, 800, "shortstack", dict(dofAngle=0.01, focalLength=6.0, sortByMaterial=2),
                "scenes/cornell.txt + dragon_5.obj (12 500 tris, 2 645 nodes), 800x800, DEPTH 8, short-stack, AA+DOF, compaction + material sort"),
    "sphere": ("cornell_sphere_low_1.npz", 800, 800, "fullstack", {},
               "scenes/cornell.txt + sphere_low_1.obj (20 tris, 63 nodes), 800x800, DEPTH 8"),
    # BASELINE configs[3]: synthetic ~1.6 M-triangle dragon (scenes/dragon.obj subdivided twice) in the cornell box, 1080p
    "dragon1m": ("cornell_sphere_low_1.npz", 1920, 1080, "fullstack", {},
                 "scenes/cornell.txt walls + scenes/dragon.obj subdivided 2x (1.6 M tris, scale 0.6), 1920x1080, DEPTH 8, full-stack"),
}
# traversal-only microbenchmarks (kdpt_trace): name -> (triangles, rays per GPU, description)   BASELINE configs[4]
TRACE_WORKLOADS = {
    "soup10m": (10_000_000, 1 << 24, "synthetic 10 M-triangle random soup (seed 12345), incoherent rays (origins uniform in the root box, "
                "directions uniform on the sphere), traversal+intersection only, full-stack"),
    "soup1m": (1_000_000, 1 << 24, "synthetic 1 M-triangle random soup (seed 12345), incoherent in-box rays, traversal only, full-stack"),
}
HBM_FALLBACK_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md, used only if MEASURED_PEAKS.json is absent


def measured_traffic(workload):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/r1_traffic.json)"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))[workload]
        return t["dram_bytes_per_launch"], t["kernel"]
    except Exception:
        return None, None


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy kernel)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def l2_reference(achieved):
    """L2-resident read bandwidth measured on this pool's B200 with tools/l2_peak.cu (profiles/r1_l2_peak.json): the
    physical ceiling of the workloads whose tree fits the 126 MB L2; informational, next to the contract's
    HBM denominator"""
    try:
        z = json.load(open(os.path.join(ROOT, "profiles", "r1_l2_peak.json")))["read_gbs"]
        pk = max(v for k, v in z.items() if int(k.split()[0]) <= 64)
        return {"l2_read_peak_gbs": pk, "frac_of_l2": (achieved / pk) if achieved else None, "hbm_read_peak_gbs": z.get("4096 MB"),
                "l2_source": "profiles/r1_l2_peak.json (ld.global.cg read kernel, best of 8-64 MB buffers)"}
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (recipe's clocks line)."""

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def load_scene(workload: str):
    import kdtreepathtraceroptimization_b200 as kdpt
    fixture, w, h, mode, extra, desc = WORKLOADS[workload]
    sc = kdpt.Scene.from_npz(os.path.join(ROOT, "tests", "golden", fixture)).with_resolution(w, h).first_frame_camera()
    if workload == "dragon1m":
        from kdtreepathtraceroptimization_b200 import meshes
        t18 = meshes.dense_dragon(os.path.join(ROOT, "tests", "golden", "dragon_100k.npz"), 2, 0.6)
        tree = kdpt.Scene.from_triangles(t18)                     # the library's host-side KD builder (reference tree)
        sc.nodes, sc.tris = tree.nodes, tree.tris
        # dragon.mtl: illum 6 -> glass (REFL 1, REFR 1, ior 1.5), scene.cpp:748-760
        m = sc.materials[-1:].copy()
        m['color'] = 1.0; m['specular_exponent'] = 1.0; m['specular_color'] = 1.0; m['hasReflective'] = 1.0
        m['hasRefractive'] = 1.0; m['indexOfRefraction'] = 1.5; m['transmittance'] = 0.0; m['emittance'] = 0.0
        sc.materials = np.concatenate([sc.materials[:-1], m])
    params = kdpt.default_params(mode=kdpt.MODES[mode], **extra)
    return sc, params, desc


# ---- CPU side: the reference's own host functions (oracle/_ref) or, if that file is absent, the oracle port ----------
def cpu_iteration_runner(sc, params):
    """returns (kind, fn(iter) -> (ray_bounces, seconds), cores)"""
    from conftest import KO_PATH, P, ko_params, ko_scene
    cores = os.cpu_count() or 1
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
    O = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
    O.ko_render_iteration.restype = ctypes.c_long
    ks, kp = ko_scene(sc), ko_params(params, sc.trace_depth)
    n = int(sc.camera['resolution'][0][0]) * int(sc.camera['resolution'][0][1])
    so = os.path.join(ROOT, "oracle", "_ref", "libkdref.so")
    R = None
    if os.path.exists(so):
        try:
            R = ctypes.CDLL(so)
            R.kdref_cpu_iteration.restype = ctypes.c_long
        except OSError:
            R = None
    img = np.zeros((n, 3), 'f4')
    # the reference has only its two lossy traversals; fullstack/ref_t1 -> traverseKDbare, shortstack/ref_t2 -> ShortHybrid
    shortstack = 1 if params.mode in (1, 3) else 0
    if R is not None:
        def run(it):
            paths = np.zeros(n, KO_PATH)
            O.ko_generate_rays(ctypes.byref(ks.cam), it, sc.trace_depth, ctypes.c_float(params.focalLength),
                               ctypes.c_float(params.dofAngle), params.antialias, P(paths))
            secs = np.zeros(2, 'f8')
            t0 = time.perf_counter()
            tot = R.kdref_cpu_iteration(P(sc.nodes), len(sc.nodes), P(sc.tris), P(sc.geoms), len(sc.geoms), P(sc.materials),
                                        len(sc.materials), P(sc.offsets), P(paths), n, it, ctypes.c_float(params.softness),
                                        params.enableSss, params.compaction, shortstack, params.sortByMaterial, P(img), cores, P(secs))
            return int(tot), time.perf_counter() - t0
        return "reference", run, cores

    def run(it):
        t0 = time.perf_counter()
        tot = O.ko_render_iteration(ctypes.byref(ks), ctypes.byref(kp), it, P(img), None, cores)
        return int(tot), time.perf_counter() - t0
    return "port", run, cores


def algorithmic_bytes_from_oracle(sc, params):
    """bytes(ray) = 32 (ray in) + 16 (hit out) + 16 V + 4 L + 48 T, plus per path-bounce 2x56 B path
    state and 24 B of accumulation; V/L/T from the instrumented CPU traversal over the rays of iteration 1."""
    from conftest import KO_CTR, P, ko_params, ko_scene
    import kdtreepathtraceroptimization_b200 as kdpt
    O = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
    O.ko_render_iteration_counted.restype = ctypes.c_long
    p2 = kdpt.default_params()
    ctypes.memmove(ctypes.byref(p2), ctypes.byref(params), ctypes.sizeof(p2))
    if p2.mode >= 2:      # counters are defined on the exact near-first traversal
        p2.mode -= 2
    ks, kp = ko_scene(sc), ko_params(p2, sc.trace_depth)
    n = int(sc.camera['resolution'][0][0]) * int(sc.camera['resolution'][0][1])
    img = np.zeros((n, 3), 'f4'); ctr = np.zeros(1, KO_CTR)
    tot = O.ko_render_iteration_counted(ctypes.byref(ks), ctypes.byref(kp), 1, P(img), None, os.cpu_count() or 1, P(ctr))
    V = float(ctr['nodes'][0]) / tot; T = float(ctr['tris'][0]) / tot
    traversal = 32 + 16 + 16 * V + 4 * T + 48 * T
    return dict(V=V, T=T, traversal_bytes=traversal, path_bytes=2 * 56 + 24, bytes_per_ray_bounce=traversal + 2 * 56 + 24,
                ray_bounces_iter1=int(tot))


def algorithmic_bytes_per_ray(workload):
    """the committed per-ray figures (profiles/algorithmic_bytes.json, regenerated by tools/algorithmic_bytes.py)"""
    return json.load(open(os.path.join(ROOT, "profiles", "algorithmic_bytes.json")))[workload]


def run_trace_workload(args):
    """BASELINE configs[4]: traversal-only microbenchmark, rays sharded in contiguous chunks, no collective."""
    import torch
    import torch.distributed as dist
    import kdtreepathtraceroptimization_b200 as kdpt
    from kdtreepathtraceroptimization_b200 import meshes
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    ntri, nrays, desc = TRACE_WORKLOADS[args.workload]
    if args.impl == "reference":
        if rank != 0:
            return
        nrays = 20000
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — the product has no CPU path")
        torch.cuda.set_device(local)
        if world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sc = kdpt.Scene.from_triangles(meshes.triangle_soup(ntri))
    lo, hi = sc.nodes['mins'][0].astype('f8'), sc.nodes['maxs'][0].astype('f8')
    rs = np.random.default_rng(777 + rank)
    d = rs.normal(size=(nrays, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([lo + rs.random((nrays, 3)) * (hi - lo), d], 1).astype('f4')
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import KO_CTR, KO_HIT, KO_STACK_T1, KO_T1, P
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
    O = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
    cores = os.cpu_count() or 1

    def cpu_run(mode, k):
        out = np.zeros(k, KO_HIT); ctr = np.zeros(1, KO_CTR)
        t0 = time.perf_counter()
        O.ko_trace_batch(mode, P(sc.nodes), len(sc.nodes), P(sc.tris), len(sc.tris), 1, P(rays[:k]), k, None, P(out), P(ctr), cores)
        return time.perf_counter() - t0, ctr[0]
    W = max(args.warmup, 3); K = args.steps
    if args.impl == "reference":
        # the reference's traverseKDbare semantics (its own function overflows nodeIDs[4000] on this 32 767-node tree:
        # SURVEY 8c) through the guarded restatement in the oracle
        # each step is a bounded sample (20 000 rays) of the batch: the whole 16.7 M-ray batch would take minutes per step on the host
        secs = 0.0
        for k in range(K):
            dt, _ = cpu_run(KO_T1, nrays); secs += dt
        v = K * nrays / secs / 1e6
        print(json.dumps({"impl": "reference", "metric": "Mrays/s (traverse+intersect)", "value": v, "unit": "Mrays/s", "n_gpus": args.gpus,
                          "steps": K, "warmup": 0, "ms_per_step": 1e3 * secs / K, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": desc, "rays_per_step": nrays},
                          "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port",
                                           "sample": f"{nrays} rays of the workload's distribution per step, {K} steps"},
                          "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return
    ctx = kdpt.Context(sc, device=local)
    ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
    d_rays = torch.from_numpy(rays).cuda(); d_hits = torch.zeros(nrays, 8, dtype=torch.float32, device="cuda")
    for _ in range(W):
        ctx.trace_device(d_rays.data_ptr(), None, nrays, d_hits.data_ptr(), kdpt.FULLSTACK, ts.cuda_stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = ClockSampler(local)
    time.sleep(0.3)
    clocks.rows.clear()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    launches0 = ctx.counters()[1]
    for k in range(K):           # 0.9 GB of tree + 0.9 GB of rays/hits per step >> 126 MB L2: no flush needed
        evs[k][0].record(ts)
        ctx.trace_device(d_rays.data_ptr(), None, nrays, d_hits.data_ptr(), kdpt.FULLSTACK, ts.cuda_stream)
        evs[k][1].record(ts)
    torch.cuda.synchronize()
    clk = clocks.stop()
    ms = sum(a.elapsed_time(b) for a, b in evs)
    launches = ctx.counters()[1] - launches0      # per step: bin keys, bin scan, bin scatter, traversal
    hrays = torch.from_numpy(rays).pin_memory().numpy()
    hhits = torch.empty(nrays * kdpt.HIT.itemsize, dtype=torch.uint8).pin_memory().numpy().view(kdpt.HIT)
    t0 = time.perf_counter()
    for k in range(3):
        kdpt._check(ctx._lib.kdpt_trace(ctx._h, hrays.ctypes.data, None, nrays, hhits.ctypes.data, kdpt.FULLSTACK), "kdpt_trace")
    e2e_s = (time.perf_counter() - t0) / 3
    t = torch.tensor([ms, e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_s = float(t[0]), float(t[1])
    if rank == 0:
        value = world * nrays * K / (ms * 1e-3) / 1e6
        ab = algorithmic_bytes_per_ray(args.workload)
        V, T, bpr = ab["V"], ab["T"], ab["bytes_per_ray_bounce"]
        peak, peak_src = peak_hbm()
        achieved = bpr * nrays / (ms / K * 1e-3) / 1e9
        # CPU baseline: a first 20 000-ray sample sizes a second one of ~10 s (bounded by the batch)
        dt0, _ = cpu_run(KO_T1, 20000)
        kcpu = int(max(20000, min(nrays, 20000 * 10.0 / max(dt0, 1e-3))))
        dt1, _ = cpu_run(KO_T1, kcpu)
        print(json.dumps({"metric": "Mrays/s (traverse+intersect)", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": K, "warmup": W,
                          "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic", "config": {"workload": desc, "rays_per_step_per_gpu": nrays, "tree": ctx.tree_info(), "tree_refined": ctx.refined_info(),
                                                          "l2": "working set (tree 0.9 GB + rays/hits 0.9 GB) exceeds L2; no flush"},
                          "clocks": clk, "gpu_launches": int(launches),
                          "e2e": {"value": world * nrays / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": nrays * 24, "d2h_bytes_per_step": nrays * 32},
                          "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                                       "traffic": measured_traffic(args.workload)[0],
                                       "kernel": "traceKernel<fullstack>", "avg_launch_ms": ms / K, "algorithmic_bytes_per_ray": bpr,
                                       "V_nodes_per_ray": V, "T_tris_per_ray": T, "peak_source": peak_src,
                                       "traffic_gbs_at_this_speed": (measured_traffic(args.workload)[0] or 0) / (ms / K * 1e-3) / 1e9,
                                       "note": "achieved counts the ALGORITHMIC bytes of the unculled near-first traversal (SURVEY 8d); chunk-box culling "
                                               "and leaf-ordered ray binning make the kernel touch far fewer (traffic = ncu DRAM bytes per launch), so "
                                               "frac > 1 and the kernel is latency/issue-bound, not HBM-bound (profiles/r1_trace_soup10m.txt)"},
                          "cpu_baseline": {"value": kcpu / dt1 / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
                                           "sample": "the first %d of the same rays through the reference's traverseKDbare semantics (oracle restatement, "
                                                     "the reference's own function overflows nodeIDs[4000] on this tree), %.1f s" % (kcpu, dt1)}}), flush=True)
    ctx.close()
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sc, params, desc = load_scene(args.workload)
    kind, run, cores = cpu_iteration_runner(sc, params)
    for w in range(min(args.warmup, 1)):
        run(1)
    tot = 0; secs = 0.0
    for k in range(args.steps):
        t, s = run(1 + k)
        tot += t; secs += s
    v = tot / secs / 1e6
    w, h = (int(x) for x in sc.camera['resolution'][0])
    line = {"impl": "reference", "metric": "Mrays/s (traverse+intersect, full render loop)", "value": v, "unit": "Mrays/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * secs / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "spp_per_step": 1, "resolution": [w, h], "trace_depth": sc.trace_depth,
                       "note": "the reference's own host functions (traverseKDbare/ShortHybrid, geoms, scatterRay) driven by "
                               "oracle/ref_shim.cu on the host cores; each step is ONE sample of the frame"
                       if kind == "reference" else "oracle port (oracle/_ref/libkdref.so absent); each step is ONE sample"},
            "spp_per_s": args.steps / secs,
            "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": kind,
                             "sample": f"{args.steps} iteration(s) of the 800x800 frame, {tot} ray-bounces"},
            "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="kdpt", choices=["kdpt", "reference"])
    ap.add_argument("--workload", default="bunny", choices=sorted(WORKLOADS) + sorted(TRACE_WORKLOADS))
    ap.add_argument("--spp-per-step", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default=None, choices=["fullstack", "shortstack", "ref_t1", "ref_t2"],
                    help="override the workload's traversal mode")
    args = ap.parse_args()
    if args.workload in TRACE_WORKLOADS:
        run_trace_workload(args)
        return
    if args.impl == "reference":
        if args.steps > 8:
            args.steps = 8          # one step = one full 800x800 sample on the CPU; keep the arm within minutes
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    import kdtreepathtraceroptimization_b200 as kdpt

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W = max(args.warmup, 3); K = args.steps; S = args.spp_per_step
    sc, params, desc = load_scene(args.workload)
    mode_name = args.mode or WORKLOADS[args.workload][3]
    params.mode = kdpt.MODES[mode_name]
    w, h = (int(x) for x in sc.camera['resolution'][0])
    npix = w * h

    ctx = kdpt.Context(sc, device=local)
    ts = torch.cuda.Stream()
    ctx.set_stream(ts.cuda_stream)
    img = torch.zeros(3 * npix, dtype=torch.float32, device="cuda")
    total_img = torch.zeros_like(img)
    ctx.bind_image(img.data_ptr())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")        # > 126 MB L2

    def step(k: int):
        # sample sharding: rank r takes iterations r+1, r+1+world, ... ; step k covers the next S of them
        ctx.render_async(1 + k * S * world + rank, S, world, params)
        if world > 1:
            with torch.cuda.stream(ts):
                total_img.copy_(img, non_blocking=True)
                dist.all_reduce(total_img)

    clocks = ClockSampler(local)            # nvidia-smi needs ~100 ms to start: launch it before the warm-up
    with torch.cuda.stream(ts):
        for k in range(W):
            flush.fill_(k & 0xff)
            step(k)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    rb0, l0 = ctx.counters()
    clocks.rows.clear()                     # keep only the samples taken during the timed region
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    with torch.cuda.stream(ts):
        for k in range(K):
            flush.fill_(k & 0xff)                                           # L2 flush between timed steps (untimed)
            evs[k][0].record(ts)
            step(W + k)
            evs[k][1].record(ts)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    rb1, l1 = ctx.counters()
    ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([ms, float(rb1 - rb0)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, rays = float(tmax[0]), float(tsum[1])
    else:
        rays = float(t[1])
    value = rays / (ms * 1e-3) / 1e6
    launches = (l1 - l0) + (2 * K if world > 1 else 0) + K      # + copy/all-reduce per step (library) + flush fill

    # ---- end to end through the C ABI with HOST buffers: pathtraceInit (flatten + H2D) + pathtrace x S + image D2H ----
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).pin_memory().numpy().view(a.dtype).reshape(a.shape)
    hsc = kdpt.Scene(nodes=pin(sc.nodes), tris=pin(sc.tris), geoms=pin(sc.geoms), materials=pin(sc.materials),
                     offsets=pin(sc.offsets), camera=sc.camera, trace_depth=sc.trace_depth)
    himg = torch.empty((h, w, 3), dtype=torch.float32).pin_memory()
    lib = kdpt.load_library()
    Ke = max(3, min(K, 10))
    e2e_s = 0.0; e2e_rays = 0; h2d = 0
    for k in range(Ke + 1):
        flush.fill_(k & 0xff); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        c2 = kdpt.Context(hsc, device=local)
        t1 = time.perf_counter()
        st = c2.render(1 + k * S * world + rank, S, world, params)
        t2 = time.perf_counter()
        kdpt._check(lib.kdpt_get_image(c2._h, ctypes.c_void_p(himg.data_ptr())), "kdpt_get_image")
        if rank == 0 and os.environ.get("KDPT_BENCH_DEBUG"):
            print(f"e2e step {k}: create {1e3 * (t1 - t0):.2f} ms, render {1e3 * (t2 - t1):.2f} ms, get_image {1e3 * (time.perf_counter() - t2):.2f} ms", file=sys.stderr)
        if world > 1:
            tmp = himg.reshape(-1).cuda(non_blocking=True)
            dist.all_reduce(tmp); torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        h2d = c2.tree_info()["device_bytes"] + sc.geoms.nbytes + sc.materials.nbytes + sc.offsets.nbytes
        c2.close()
        if k > 0:
            e2e_s += dt; e2e_rays += st["ray_bounces"]
    clk = clocks.stop()         # sampled over the device-timed steps and the end-to-end steps
    et = torch.tensor([e2e_s, float(e2e_rays)], dtype=torch.float64, device="cuda")
    if world > 1:
        emax = et.clone(); dist.all_reduce(emax, op=dist.ReduceOp.MAX)
        esum = et.clone(); dist.all_reduce(esum, op=dist.ReduceOp.SUM)
        e2e_s, e2e_rays = float(emax[0]), float(esum[1])
    e2e_value = e2e_rays / e2e_s / 1e6

    out = None
    if rank == 0:
        # ---- roofline of the dominant kernel (bounceKernel): live CUDA-event durations via TESTINGMODE ----
        peak, peak_src = peak_hbm()
        roof = None; cpu = None
        try:
            ab = algorithmic_bytes_per_ray(args.workload)
            pt = kdpt.default_params(); ctypes.memmove(ctypes.byref(pt), ctypes.byref(params), ctypes.sizeof(pt)); pt.testingmode = 1
            ctx.render(1, S, 1, pt)
            st = ctx.render(1 + S, S, 1, pt)
            nb = max(1, st["intersect_launches"])
            launch_ms = st["intersect_ms"] / nb
            bytes_per_launch = ab["bytes_per_ray_bounce"] * st["ray_bounces"] / nb
            achieved = bytes_per_launch / (launch_ms * 1e-3) / 1e9
            traffic, traffic_kernel = measured_traffic(args.workload)
            roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "traffic_of": traffic_kernel,
                    "kernel": "bounceKernel<%s>" % mode_name, "avg_launch_ms": launch_ms,
                    "kernel_share_of_step": st["intersect_ms"] / st["total_ms"],
                    "algorithmic_bytes_per_ray_bounce": ab["bytes_per_ray_bounce"], "V_nodes_per_ray": ab["V"], "T_tris_per_ray": ab["T"],
                    "peak_source": peak_src, **l2_reference(achieved),
                    "note": "tree + triangles of this workload are %d KB, i.e. L1/L2-resident: the HBM figure is the "
                            "contract's denominator, not the physical bound (see DESIGN.md)" % (ctx.tree_info()["device_bytes"] // 1024)}
        except Exception as e:      # never let the roofline bookkeeping kill the bench line
            roof = {"bound": "hbm", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None, "error": repr(e)}
        if world == 1 and not args.no_cpu_baseline:
            try:
                kind, run, cores = cpu_iteration_runner(sc, params)
                tot = 0; secs = 0.0; it = 0
                while secs < 10.0 and it < 400:         # ~10 s of the host arm
                    a, b = run(1 + it); tot += a; secs += b; it += 1
                cpu = {"value": tot / secs / 1e6, "unit": "Mrays/s", "cores": cores, "kind": kind,
                       "sample": f"{it} full {w}x{h} iteration(s) of the same workload ({tot} ray-bounces, {secs:.1f} s)"}
            except Exception as e:
                cpu = {"value": None, "unit": "Mrays/s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: " + repr(e)}
        out = {"metric": "Mrays/s (traverse+intersect, full render loop)", "value": value, "unit": "Mrays/s", "n_gpus": world,
               "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f32", "data": "synthetic",
               "config": {"workload": desc, "spp_per_step": S, "resolution": [w, h], "trace_depth": sc.trace_depth,
                          "traversal": mode_name, "tree": ctx.tree_info(), "tree_refined": ctx.refined_info(), "parallelism": f"sample-sharded x{world}, tree replicated, "
                          "NCCL all-reduce of the radiance buffer per step" if world > 1 else "single GPU",
                          "l2": "flushed between timed steps (256 MiB fill, untimed)"},
               "spp_per_s": S * world * K / (ms * 1e-3),
               "clocks": clk, "gpu_launches": int(launches),
               "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(npix * 12),
                       "steps": Ke, "timing": "host wall clock around kdpt_create + kdpt_render + kdpt_get_image (synchronous C-ABI calls, pinned host buffers)"},
               "roofline": roof}
        if cpu is not None:
            out["cpu_baseline"] = cpu
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
