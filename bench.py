#!/usr/bin/env python
"""bench.py — headline benchmark of the KD-tree path-tracing hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME] [--mode MODE]

Default workload = BASELINE.json configs[3], the north-star configuration: the cornell box (scenes/cornell.txt) with a
~1.6 M-triangle dense dragon (scenes/dragon.obj subdivided twice; the reference's own large dragon is missing upstream),
1920x1080, KD full-stack traversal, sample-sharded over the GPUs with an NCCL all-reduce of the radiance buffer.
One STEP = one batch of `--spp-per-step` samples (iterations) of the whole render loop over the frame:
generate rays -> [intersect(+scatter) -> shade -> gather -> compact (-> sort)] x bounces -> accumulate.
metric = Mrays/s = ray-bounces traversed+intersected per second, whole job (all ranks). spp/s is reported next to it.
Other workloads (`--workload`): bunny (configs[1]), dragon5 (configs[2]), sphere (configs[0] on the GPU), and the
traversal-only microbenchmarks soup10m (configs[4]) / soup1m.

N > 1: one process per GPU (torchrun), samples sharded by iteration index (rank r renders iterations r+1, r+1+N, ...;
the tree is replicated), and every step ends with one NCCL all-reduce of the W x H x 3 float radiance buffer.
Scaling is weak: every rank renders `spp-per-step` samples per step.

The product path is libkdpt.so (hand-written sm_100a kernels behind the C ABI of include/kdpt.h). oracle/ is only
used for the reported cpu_baseline, the sampled parity check of what was rendered, and `--impl reference` (which never
loads libkdpt.so).
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
    # BASELINE configs[3]: synthetic ~1.6 M-triangle dragon (scenes/dragon.obj subdivided twice) in the cornell box, 1080p
    "dragon1m": ("cornell_sphere_low_1.npz", 1920, 1080, "fullstack", {},
                 "scenes/cornell.txt walls + scenes/dragon.obj subdivided 2x (1.6 M tris, scale 0.6), 1920x1080, DEPTH 8, full-stack"),
    "bunny": ("cornell8_stanford_bunny.npz", 800, 800, "fullstack", {},
              "scenes/cornell8.txt + stanford_bunny.obj (309 tris, 265 nodes), 800x800, DEPTH 4, KD full-stack traversal"),
    "dragon5": ("cornell_dragon_5.npz", 800, 800, "shortstack", dict(dofAngle=0.01, focalLength=6.0, sortByMaterial=2),
                "scenes/cornell.txt + dragon_5.obj (12 500 tris, 2 645 nodes), 800x800, DEPTH 8, short-stack, AA+DOF, compaction + material sort"),
    "sphere": ("cornell_sphere_low_1.npz", 800, 800, "fullstack", {},
               "scenes/cornell.txt + sphere_low_1.obj (20 tris, 63 nodes), 800x800, DEPTH 8"),
}
# traversal-only microbenchmarks (kdpt_trace): name -> (triangles, rays per GPU, description)   BASELINE configs[4]
TRACE_WORKLOADS = {
    "soup10m": (10_000_000, 1 << 24, "synthetic 10 M-triangle random soup (seed 12345), incoherent rays (origins uniform in the root box, "
                "directions uniform on the sphere), traversal+intersection only, full-stack"),
    "soup1m": (1_000_000, 1 << 24, "synthetic 1 M-triangle random soup (seed 12345), incoherent in-box rays, traversal only, full-stack"),
}
DEFAULT_WORKLOAD = "dragon1m"
HBM_FALLBACK_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md, used only if MEASURED_PEAKS.json is absent
L2_BYTES = 126.5e6          # cudaDeviceProp::l2CacheSize on this pool's B200 (profiles/r1_l2_peak.json)
MODE_NAMES = {"fullstack": 0, "shortstack": 1, "ref_t1": 2, "ref_t2": 3, "brute": 4, "brute_bbox": 5}


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def profile_json(name):
    try:
        return json.load(open(os.path.join(ROOT, "profiles", name)))
    except Exception:
        return {}


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy kernel)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def peak_l2():
    """L2-resident read bandwidth measured on this pool's B200 with tools/l2_peak.cu (profiles/r1_l2_peak.json)"""
    try:
        z = profile_json("r1_l2_peak.json")["read_gbs"]
        return max(v for k, v in z.items() if int(k.split()[0]) <= 64), "profiles/r1_l2_peak.json (ld.global.cg read kernel, best of 8-64 MB buffers)"
    except Exception:
        return None, "unmeasured"


def roofline_object(workload, kernel, bytes_per_launch, launch_ms, working_set_bytes, extra, units_per_launch=None):
    """The contract's roofline object for the dominant kernel. `achieved` = algorithmic bytes of THIS design's traversal
    (profiles/algorithmic_bytes.json: node records, chunk boxes, triangles of the visited leaves, hit records, ray / path
    state — counted by running the product's own traversal functions on the host over the workload's rays) per launch
    divided by the launch's CUDA-event duration. The bound follows the working set: trees that fit the 126 MB L2 are held
    against the measured L2 read bandwidth, larger ones against the measured HBM bandwidth; both fractions are printed."""
    hbm, hbm_src = peak_hbm()
    l2, l2_src = peak_l2()
    achieved = bytes_per_launch / (launch_ms * 1e-3) / 1e9
    in_l2 = working_set_bytes <= L2_BYTES and l2 is not None
    peak = l2 if in_l2 else hbm
    tr = profile_json("r2_traffic.json").get(workload, {})
    # ncu's DRAM bytes of the same kernels, per launch: the capture's bytes per ray(-bounce) x the rays of one of this run's launches
    traffic = tr.get("dram_bytes_per_unit") * units_per_launch if (tr.get("dram_bytes_per_unit") and units_per_launch) else tr.get("dram_bytes_per_launch")
    roof = {"bound": "l2" if in_l2 else "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_of": tr.get("kernel"), "kernel": kernel, "avg_launch_ms": launch_ms,
            "algorithmic_bytes_per_launch": bytes_per_launch, "working_set_bytes": int(working_set_bytes),
            "peak_source": l2_src if in_l2 else hbm_src, "frac_of_hbm": achieved / hbm, "frac_of_l2": (achieved / l2) if l2 else None,
            "hbm_peak_gbs": hbm, "l2_read_peak_gbs": l2,
            "ncu": {k: tr.get(k) for k in ("dram_gbs", "dram_frac_of_peak", "l2_gbs", "l2_frac_of_peak", "l1_hit_pct", "l2_hit_pct",
                                           "issue_active_pct", "occupancy_pct", "lanes_per_instruction", "dram_bytes_per_unit", "source") if k in tr}}
    roof.update(extra)
    return roof


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


# ---- scene set-up -----------------------------------------------------------------------------------------------------
def oracle_lib():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
    O = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
    O.ko_render_iteration.restype = ctypes.c_long
    return O


def load_scene(workload: str, product: bool = True):
    """The workload's scene arrays + render parameters. product=True goes through the library's own host code (camera
    set-up, KD builder); product=False (the `--impl reference` arm) uses only numpy and the oracle, so that libkdpt.so is
    never loaded by that arm."""
    import kdtreepathtraceroptimization_b200 as kdpt        # numpy dtypes + ctypes mirror; loads libkdpt.so only on use
    fixture, w, h, mode, extra, desc = WORKLOADS[workload]
    sc = kdpt.Scene.from_npz(os.path.join(ROOT, "tests", "golden", fixture)).with_resolution(w, h)
    if product:
        sc.first_frame_camera()
    else:
        cam = np.ascontiguousarray(sc.camera.copy())
        oracle_lib().ko_camera_runcuda(P(cam))
        sc.camera = cam
    if workload == "dragon1m":
        from kdtreepathtraceroptimization_b200 import meshes
        t18 = meshes.dense_dragon(os.path.join(ROOT, "tests", "golden", "dragon_100k.npz"), 2, 0.6)
        if product:
            tree = kdpt.Scene.from_triangles(t18)                 # the library's host-side KD builder (reference tree)
            sc.nodes, sc.tris = tree.nodes, tree.tris
        else:
            from conftest import Oracle
            sc.nodes, sc.tris, _ = Oracle(oracle_lib()).build_tree(t18, None, 13)
        # dragon.mtl: illum 6 -> glass (REFL 1, REFR 1, ior 1.5), scene.cpp:748-760
        m = sc.materials[-1:].copy()
        m['color'] = 1.0; m['specular_exponent'] = 1.0; m['specular_color'] = 1.0; m['hasReflective'] = 1.0
        m['hasRefractive'] = 1.0; m['indexOfRefraction'] = 1.5; m['transmittance'] = 0.0; m['emittance'] = 0.0
        sc.materials = np.concatenate([sc.materials[:-1], m])
    params = kdpt.default_params(mode=MODE_NAMES[mode], **extra)
    return sc, params, desc


# ---- CPU side: the reference's own host functions (oracle/_ref) or, if that file is absent, the oracle port ----------
def cpu_iteration_runner(sc, params):
    """returns (kind, fn(iter, stride, threads) -> (ray_bounces, seconds)): ONE iteration of the render loop over every
    `stride`-th pixel of the frame (ray generation included in the time) on `threads` host threads"""
    from conftest import KO_PATH, ko_params, ko_scene
    O = oracle_lib()
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
    shortstack = 1 if params.mode in (1, 3, 4, 5) else 0
    paths = np.zeros(n, KO_PATH)

    def run(it, stride=1, threads=None):
        threads = threads or (os.cpu_count() or 1)
        t0 = time.perf_counter()
        # generateRayFromCamera has no host form in the reference (a __global__ kernel): the oracle's restatement makes the rays
        O.ko_generate_rays(ctypes.byref(ks.cam), it, sc.trace_depth, ctypes.c_float(params.focalLength),
                           ctypes.c_float(params.dofAngle), params.antialias, P(paths))
        sub = np.ascontiguousarray(paths[::stride]) if stride > 1 else paths
        if R is not None:
            secs = np.zeros(2, 'f8')
            tot = R.kdref_cpu_iteration(P(sc.nodes), len(sc.nodes), P(sc.tris), P(sc.geoms), len(sc.geoms), P(sc.materials),
                                        len(sc.materials), P(sc.offsets), P(sub), len(sub), it, ctypes.c_float(params.softness),
                                        params.enableSss, params.compaction, shortstack, params.sortByMaterial, P(img), threads, P(secs))
        else:
            if stride > 1:
                raise RuntimeError("the oracle port renders whole frames only")
            tot = O.ko_render_iteration(ctypes.byref(ks), ctypes.byref(kp), it, P(img), None, threads)
        return int(tot), time.perf_counter() - t0
    return ("reference" if R is not None else "port"), run


def cpu_stride_for(workload):
    """pixel stride of the bounded CPU sample: a full 1080p iteration of the 1.6 M dragon takes the reference's host functions
    ~45 s on 16 cores"""
    return 17 if workload == "dragon1m" else 1


def algorithmic_bytes_per_ray(workload):
    """the committed per-ray figures (profiles/algorithmic_bytes.json, regenerated by tools/algorithmic_bytes.py)"""
    return json.load(open(os.path.join(ROOT, "profiles", "algorithmic_bytes.json")))[workload]


def soup_rays(sc, nrays, seed):
    lo, hi = sc.nodes['mins'][0].astype('f8'), sc.nodes['maxs'][0].astype('f8')
    rs = np.random.default_rng(seed)
    d = rs.normal(size=(nrays, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.concatenate([lo + rs.random((nrays, 3)) * (hi - lo), d], 1).astype('f4')


def run_trace_workload(args):
    """BASELINE configs[4]: traversal-only microbenchmark, rays sharded in contiguous chunks, no collective."""
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    ntri, nrays, desc = TRACE_WORKLOADS[args.workload]
    from kdtreepathtraceroptimization_b200 import meshes
    from conftest import KO_CTR, KO_HIT, KO_STACK_T1, KO_T1, Oracle, assert_hits_equal
    cores = os.cpu_count() or 1
    W = max(args.warmup, 3); K = args.steps
    if args.impl == "reference":
        if rank != 0:
            return
        # the reference's traverseKDbare semantics (its own function overflows nodeIDs[4000] on this 32 767-node tree:
        # SURVEY 8c) through the guarded restatement in the oracle; tree from the oracle's restatement of the reference builder.
        # Each step is a bounded sample (20 000 rays) of the batch: the whole 16.7 M-ray batch would take minutes per step.
        O = oracle_lib()
        nodes, tris, _ = Oracle(O).build_tree(meshes.triangle_soup(ntri), None, 13)
        import kdtreepathtraceroptimization_b200 as kdpt
        sc = kdpt.Scene.tree_only(nodes, tris)
        nr = 20000
        rays = soup_rays(sc, nr, 777)
        out = np.zeros(nr, KO_HIT); ctr = np.zeros(1, KO_CTR)
        secs = 0.0
        K = min(K, 8)
        for k in range(K):
            t0 = time.perf_counter()
            O.ko_trace_batch(KO_T1, P(nodes), len(nodes), P(tris), len(tris), 1, P(rays), nr, None, P(out), P(ctr), cores)
            secs += time.perf_counter() - t0
        v = K * nr / secs / 1e6
        print(json.dumps({"impl": "reference", "metric": "Mrays/s (traverse+intersect)", "value": v, "unit": "Mrays/s", "n_gpus": args.gpus,
                          "steps": K, "warmup": 0, "ms_per_step": 1e3 * secs / K, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": desc, "rays_per_step": nr},
                          "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port",
                                           "sample": f"{nr} rays of the workload's distribution per step, {K} steps"},
                          "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return
    import torch
    import torch.distributed as dist
    import kdtreepathtraceroptimization_b200 as kdpt
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sc = kdpt.Scene.from_triangles(meshes.triangle_soup(ntri))
    rays = soup_rays(sc, nrays, 777 + rank)
    t0 = time.perf_counter()
    ctx = kdpt.Context(sc, device=local)
    create_s = time.perf_counter() - t0
    ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
    d_rays = torch.from_numpy(rays).cuda(); d_hits = torch.zeros(nrays, 8, dtype=torch.float32, device="cuda")
    mode = MODE_NAMES[args.mode or "fullstack"]
    for _ in range(W):
        ctx.trace_device(d_rays.data_ptr(), None, nrays, d_hits.data_ptr(), mode, ts.cuda_stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = ClockSampler(local)
    time.sleep(0.3)
    clocks.rows.clear()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    launches0 = ctx.counters()[1]
    for k in range(K):           # ~1 GB of tree + 0.9 GB of rays/hits per step >> 126 MB L2: no flush needed
        evs[k][0].record(ts)
        ctx.trace_device(d_rays.data_ptr(), None, nrays, d_hits.data_ptr(), mode, ts.cuda_stream)
        evs[k][1].record(ts)
    torch.cuda.synchronize()
    clk = clocks.stop()
    ms = sum(a.elapsed_time(b) for a, b in evs)
    launches = ctx.counters()[1] - launches0      # per step: bin keys, bin scan, bin scatter, traversal
    hrays = torch.from_numpy(rays).pin_memory().numpy()
    hhits = torch.empty(nrays * kdpt.HIT.itemsize, dtype=torch.uint8).pin_memory().numpy().view(kdpt.HIT)
    t0 = time.perf_counter()
    for k in range(3):
        kdpt._check(ctx._lib.kdpt_trace(ctx._h, hrays.ctypes.data, None, nrays, hhits.ctypes.data, mode), "kdpt_trace")
    e2e_s = (time.perf_counter() - t0) / 3
    t = torch.tensor([ms, e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_s = float(t[0]), float(t[1])
    check = {"ok": None}
    if rank == 0:
        value = world * nrays * K / (ms * 1e-3) / 1e6
        ab = algorithmic_bytes_per_ray(args.workload)
        O = oracle_lib()
        # what was traced is checked: a sample of this run's hits against the oracle's stack traversal (== brute force)
        pick = np.linspace(0, nrays - 1, 1024).astype(np.int64)
        check = {"what": "1024 of this run's hits vs the oracle's stack traversal, bit for bit", "ok": True}
        try:
            if mode == 0:
                oh = Oracle(O).trace(KO_STACK_T1, sc.nodes, sc.tris, rays[pick], None, 1, nthreads=cores)
                assert_hits_equal(hhits[pick], oh, sc.tris, "bench sample")
            else:
                check["ok"] = None
        except AssertionError as e:
            check = {"ok": False, "error": str(e)[:300]}

        def cpu_run(k, threads):
            out = np.zeros(k, KO_HIT); ctr = np.zeros(1, KO_CTR)
            t0 = time.perf_counter()
            O.ko_trace_batch(KO_T1, P(sc.nodes), len(sc.nodes), P(sc.tris), len(sc.tris), 1, P(rays[:k]), k, None, P(out), P(ctr), threads)
            return time.perf_counter() - t0
        cpu = None
        if not args.no_cpu_baseline:
            # CPU baseline: a first 20 000-ray sample sizes a second one of ~10 s (bounded by the batch); then 1 thread
            dt0 = cpu_run(20000, cores)
            kcpu = int(max(20000, min(nrays, 20000 * 10.0 / max(dt0, 1e-3))))
            dt1 = cpu_run(kcpu, cores)
            k1 = max(2000, kcpu // (2 * cores))
            dt2 = cpu_run(k1, 1)
            cpu = {"value": kcpu / dt1 / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
                   "one_thread": {"value": k1 / dt2 / 1e6, "cores": 1, "sample": "%d rays, %.1f s" % (k1, dt2)},
                   "sample": "the first %d of the same rays through the reference's traverseKDbare semantics (oracle restatement, "
                             "the reference's own function overflows nodeIDs[4000] on this tree), %.1f s" % (kcpu, dt1)}
        ws = ctx.tree_info()["device_bytes"]
        tmode = args.mode or "fullstack"
        kname = ("traceLeanKernel<%s>" if tmode in ("fullstack", "shortstack") and ctx.refined_info()["num_nodes"] > 0 else "traceKernel<%s>") % tmode
        roof = roofline_object(args.workload, kname, ab["bytes_per_ray_bounce"] * nrays, ms / K, ws,
                               {"algorithmic_bytes_per_ray": ab["bytes_per_ray_bounce"], "per_ray": {k: ab.get(k) for k in ("V", "B", "T", "H")},
                                "unculled_reference_tree_bytes_per_ray": ab.get("unculled", {}).get("bytes_per_ray_bounce")}, units_per_launch=nrays)
        line = {"metric": "Mrays/s (traverse+intersect)", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": {"workload": desc, "rays_per_step_per_gpu": nrays, "tree": ctx.tree_info(), "tree_refined": ctx.refined_info(),
                                                "traversal": args.mode or "fullstack", "create_s": create_s,
                                                "l2": "working set (tree + 0.9 GB of rays/hits) exceeds L2; no flush"},
                "clocks": clk, "gpu_launches": int(launches), "check": check,
                "e2e": {"value": world * nrays / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": nrays * 24, "d2h_bytes_per_step": nrays * 32},
                "roofline": roof}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    if rank == 0 and check.get("ok") is False:
        raise SystemExit("bench.py: the sampled hits differ from the oracle: " + check.get("error", ""))


def run_reference_arm(args):
    """The reference's own CPU implementation of the path on this box's host cores. Never loads libkdpt.so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sc, params, desc = load_scene(args.workload, product=False)
    if args.mode:
        params.mode = MODE_NAMES[args.mode]
    kind, run = cpu_iteration_runner(sc, params)
    cores = os.cpu_count() or 1
    stride = cpu_stride_for(args.workload) if kind == "reference" else 1
    K = min(args.steps, 8)
    for w in range(min(args.warmup, 1)):
        run(1, max(stride, 8) if kind == "reference" else 1)
    tot = 0; secs = 0.0
    for k in range(K):
        t, s = run(1 + k, stride)
        tot += t; secs += s
    v = tot / secs / 1e6
    w, h = (int(x) for x in sc.camera['resolution'][0])
    t1, s1 = run(1, stride * 8 if kind == "reference" else 1, 1)
    line = {"impl": "reference", "metric": "Mrays/s (traverse+intersect, full render loop)", "value": v, "unit": "Mrays/s",
            "n_gpus": args.gpus, "steps": K, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * secs / max(K, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "spp_per_step": 1.0 / stride, "resolution": [w, h], "trace_depth": sc.trace_depth,
                       "traversal": "traverseKDbareShortHybrid" if params.mode in (1, 3) else "traverseKDbare",
                       "note": ("the reference's own host functions (traverseKDbare/ShortHybrid, geoms, scatterRay) driven by "
                                "oracle/ref_shim.cu on the host cores" if kind == "reference" else "oracle port (oracle/_ref/libkdref.so absent)")
                               + "; each step is ONE sample of every %d-th pixel of the frame, ray generation included" % stride},
            "spp_per_s": K / stride / secs,
            "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": kind,
                             "one_thread": {"value": t1 / s1 / 1e6, "cores": 1, "sample": "one sample of every %d-th pixel, %.1f s" % (stride * 8, s1)},
                             "sample": f"{K} step(s) of one sample of every {stride}-th pixel of the {w}x{h} frame, {tot} ray-bounces"},
            "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="kdpt", choices=["kdpt", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS) + sorted(TRACE_WORKLOADS))
    ap.add_argument("--spp-per-step", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary measurements (reference-compatible mode line)")
    ap.add_argument("--mode", default=None, choices=["fullstack", "shortstack", "ref_t1", "ref_t2", "brute", "brute_bbox"],
                    help="override the workload's traversal mode")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 20 if args.workload in ("dragon1m",) + tuple(TRACE_WORKLOADS) else 60
    if args.workload in TRACE_WORKLOADS:
        run_trace_workload(args)
        return
    if args.impl == "reference":
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
    params.mode = MODE_NAMES[mode_name]
    w, h = (int(x) for x in sc.camera['resolution'][0])
    npix = w * h

    # pathtraceInit, once: flatten + leaf refinement + upload from pinned host arrays (the reference runs it per camera move)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).pin_memory().numpy().view(a.dtype).reshape(a.shape)
    hsc = kdpt.Scene(nodes=pin(sc.nodes), tris=pin(sc.tris), geoms=pin(sc.geoms), materials=pin(sc.materials),
                     offsets=pin(sc.offsets), camera=sc.camera, trace_depth=sc.trace_depth)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ctx = kdpt.Context(hsc, device=local)
    create_s = time.perf_counter() - t0
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

    # ---- end to end through the C ABI with HOST buffers, on the context that stays alive across frames: per step the
    # camera goes host -> device (kdpt_set_camera, which also clears the radiance buffer as pathtraceInit does), S samples
    # are rendered, the shards are summed on the device (NCCL on the bound buffer, world > 1) and the radiance buffer
    # comes back to pinned host memory on rank 0 (kdpt_get_image; the reference copies it every iteration,
    # pathtrace.cu:2631). Host wall clock around the synchronous calls, barrier before each step, max over ranks.
    himg = torch.empty((h, w, 3), dtype=torch.float32).pin_memory()
    lib = kdpt.load_library()
    cam = np.ascontiguousarray(sc.camera)
    Ke = max(3, min(K, 10))
    e2e_s = 0.0; e2e_rays = 0
    for k in range(Ke + 1):
        flush.fill_(k & 0xff); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        kdpt._check(lib.kdpt_set_camera(ctx._h, cam.ctypes.data_as(ctypes.c_void_p)), "kdpt_set_camera")
        st = ctx.render(1 + k * S * world + rank, S, world, params)
        if world > 1:
            with torch.cuda.stream(ts):
                dist.all_reduce(img)
        if rank == 0:
            kdpt._check(lib.kdpt_get_image(ctx._h, ctypes.c_void_p(himg.data_ptr())), "kdpt_get_image")
        else:
            ts.synchronize()
        dt = time.perf_counter() - t0
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
    check = {"ok": None}
    if rank == 0:
        # ---- roofline of the dominant kernels (the bounce: front + lean traversal + finish): live CUDA-event durations via TESTINGMODE ----
        roof = None; cpu = None; extras = {}
        ab = {}
        try:
            ab = algorithmic_bytes_per_ray(args.workload)
            pt = kdpt.default_params(); ctypes.memmove(ctypes.byref(pt), ctypes.byref(params), ctypes.sizeof(pt)); pt.testingmode = 1
            ctx.render(1, S, 1, pt)
            st = ctx.render(1 + S, S, 1, pt)
            nb = max(1, st["intersect_launches"])
            launch_ms = st["intersect_ms"] / nb
            bytes_per_launch = ab["bytes_per_ray_bounce"] * st["ray_bounces"] / nb
            ws = ctx.tree_info()["device_bytes"] + 2 * 56 * npix * min(S, 64)
            kname = ("frontKernel + traceQueueKernel<%s> + finishQueueKernel (one bounce, wavefront form)" if ctx.refined_info()["num_nodes"] > 0 and mode_name in ("fullstack", "shortstack")
                     else "bounceKernel<%s>") % mode_name
            roof = roofline_object(args.workload, kname, bytes_per_launch, launch_ms, ws,
                                   {"kernel_share_of_step": st["intersect_ms"] / st["total_ms"],
                                    "algorithmic_bytes_per_ray_bounce": ab["bytes_per_ray_bounce"],
                                    "per_ray_bounce": {k: ab.get(k) for k in ("V", "B", "T", "H", "traversal_bytes", "path_bytes")},
                                    "unculled_reference_tree_bytes_per_ray_bounce": ab.get("unculled", {}).get("bytes_per_ray_bounce")},
                                   units_per_launch=st["ray_bounces"] / nb)
        except Exception as e:      # never let the roofline bookkeeping kill the bench line
            roof = {"bound": "hbm", "achieved": None, "peak": peak_hbm()[0], "unit": "GB/s", "frac": None, "traffic": None, "error": repr(e)}
        # ---- what was rendered is checked: ray-bounces of iteration 1 against the oracle's count for this workload
        # (profiles/algorithmic_bytes.json; device libm vs glibc in scatterRay moves a few paths, hence 0.2 %) ----
        try:
            if "ray_bounces_iter1" in ab and args.mode is None:
                ctx.clear_image()
                got = ctx.render(1, 1, 1, params)["ray_bounces"]
                exp = ab["ray_bounces_iter1"]
                check = {"what": "ray-bounces of iteration 1 vs the CPU oracle's count", "got": int(got), "expected": int(exp),
                         "ok": bool(abs(got - exp) <= 2e-3 * exp)}
        except Exception as e:
            check = {"ok": None, "error": repr(e)}
        if not args.no_extras:
            # the mode the documented drop-in shim selects (SHORTSTACK -> KDPT_REF_T2): bit-compatible with the
            # reference's own lossy traversal
            try:
                pr = kdpt.default_params(); ctypes.memmove(ctypes.byref(pr), ctypes.byref(params), ctypes.sizeof(pr)); pr.mode = MODE_NAMES["ref_t2"]
                s2 = max(1, S // 8)
                ctx.render(1, s2, 1, pr)
                sr = ctx.render(1 + s2, s2, 1, pr)
                extras["ref_t2"] = {"mode": "ref_t2", "value": sr["ray_bounces"] / (sr["total_ms"] * 1e-3) / 1e6, "unit": "Mrays/s",
                                    "spp": s2, "ms": sr["total_ms"], "what": "the same workload in the reference-compatible mode, one GPU, device-timed"}
            except Exception as e:
                extras["ref_t2"] = {"error": repr(e)}
        if world == 1 and not args.no_cpu_baseline:
            try:
                kind, run = cpu_iteration_runner(sc, params)
                cores = os.cpu_count() or 1
                stride = cpu_stride_for(args.workload) if kind == "reference" else 1
                tot = 0; secs = 0.0; it = 0
                while secs < 10.0 and it < 400:         # ~10 s of the host arm
                    a, b = run(1 + it, stride); tot += a; secs += b; it += 1
                a1, b1 = run(1, stride * 8 if kind == "reference" else 1, 1)
                cpu = {"value": tot / secs / 1e6, "unit": "Mrays/s", "cores": cores, "kind": kind,
                       "one_thread": {"value": a1 / b1 / 1e6, "cores": 1, "sample": "one sample of every %d-th pixel (%d ray-bounces, %.1f s)" % (stride * 8, a1, b1)},
                       "sample": f"{it} sample(s) of every {stride}-th pixel of the {w}x{h} frame of the same workload ({tot} ray-bounces, {secs:.1f} s), ray generation included"}
            except Exception as e:
                cpu = {"value": None, "unit": "Mrays/s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: " + repr(e)}
        out = {"metric": "Mrays/s (traverse+intersect, full render loop)", "value": value, "unit": "Mrays/s", "n_gpus": world,
               "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f32", "data": "synthetic",
               "config": {"workload": desc, "spp_per_step": S, "resolution": [w, h], "trace_depth": sc.trace_depth,
                          "traversal": mode_name, "tree": ctx.tree_info(), "tree_refined": ctx.refined_info(), "create_s": create_s,
                          "parallelism": f"sample-sharded x{world}, tree replicated, "
                          "NCCL all-reduce of the radiance buffer per step" if world > 1 else "single GPU",
                          "l2": "flushed between timed steps (256 MiB fill, untimed)"},
               "spp_per_s": S * world * K / (ms * 1e-3),
               "clocks": clk, "gpu_launches": int(launches), "check": check,
               "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": int(kdpt.CAMERA.itemsize + ctypes.sizeof(params)),
                       "d2h_bytes_per_step": int(npix * 12), "steps": Ke, "spp_per_s": S * world * Ke / e2e_s,
                       "one_time_create_s": create_s, "one_time_h2d_bytes": int(ctx.tree_info()["device_bytes"]),
                       "timing": "host wall clock around kdpt_set_camera + kdpt_render + (NCCL all-reduce on the bound device buffer) + "
                                 "kdpt_get_image into pinned host memory on rank 0; the context (tree) stays resident, as it does between "
                                 "frames in the reference"},
               "roofline": roof}
        out.update(extras)
        if cpu is not None:
            out["cpu_baseline"] = cpu
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out), flush=True)
        if check.get("ok") is False:
            raise SystemExit("bench.py: rendered ray-bounce count differs from the oracle's: %r" % (check,))


if __name__ == "__main__":
    main()
