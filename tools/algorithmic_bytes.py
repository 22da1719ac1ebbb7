"""Regenerate profiles/algorithmic_bytes.json: the bytes the traversal of THIS design has to read per ray, for every
bench workload (re-based in round 2 on what the kernels actually walk — VERDICT r1 item 1b).

For every ray-bounce of iteration 1 (render workloads; the rays and the analytic-geom t_min come from the CPU oracle's
render loop) or for a seeded sample of the traversal-only workloads' rays, the product's own traversal function
(kdpt_core.cuh:traceExact on the flatten output, compiled for the host by tests/hostsim) is run with counters:

    V = node records read (16 B)           B = chunk / group boxes tested (32 B)
    T = triangles of the visited leaves / surviving chunks (40 B hot record)
    H = triangles glm accepted (40 B normals + 4 B entryOf + the hot record again is NOT counted)
    traversal bytes = 16 V + 32 B + 40 T + 44 H          (+ 48 B ray in / hit out for the traversal-only workloads)
    render workloads add 2 x 56 B of path state and 24 B of accumulation per ray-bounce

`unculled` keeps round 1's figure (SURVEY's formula on the reference tree with NO leaf culling, from the instrumented
oracle) for comparison. Development tool (uses oracle/ and tests/hostsim); bench.py only reads the committed JSON."""
import ctypes, json, os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import kdtreepathtraceroptimization_b200 as kdpt
from kdtreepathtraceroptimization_b200 import meshes
from conftest import KO_CTR, KO_HIT, KO_STACK_T1, P, ko_params, ko_scene

subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "hostsim")])
O = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
O.ko_collect_all_bounce_rays.restype = ctypes.c_long
O.ko_render_iteration_counted.restype = ctypes.c_long
HS = ctypes.CDLL(os.path.join(ROOT, "tests", "hostsim", "_build", "libkdpt_hostsim.so"))
cores = os.cpu_count() or 1
path = os.path.join(ROOT, "profiles", "algorithmic_bytes.json")
old = json.load(open(path)) if os.path.exists(path) else {}
only = sys.argv[1:]


def counted(sc, rays, tmin, mode, nmat):
    n = len(rays)
    out = np.zeros(n, kdpt.HIT); ctr = np.zeros(5, 'u8')
    HS.hostsim_set_refine(-1)
    rc = HS.hostsim_trace(P(sc.nodes), len(sc.nodes), P(sc.tris), len(sc.tris), nmat, P(rays), P(tmin) if tmin is not None else None,
                          n, P(out), mode, P(ctr), None)
    assert rc == 0
    V, _, T, B, H = (float(x) / n for x in ctr)
    info = np.zeros(4, 'i8')
    HS.hostsim_refined_info(P(sc.nodes), len(sc.nodes), P(sc.tris), len(sc.tris), -1, P(info))
    return dict(V=V, B=B, T=T, H=H, traversal_bytes=16 * V + 32 * B + 40 * T + 44 * H,
                refined_tree=dict(nodes=int(info[0]), leaf_entries=int(info[1]), max_depth=int(info[2]), max_leaf=int(info[3])))


out = dict(old)
out["_comment"] = __doc__.strip().splitlines()[0] + " (tools/algorithmic_bytes.py)"
for w in bench.WORKLOADS:
    if only and w not in only:
        continue
    t0 = time.time()
    sc, params, desc = bench.load_scene(w)
    ks, kp = ko_scene(sc), ko_params(params, sc.trace_depth)
    npix = int(sc.camera['resolution'][0][0]) * int(sc.camera['resolution'][0][1])
    cap = npix * 8
    rays = np.zeros((cap, 6), 'f4'); per = np.zeros(8, 'i8')
    n = O.ko_collect_all_bounce_rays(ctypes.byref(ks), ctypes.byref(kp), 1, P(rays), cap, P(per), cores)
    rays = rays[:n].copy()
    tmin = np.zeros(n, 'f4')
    O.ko_geoms_tmin(ctypes.byref(ks), P(rays), n, P(tmin))
    c = counted(sc, rays, tmin, params.mode if params.mode < 2 else params.mode - 2, len(sc.materials))
    c["path_bytes"] = 2 * 56 + 24
    c["bytes_per_ray_bounce"] = c["traversal_bytes"] + c["path_bytes"]
    c["ray_bounces_iter1"] = int(n); c["ray_bounces_per_bounce"] = [int(x) for x in per]
    if w in old and "unculled" in old[w]:
        c["unculled"] = old[w]["unculled"]
    elif w in old and "V" in old[w] and "B" not in old[w]:      # round 1's entry
        c["unculled"] = {k: old[w][k] for k in ("V", "T", "traversal_bytes", "bytes_per_ray_bounce")}
    out[w] = c
    print(w, "%.0f s" % (time.time() - t0), json.dumps(c), flush=True)
    json.dump(out, open(path, "w"), indent=1)
for w, (ntri, nrays, desc) in bench.TRACE_WORKLOADS.items():
    if only and w not in only:
        continue
    sc = kdpt.Scene.from_triangles(meshes.triangle_soup(ntri))
    k = 20000
    rays = bench.soup_rays(sc, k, 777)
    c = counted(sc, rays, None, 0, 1)
    c["traversal_bytes"] += 48; c["path_bytes"] = 0; c["bytes_per_ray_bounce"] = c["traversal_bytes"]
    c["sample"] = "20 000 seeded in-box rays"
    if w in old and "unculled" in old[w]:
        c["unculled"] = old[w]["unculled"]
    elif w in old and "V" in old[w] and "B" not in old[w]:
        c["unculled"] = {k2: old[w][k2] for k2 in ("V", "T", "traversal_bytes", "bytes_per_ray_bounce")}
    out[w] = c
    print(w, json.dumps(c), flush=True)
    json.dump(out, open(path, "w"), indent=1)
