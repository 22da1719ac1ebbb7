"""Regenerate profiles/algorithmic_bytes.json: the per-ray figures of SURVEY §8(d) for every bench workload, counted by
the instrumented CPU oracle (near-first stack traversal on the reference tree, NO leaf culling):
    bytes(ray) = 32 (ray in) + 16 (hit out) + 16 V + 4 T + 48 T      V = nodes read, T = triangle tests
plus, for the render-loop workloads, 2 x 56 B of path state and 24 B of accumulation per path-bounce.
Development tool (uses oracle/); bench.py only reads the committed JSON."""
import ctypes, json, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import kdtreepathtraceroptimization_b200 as kdpt
from kdtreepathtraceroptimization_b200 import meshes
from conftest import KO_CTR, KO_HIT, KO_STACK_T1, P

subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
O = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
out = {"_comment": __doc__.strip().splitlines()[0] + " (tools/algorithmic_bytes.py)"}
for w in bench.WORKLOADS:
    sc, params, desc = bench.load_scene(w)


    ab = bench.algorithmic_bytes_from_oracle(sc, params)
    out[w] = ab
    print(w, ab, flush=True)
for w, (ntri, nrays, desc) in bench.TRACE_WORKLOADS.items():
    sc = kdpt.Scene.from_triangles(meshes.triangle_soup(ntri))
    lo, hi = sc.nodes['mins'][0].astype('f8'), sc.nodes['maxs'][0].astype('f8')
    rs = np.random.default_rng(777)
    k = 20000
    d = rs.normal(size=(k, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([lo + rs.random((k, 3)) * (hi - lo), d], 1).astype('f4')
    res = np.zeros(k, KO_HIT); ctr = np.zeros(1, KO_CTR)
    O.ko_trace_batch(KO_STACK_T1, P(sc.nodes), len(sc.nodes), P(sc.tris), len(sc.tris), 1, P(rays), k, None, P(res), P(ctr), os.cpu_count())
    V = float(ctr['nodes'][0]) / k; T = float(ctr['tris'][0]) / k
    out[w] = dict(V=V, T=T, traversal_bytes=48 + 16 * V + 52 * T, path_bytes=0, bytes_per_ray_bounce=48 + 16 * V + 52 * T, sample="20 000 seeded in-box rays")
    print(w, out[w], flush=True)
json.dump(out, open(os.path.join(ROOT, "profiles", "algorithmic_bytes.json"), "w"), indent=1)
