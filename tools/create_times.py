"""development tool: where start-up time goes on the large configs (KD build on host vs GPU, kdpt_create = flatten + upload)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kdtreepathtraceroptimization_b200 as kdpt
from kdtreepathtraceroptimization_b200 import meshes
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
for name, t18 in (("dragon 1.6M", meshes.dense_dragon(os.path.join(G, "dragon_100k.npz"), 2, 0.6)), ("soup 10M", meshes.triangle_soup(10_000_000))):
    t = time.perf_counter(); a = kdpt.Scene.from_triangles(t18); th = time.perf_counter() - t
    t = time.perf_counter(); b = kdpt.Scene.from_triangles_gpu(t18); tg = time.perf_counter() - t
    t = time.perf_counter(); b = kdpt.Scene.from_triangles_gpu(t18); tg2 = time.perf_counter() - t
    same = a.nodes.tobytes() == b.nodes.tobytes() and a.tris.tobytes() == b.tris.tobytes()
    t = time.perf_counter(); ctx = kdpt.Context(a); tc = time.perf_counter() - t
    info = ctx.tree_info(); ctx.close()
    t = time.perf_counter(); ctx = kdpt.Context(a); tc2 = time.perf_counter() - t; ctx.close()
    print(f"{name}: host build {th:.2f} s | gpu build {tg:.2f} s then {tg2:.2f} s (device {b.build_device_ms:.0f} ms, identical={same}) | "
          f"kdpt_create {tc:.2f} s then {tc2:.2f} s | {info}", flush=True)
