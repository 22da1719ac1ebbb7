"""Traversal-only microbenchmark (kdpt_trace_device on device-resident rays) — development tool, not the bench."""
import argparse, json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kdtreepathtraceroptimization_b200 as kdpt

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="tests/golden/cornell_dragon_5.npz")
ap.add_argument("--rays", type=int, default=1 << 22)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--modes", default="fullstack,shortstack,ref_t1,ref_t2")
ap.add_argument("--cls", default="inside")
ap.add_argument("--soup", type=int, default=0, help="synthetic N-triangle soup (kdtreepathtraceroptimization_b200.meshes), tree from the library host builder")
args = ap.parse_args()
if args.soup:
    from kdtreepathtraceroptimization_b200 import meshes
    t0 = time.time(); sc = kdpt.Scene.from_triangles(meshes.triangle_soup(args.soup)); print("host build %.1fs" % (time.time() - t0))
else:
    sc = kdpt.Scene.from_npz(args.scene)
ctx = kdpt.Context(sc, 0)
print(json.dumps(ctx.tree_info()))
rng = np.random.default_rng(1)
lo, hi = sc.nodes['mins'][0], sc.nodes['maxs'][0]
n = args.rays
d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
if args.cls == "inside":
    o = lo + rng.random((n, 3)) * (hi - lo)
else:
    c = (lo + hi) / 2; rad = np.linalg.norm(hi - lo) / 2
    o = c + d * rad * 2.5
    tgt = lo + rng.random((n, 3)) * (hi - lo); d = tgt - o; d /= np.linalg.norm(d, axis=1, keepdims=True)
rays = torch.from_numpy(np.concatenate([o, d], 1).astype('f4')).cuda()
hits = torch.zeros(n, 8, dtype=torch.float32, device="cuda")
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); st = ts.cuda_stream
for name in args.modes.split(","):
    m = kdpt.MODES[name]
    ctx.trace_device(rays.data_ptr(), None, n, hits.data_ptr(), m, st)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        ctx.trace_device(rays.data_ptr(), None, n, hits.data_ptr(), m, st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    nh = int((hits[:, 4].view(torch.int32) >= 0).sum())
    print(f"{name:10s} {ms:9.3f} ms  {n / ms / 1e3:9.2f} Mrays/s  hits {nh}")
