"""The "reference-on-B200" row: the reference's OWN CUDA kernels (src/pathtrace.cu recompiled for sm_100, default
fmad as in its own build; oracle/_ref/libkdref_gpu.so = the temp copy with one guard slot in front of `nodeIDs[]`, see
oracle/Makefile — the unmodified build, --lib libkdref.so, faults on the B200) rendering the same scene on this GPU, timed beside libkdpt.so in
the reference-compatible mode, and the two images compared. TEST / MEASUREMENT INFRASTRUCTURE: needs oracle/_ref (built
where /root/reference exists; the .so travels to the GPU box).

    python tools/ref_gpu_row.py [--scene cornell8.txt --obj stanford_bunny.obj] [--res 800x800] [--spp 32] [--check]
"""
import argparse, ctypes, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kdtreepathtraceroptimization_b200 as kdpt

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="cornell8.txt"); ap.add_argument("--obj", default="stanford_bunny.obj")
ap.add_argument("--res", default="800x800"); ap.add_argument("--spp", type=int, default=32)
ap.add_argument("--t1", action="store_true", help="SHORTSTACK off (traverseKDbare) instead of the reference's default short-stack")
ap.add_argument("--lib", default="libkdref_gpu.so", help="libkdref_gpu.so (nodeIDs guard slot) or libkdref.so (unmodified)")
ap.add_argument("--check", action="store_true", help="exit non-zero unless the images agree (RMSE < 2 %% of the mean radiance)")
args = ap.parse_args()
W, H = (int(v) for v in args.res.split("x"))
scenes = os.path.join(ROOT, "tests", "golden", "scenes")
so = os.path.join(ROOT, "oracle", "_ref", args.lib)
if not os.path.exists(so):
    print(json.dumps({"unavailable": "oracle/_ref/%s not built" % args.lib})); sys.exit(0 if not args.check else 3)
R = ctypes.CDLL(so)
P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
if not os.path.exists(os.path.join(scenes, args.obj)) and os.path.exists(os.path.join(scenes, args.obj + ".gz")):
    # the larger reference meshes are kept compressed in the fixture directory: unpack next to the scene files' copies
    import gzip, shutil, tempfile
    tmpd = tempfile.mkdtemp(prefix="kdpt_scenes_")
    for f in os.listdir(scenes):
        if not f.endswith(".gz"):
            shutil.copy(os.path.join(scenes, f), tmpd)
    with gzip.open(os.path.join(scenes, args.obj + ".gz")) as fi, open(os.path.join(tmpd, args.obj), "wb") as fo:
        shutil.copyfileobj(fi, fo)
    scenes = tmpd
devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1); os.dup2(devnull, 1)
rc = R.kdref_load(os.path.join(scenes, args.scene).encode(), os.path.join(scenes, args.obj).encode(), (scenes + "/").encode())
os.dup2(saved, 1)
assert rc == 0, "kdref_load failed"
s = (ctypes.c_int * 9)(); R.kdref_sizes(s)
nn, nt, ng, nm, nshape, depth, iters, rx, ry = list(s)
nodes = np.zeros(nn, kdpt.NODE); tris = np.zeros(nt, kdpt.TRI); geoms = np.zeros(ng, kdpt.GEOM); mats = np.zeros(nm, kdpt.MATERIAL)
offs = np.zeros(max(nshape, 1), 'i4'); cam = np.zeros(1, kdpt.CAMERA)
R.kdref_copy(P(nodes), P(tris), P(geoms), P(mats), P(offs), P(cam))
# exactly what the reference's own loader produced, uninitialised Material::transmittance included (both sides see the same bytes)
sc = kdpt.Scene(nodes=nodes, tris=tris, geoms=geoms, materials=mats, offsets=offs[:nshape].copy(), camera=cam, trace_depth=depth, iterations=iters)
sc = sc.with_resolution(W, H).first_frame_camera()
camb = np.ascontiguousarray(sc.camera)
N = args.spp
mode = kdpt.REF_T1 if args.t1 else kdpt.REF_T2
R.kdref_gpu_render.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_float,
                               ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
ref = np.zeros((H, W, 3), 'f4')
os.dup2(devnull, 1)
rc = R.kdref_gpu_render(P(camb), 1, 2, 6.0, 0.0, 1, 0.0, 0, 1, 0 if args.t1 else 1, 0, P(ref))          # warm-up (context, thrust temp)
t0 = time.perf_counter()
rc2 = R.kdref_gpu_render(P(camb), 1, N, 6.0, 0.0, 1, 0.0, 0, 1, 0 if args.t1 else 1, 0, P(ref))
t_ref = time.perf_counter() - t0
# once more with the reference's TESTINGMODE: it prints the summed CUDA-event time of its intersect kernel per iteration
# (pathtrace.cu:2565-2574, 2612-2616 — the number all of the reference's published benchmarks are made of)
import tempfile
tf = tempfile.TemporaryFile(mode="w+b"); os.dup2(tf.fileno(), 1)
ref2 = np.zeros((H, W, 3), 'f4')
rc3 = R.kdref_gpu_render(P(camb), 1, N, 6.0, 0.0, 1, 0.0, 0, 1, 0 if args.t1 else 1, 1, P(ref2))
ctypes.CDLL(None).fflush(None)
os.dup2(saved, 1)
tf.seek(0)
ref_isect_ms = [float(l.split(b"=")[1]) for l in tf.read().splitlines() if b"pathTrace time" in l]
assert rc == 0 and rc2 == 0 and rc3 == 0, "reference GPU render failed"
p = kdpt.default_params(mode=mode, sortByMaterial=1)
with kdpt.Context(sc) as ctx:
    ctx.clear_image(); ctx.render(1, 2, 1, p)                         # warm-up
    ctx.clear_image()
    t0 = time.perf_counter(); st = ctx.render(1, N, 1, p); img = ctx.image(); t_own = time.perf_counter() - t0
    pt = kdpt.default_params(mode=mode, sortByMaterial=1, testingmode=1)
    ctx.clear_image(); stt = ctx.render(1, N, 1, pt)
    # the exact-hit mode of the same traversal family (what a user who does not need the reference's dropped hits runs)
    px = kdpt.default_params(mode=kdpt.FULLSTACK if args.t1 else kdpt.SHORTSTACK, sortByMaterial=1, testingmode=1)
    ctx.clear_image(); ctx.render(1, 2, 1, px); ctx.clear_image(); stx = ctx.render(1, N, 1, px)
    # the reference's cadence: one call and one image copy per sample (what the INTEGRATION.md shim does)
    ctx.clear_image()
    t0 = time.perf_counter()
    for it in range(1, N + 1):
        ctx.render(it, 1, 1, p); img1 = ctx.image()
    t_own1 = time.perf_counter() - t0
img = img.reshape(H, W, 3)
rays = st["ray_bounces"]
mean = float(img.mean())
rmse = float(np.sqrt(((img.astype('f8') - ref) ** 2).mean()))
close = float((np.abs(img - ref) <= 1e-3 * N + 1e-3 * np.abs(ref)).all(axis=2).mean())
out = {"scene": args.scene, "obj": args.obj, "resolution": [W, H], "spp": N, "mode": "ref_t1" if args.t1 else "ref_t2",
       "ray_bounces": int(rays),
       "lib": args.lib,
       "reference_gpu_testingmode": {"intersect_ms_per_iteration": float(np.mean(ref_isect_ms)) if ref_isect_ms else None,
                                     "iterations": len(ref_isect_ms),
                                     "mrays_per_s_intersect_kernel": (rays / N) / (np.mean(ref_isect_ms) * 1e-3) / 1e6 if ref_isect_ms else None,
                                     "what": "the reference's own TESTINGMODE number: CUDA-event time of pathTraceOneBounceKDbare summed over the bounces of one iteration"},
       "libkdpt_testingmode_same_mode": {"intersect_ms_per_iteration": stt["intersect_ms"] / N,
                                         "mrays_per_s_intersect_kernel": stt["ray_bounces"] / (stt["intersect_ms"] * 1e-3) / 1e6,
                                         "what": "kdpt_params.testingmode: CUDA-event time of bounceKernel (intersect + scatter + shade + gather fused) per iteration"},
       "libkdpt_testingmode_exact_mode": {"intersect_ms_per_iteration": stx["intersect_ms"] / N,
                                          "mrays_per_s_intersect_kernel": stx["ray_bounces"] / (stx["intersect_ms"] * 1e-3) / 1e6,
                                          "mrays_per_s_whole_loop": stx["ray_bounces"] / (stx["total_ms"] * 1e-3) / 1e6},
       "reference_gpu": {"seconds": t_ref, "mrays_per_s": rays / t_ref / 1e6, "spp_per_s": N / t_ref,
                         "what": "pathtraceInit + N x pathtrace() + pathtraceFree of the UNMODIFIED src/pathtrace.cu (sm_100 build) on this GPU"},
       "libkdpt_same_mode": {"seconds": t_own, "mrays_per_s": rays / t_own / 1e6, "spp_per_s": N / t_own, "what": "kdpt_render(1..N) + kdpt_get_image, host wall clock"},
       "libkdpt_reference_cadence": {"seconds": t_own1, "mrays_per_s": rays / t_own1 / 1e6, "what": "N x (kdpt_render(it, 1) + kdpt_get_image), as the INTEGRATION.md shim"},
       "image_agreement": {"rmse_over_mean": rmse / mean if mean > 0 else None, "pixels_within_1e-3_per_sample": close,
                           "bit_identical_to_per_sample_calls": bool(img1.tobytes() == img.tobytes()),
                           "note": "the reference build contracts FMAs, libkdpt does not (its parity target is the reference's host arithmetic): a few "
                                   "rays change their hit, the images agree statistically"}}
print(json.dumps(out))
if args.check and not (mean > 0 and rmse / mean < 0.02):
    sys.exit(4)
