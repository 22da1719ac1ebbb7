"""GPU parity of the render loop (generate rays -> intersect+scatter -> shade -> gather -> compact -> sort) against the
oracle's restatement of pathtrace() (oracle/kdpt_oracle.c, pinned piecewise against the reference by
tests/test_oracle_cpu.py). Integer results and everything built from + - * / sqrt are compared bit-for-bit; values that
pass through libm (sinf/cosf/acosf/powf/pow on the device vs glibc in the oracle) with the tolerances stated below."""
import ctypes
import os

import numpy as np
import pytest

import kdtreepathtraceroptimization_b200 as kdpt
from conftest import (GOLDEN, KDPT_TO_KO, KO_ISECT, KO_PATH, P, bits, ko_params, ko_scene)

pytestmark = pytest.mark.gpu
LIBM_RTOL = 1e-5      # north star: 1e-5 relative for floating point
CASES = [("cornell_sphere_low_1.npz", dict()), ("cornell8_stanford_bunny.npz", dict()),
         ("cornell_dragon_5.npz", dict(dofAngle=0.01, focalLength=6.0))]


def _scene(name, w=160, h=160):
    sc = kdpt.Scene.from_npz(os.path.join(GOLDEN, name))
    return sc.with_resolution(w, h).first_frame_camera()


@pytest.mark.parametrize("name,kw", CASES)
def test_first_frame_camera_matches_oracle(oracle, name, kw):
    sc = kdpt.Scene.from_npz(os.path.join(GOLDEN, name))
    cam = sc.camera.copy()
    oracle.lib.ko_camera_runcuda(P(cam))
    sc.first_frame_camera()
    assert cam.tobytes() == sc.camera.tobytes()


@pytest.mark.parametrize("name,kw", CASES)
@pytest.mark.parametrize("aa", [0, 1])
def test_generate_rays_bit_exact(oracle, name, kw, aa):
    """generateRayFromCamera (pathtrace.cu:315-397): AA + DOF jitter; the per-iteration constants come from the same
    glibc on both sides, the per-pixel part is + - * / sqrt only -> bit-exact"""
    sc = _scene(name)
    p = kdpt.default_params(antialias=aa, **kw)
    with kdpt.Context(sc) as ctx:
        for it in (1, 2, 977):
            g = ctx.generate_rays(it, p)
            o = np.zeros(len(g), KO_PATH)
            ks = ko_scene(sc)
            oracle.lib.ko_generate_rays(ctypes.byref(ks.cam), it, sc.trace_depth, ctypes.c_float(p.focalLength),
                                        ctypes.c_float(p.dofAngle), aa, P(o))
            for k in ('origin', 'direction', 'color', 'sdepth'):
                assert (bits(g[k]) == bits(o[k])).all(), (it, k)
            for k in ('isinside', 'pixelIndex', 'remainingBounces'):
                assert (g[k] == o[k]).all(), (it, k)


def _oracle_bounce(oracle, sc, mode, it, depth, paths, stages, softness=0.0, sss=0):
    o = paths.copy().view(KO_PATH) if paths.dtype != KO_PATH else paths.copy()
    isects = np.zeros(len(o), KO_ISECT)
    ks = ko_scene(sc)
    oracle.lib.ko_bounce.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int,
                                     ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    oracle.lib.ko_bounce(ctypes.byref(ks), KDPT_TO_KO[mode], it, depth, softness, sss, o.ctypes.data, len(o), isects.ctypes.data, stages)
    return o, isects


def _close(a, b, rtol=LIBM_RTOL, atol=1e-6):
    a = np.asarray(a, 'f8'); b = np.asarray(b, 'f8')
    return np.abs(a - b) <= atol + rtol * np.abs(b)


@pytest.mark.parametrize("name,kw", CASES)
@pytest.mark.parametrize("mode", [kdpt.FULLSTACK, kdpt.SHORTSTACK, kdpt.REF_T1, kdpt.REF_T2])
@pytest.mark.parametrize("lane_queue", [1, 2])
def test_bounce_stages_match_oracle(oracle, name, kw, mode, lane_queue):
    """pathTraceOneBounceKDbare + shadeMaterial over three consecutive bounces, feeding the ORACLE's paths to both
    sides at every bounce so that one libm ulp cannot cascade. Hit decisions (t, normal, material, remainingBounces,
    colour) are bit-exact; the scattered ray goes through device libm -> LIBM_RTOL."""
    sc = _scene(name, 128, 128)
    p = kdpt.default_params(mode=mode, laneQueue=lane_queue, **kw)     # both traversal schedules of the bounce kernel
    with kdpt.Context(sc) as ctx:
        paths = ctx.generate_rays(3, p)
        nhit = 0
        for depth in range(3):
            gp, gi = ctx.bounce(3, depth, paths, kdpt.STAGE_INTERSECT | kdpt.STAGE_SHADE, params=p)
            op, oi = _oracle_bounce(oracle, sc, mode, 3, depth, paths, 3)
            assert (bits(gi['t']) == bits(oi['t'])).all(), depth
            assert (gi['materialId'] == oi['materialId']).all()
            assert (bits(gi['surfaceNormal']) == bits(oi['normal'])).all()
            assert (gp['remainingBounces'] == op['remainingBounces']).all()
            assert (gp['materialIdHit'] == op['materialIdHit']).all()
            assert (bits(gp['color']) == bits(op['color'])).all()
            assert (gp['isinside'] == op['isinside']).all()
            live = op['remainingBounces'] > 0
            ok = _close(gp['origin'], op['origin']).all(1) & _close(gp['direction'], op['direction']).all(1) & _close(gp['sdepth'], op['sdepth'])
            assert ok[live].all(), (depth, (~ok[live]).sum())
            nhit += int((oi['t'] > 0).sum())
            paths = op.view(kdpt.PATH)
        assert nhit > 1000


def test_bounce_with_softness_and_sss(oracle):
    """soft reflections / refraction jitter (randSphericalVec, rotateVector: double-precision trig) and the SSS
    branches of scatterRay + shadeMaterial"""
    sc = _scene("cornell8_stanford_bunny.npz", 128, 128)
    p = kdpt.default_params(mode=kdpt.FULLSTACK, softness=0.3, enableSss=1)
    with kdpt.Context(sc) as ctx:
        paths = ctx.generate_rays(5, p)
        for depth in range(3):
            gp, gi = ctx.bounce(5, depth, paths, 3, params=p)
            op, oi = _oracle_bounce(oracle, sc, kdpt.FULLSTACK, 5, depth, paths, 3, softness=0.3, sss=1)
            assert (bits(gi['t']) == bits(oi['t'])).all()
            assert (gp['remainingBounces'] == op['remainingBounces']).all()
            live = op['remainingBounces'] > 0
            ok = _close(gp['origin'], op['origin'], 1e-4, 1e-5).all(1) & _close(gp['direction'], op['direction'], 1e-4, 1e-5).all(1)
            assert ok[live].mean() > 0.999
            assert _close(gp['color'], op['color'], 1e-4, 1e-6).all()
            paths = op.view(kdpt.PATH)


def _oracle_render(oracle, sc, p, iters):
    w, h = (int(x) for x in sc.camera['resolution'][0])
    img = np.zeros((h, w, 3), 'f4')
    ks = ko_scene(sc)
    kp = ko_params(p, sc.trace_depth)
    total = 0
    for it in iters:
        total += oracle.lib.ko_render_iteration(ctypes.byref(ks), ctypes.byref(kp), it, P(img), None, 8)
    return img, total


@pytest.mark.parametrize("name,kw", CASES)
@pytest.mark.parametrize("mode", [kdpt.FULLSTACK, kdpt.REF_T2])
def test_render_iterations_match_oracle(oracle, name, kw, mode):
    """whole loop incl. compaction, gather and the iteration-2 material sort. Per-pixel libm ulps can flip a rare
    decision, so the bar is: ray-bounce counts within 0.1 %, image RMSE < 2 % of the mean radiance, and >= 99.5 % of
    the pixels within 1e-3."""
    sc = _scene(name, 160, 160)
    p = kdpt.default_params(mode=mode, **kw)
    iters = [1, 2, 3]
    with kdpt.Context(sc) as ctx:
        ctx.clear_image()
        st = ctx.render(1, 3, 1, p)
        g = ctx.image()
    o, total = _oracle_render(oracle, sc, p, iters)
    assert abs(int(st['ray_bounces']) - int(total)) <= max(2, int(1e-3 * total)), (st['ray_bounces'], total)
    rmse = float(np.sqrt(np.mean((g.astype('f8') - o) ** 2)))
    assert rmse < 0.02 * max(float(o.mean()), 1e-3), rmse
    assert (np.abs(g - o).max(2) < 1e-3).mean() > 0.995


def test_render_without_compaction_and_sort_every_bounce(oracle):
    sc = _scene("cornell_sphere_low_1.npz", 128, 128)
    for kw in (dict(compaction=0, sortByMaterial=0), dict(compaction=1, sortByMaterial=2)):
        p = kdpt.default_params(mode=kdpt.SHORTSTACK, **kw)
        with kdpt.Context(sc) as ctx:
            ctx.clear_image()
            st = ctx.render(1, 2, 1, p)
            g = ctx.image()
        o, total = _oracle_render(oracle, sc, p, [1, 2])
        assert abs(int(st['ray_bounces']) - int(total)) <= max(2, int(2e-3 * total))
        assert (np.abs(g - o).max(2) < 1e-3).mean() > 0.99


def test_sample_sharding_is_a_partition():
    """iter_stride: ranks rendering {1,3,5} and {2,4,6} sum to the single-context result over 1..6 (the multi-GPU
    contract: identical samples, only the float summation order of the final reduction differs)"""
    sc = _scene("cornell8_stanford_bunny.npz", 128, 128)
    p = kdpt.default_params(mode=kdpt.FULLSTACK, sortByMaterial=0)
    with kdpt.Context(sc) as a, kdpt.Context(sc) as b, kdpt.Context(sc) as c:
        for x in (a, b, c):
            x.clear_image()
        a.render(1, 6, 1, p)
        b.render(1, 3, 2, p)
        c.render(2, 3, 2, p)
        full, parts = a.image(), b.image() + c.image()
    assert np.allclose(full, parts, rtol=1e-5, atol=1e-6)


def test_peer_allreduce_of_sample_shards():
    """kdpt_allreduce_image: two contexts of one process, one per GPU, render disjoint sample shards; after the
    peer-memory all-reduce both hold the same bits, equal to the single-GPU image up to float summation order"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    sc = _scene("cornell8_stanford_bunny.npz", 200, 150)
    p = kdpt.default_params(mode=kdpt.FULLSTACK)
    with kdpt.Context(sc, device=0) as one:
        one.clear_image(); one.render(1, 8, 1, p); full = one.image()
    with kdpt.Context(sc, device=0) as a, kdpt.Context(sc, device=1) as b:
        a.clear_image(); b.clear_image()
        a.render_async(1, 4, 2, p); b.render_async(2, 4, 2, p)
        kdpt.allreduce_images([a, b])
        ia, ib = a.image(), b.image()
    assert ia.tobytes() == ib.tobytes()
    assert np.allclose(ia, full, rtol=1e-5, atol=1e-6)
    with kdpt.Context(sc, device=0) as a, kdpt.Context(sc, device=0) as b:
        with pytest.raises(kdpt.KdptError):
            kdpt.allreduce_images([a, b])         # two contexts on one device


def test_headless_cli_matches_the_library(tmp_path):
    """kdpt_render (C++ driver over the C ABI: scene .txt + .obj -> first-frame camera -> spp iterations -> saveImage
    semantics: x mirrored, / samples, clamp*255 truncated) writes the same pixels as the Python mirror computes"""
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "kdtreepathtraceroptimization_b200", "kdpt_render")
    scenes = os.path.join(GOLDEN, "scenes")
    out = tmp_path / "bunny.ppm"
    r = subprocess.run([exe, os.path.join(scenes, "cornell8.txt"), os.path.join(scenes, "stanford_bunny.obj"), "--spp", "6",
                        "--res", "96x64", "--mode", "fullstack", "--out", str(out), "--json"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    import json
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["spp"] == 6 and info["ray_bounces"] > 96 * 64 * 6
    data = out.read_bytes()
    hdr = b"P6\n96 64\n255\n"
    assert data.startswith(hdr)
    got = np.frombuffer(data[len(hdr):], np.uint8).reshape(64, 96, 3)
    sc = kdpt.Scene.load(os.path.join(scenes, "cornell8.txt"), os.path.join(scenes, "stanford_bunny.obj")).with_resolution(96, 64).first_frame_camera()
    with kdpt.Context(sc) as ctx:
        ctx.clear_image()
        ctx.render(1, 6, 1, kdpt.default_params(mode=kdpt.FULLSTACK))
        img = ctx.image()
    want = (np.clip(img / np.float32(6), 0, 1) * np.float32(255)).astype(np.uint8)[:, ::-1, :]
    assert (got == want).all()
    png = tmp_path / "bunny.png"
    r = subprocess.run([exe, os.path.join(scenes, "cornell8.txt"), os.path.join(scenes, "stanford_bunny.obj"), "--spp", "1",
                        "--res", "32x32", "--out", str(png)], capture_output=True, text=True)
    assert r.returncode == 0 and png.read_bytes()[:8] == b"\x89PNG\r\n\x1a\n"


def test_headless_cli_checkpoint_resume_cache_and_hdr(tmp_path):
    """kdpt_render: a render interrupted after 4 of 10 samples and resumed from its checkpoint writes the same bytes as an
    uninterrupted one (iteration numbers seed the RNG; accumulation is in iteration order); the scene cache is hit on
    the second run; the .hdr output decodes to radiance / spp within RGBE precision"""
    import json, subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "kdtreepathtraceroptimization_b200", "kdpt_render")
    scenes = os.path.join(GOLDEN, "scenes")
    base = [exe, os.path.join(scenes, "cornell8.txt"), os.path.join(scenes, "stanford_bunny.obj"), "--res", "80x60", "--mode", "shortstack",
            "--cache-dir", str(tmp_path), "--json"]

    def run(*extra):
        r = subprocess.run(base + list(extra), capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        return json.loads(r.stdout.strip().splitlines()[-1])
    full = run("--spp", "10", "--out", str(tmp_path / "full.pfm"), "--checkpoint", str(tmp_path / "full.ckp"), "--checkpoint-every", "3")
    assert full["cache_hit"] == 0
    part = run("--spp", "4", "--checkpoint", str(tmp_path / "part.ckp"))
    assert part["cache_hit"] == 1
    rest = run("--spp", "10", "--resume", str(tmp_path / "part.ckp"), "--out", str(tmp_path / "resumed.pfm"), "--checkpoint", str(tmp_path / "resumed.ckp"))
    assert (tmp_path / "full.pfm").read_bytes() == (tmp_path / "resumed.pfm").read_bytes()
    assert (tmp_path / "full.ckp").read_bytes() == (tmp_path / "resumed.ckp").read_bytes()
    assert part["ray_bounces"] + rest["ray_bounces"] == full["ray_bounces"]
    r = subprocess.run(base + ["--spp", "10", "--res", "64x64", "--resume", str(tmp_path / "part.ckp")], capture_output=True, text=True)
    assert r.returncode != 0 and "not a checkpoint" in r.stderr                    # other resolution
    other = [a if a != "shortstack" else "fullstack" for a in base]
    r = subprocess.run(other + ["--spp", "10", "--resume", str(tmp_path / "part.ckp")], capture_output=True, text=True)
    assert r.returncode != 0 and "not a checkpoint" in r.stderr                    # other render parameters
    r = subprocess.run(base + ["--spp", "3", "--resume", str(tmp_path / "part.ckp")], capture_output=True, text=True)
    assert r.returncode != 0 and "already holds" in r.stderr                       # more samples in the file than asked for
    run("--spp", "10", "--resume", str(tmp_path / "full.ckp"), "--out", str(tmp_path / "img.hdr"))   # nothing left to render
    raw = (tmp_path / "img.hdr").read_bytes()
    head = b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y 60 +X 80\n"
    assert raw.startswith(head) and len(raw) == len(head) + 80 * 60 * 4
    rgbe = np.frombuffer(raw[len(head):], np.uint8).reshape(60, 80, 4).astype(np.float64)
    dec = rgbe[..., :3] * np.exp2(rgbe[..., 3:] - 136.0) * (rgbe[..., 3:] > 0)
    pfm = (tmp_path / "full.pfm").read_bytes()
    ph = b"PF\n80 60\n-1.0\n"
    want = np.frombuffer(pfm[len(ph):], '<f4').reshape(60, 80, 3)[::-1]                # PFM rows are bottom-up
    scale = np.maximum(want.max(axis=2, keepdims=True), 1e-30)
    assert (np.abs(dec - want) <= scale / 128 + 1e-30).all() and want.max() > 0


@pytest.mark.parametrize("name,kw", CASES)
@pytest.mark.parametrize("mode", [kdpt.FULLSTACK, kdpt.SHORTSTACK, kdpt.REF_T1])
def test_lane_queue_schedules_are_bit_identical(name, kw, mode):
    """kdpt_params.laneQueue only changes WHEN a path's traversal runs (at once, pooled in the per-warp queue, or the
    kernel's own per-launch choice): images and ray counts must be identical bit for bit, with and without compaction
    and with the material sort on every bounce"""
    sc = _scene(name, 160, 128)
    for extra in (dict(), dict(compaction=0, sortByMaterial=0), dict(sortByMaterial=2)):
        got = []
        for lq in (1, 2, 0):
            p = kdpt.default_params(mode=mode, laneQueue=lq, **kw, **extra)
            with kdpt.Context(sc) as ctx:
                ctx.clear_image()
                st = ctx.render(1, 20, 1, p)            # 20 iterations = two batches: the second one uses the measured fractions
                got.append((ctx.image().tobytes(), int(st['ray_bounces'])))
        assert got[0] == got[1] == got[2], (extra, [g[1] for g in got])
    with pytest.raises(kdpt.KdptError):
        with kdpt.Context(sc) as ctx:
            ctx.render(1, 1, 1, kdpt.default_params(laneQueue=3))


@pytest.mark.parametrize("name,kw", CASES)
@pytest.mark.parametrize("mode", [kdpt.FULLSTACK, kdpt.SHORTSTACK])
def test_node_forms_render_identical_images(name, kw, mode):
    """kdpt_build_opts.nodeForm only changes how the exact modes' tree is laid out for the traversal kernel of the wavefront
    bounce: images and ray counts are identical bit for bit"""
    sc = _scene(name, 160, 128)
    got = []
    for form in (1, 2):
        with kdpt.Context(sc, node_form=form) as ctx:
            ctx.clear_image()
            st = ctx.render(1, 6, 1, kdpt.default_params(mode=mode, **kw))
            got.append((ctx.image().tobytes(), int(st['ray_bounces'])))
    assert got[0] == got[1] and got[0][1] > 0


def test_checkpoint_resume_is_bit_identical():
    """render 1..4, save, restore into a fresh context, render 5..8 == render 1..8 in one go (kdpt_get_image /
    kdpt_set_image; accumulation is iteration-ordered, so the result is bit-identical)"""
    sc = _scene("cornell8_stanford_bunny.npz", 96, 96)
    p = kdpt.default_params(mode=kdpt.FULLSTACK)
    with kdpt.Context(sc) as a:
        a.clear_image(); a.render(1, 8, 1, p); full = a.image()
    with kdpt.Context(sc) as b:
        b.clear_image(); b.render(1, 4, 1, p); ckpt = b.image()
    with kdpt.Context(sc) as c:
        c.set_image(ckpt); c.render(5, 4, 1, p); resumed = c.image()
    assert full.tobytes() == resumed.tobytes()


@pytest.mark.parametrize("variant", ["no_mesh", "no_geoms", "depth1", "depth10", "odd_res", "no_compaction_depth10"])
def test_render_edge_cases_match_oracle(oracle, variant):
    """scene without a mesh, mesh without analytic geoms, DEPTH 1 and DEPTH 10 (beyond the reference's 8-bounce cap:
    survivors are dropped with compaction on, added by finalGather with it off), resolution not a multiple of 32"""
    sc = _scene("cornell_sphere_low_1.npz", 96, 64)
    kw = dict(mode=kdpt.FULLSTACK, sortByMaterial=0)
    if variant == "no_mesh":
        sc.nodes = np.zeros(0, kdpt.NODE); sc.tris = np.zeros(0, kdpt.TRI)
    elif variant == "no_geoms":
        sc.geoms = np.zeros(0, kdpt.GEOM)
    elif variant == "depth1":
        sc.trace_depth = 1
    elif variant == "depth10":
        sc.trace_depth = 10
    elif variant == "odd_res":
        sc = _scene("cornell_sphere_low_1.npz", 77, 53)
    elif variant == "no_compaction_depth10":
        sc.trace_depth = 10; kw["compaction"] = 0
    p = kdpt.default_params(**kw)
    with kdpt.Context(sc) as ctx:
        ctx.clear_image()
        st = ctx.render(1, 3, 1, p)
        g = ctx.image()
    o, total = _oracle_render(oracle, sc, p, [1, 2, 3])
    assert abs(int(st['ray_bounces']) - int(total)) <= max(2, int(2e-3 * total)), (st['ray_bounces'], total)
    assert (np.abs(g - o).max(2) < 1e-3).mean() > 0.99
    if variant != "no_geoms":
        assert o.sum() > 0


@pytest.mark.parametrize("n", [1, 31, 32, 33, 1000, 4097, 100003, 640000])
def test_compaction_and_sort_are_bit_exact(n):
    """C1 / C2 are index work: the render loop's own compaction kernels (warp-ballot counts -> scan -> stable scatter)
    and its radix sort, run on caller-owned PathSegment records through kdpt_compact_paths / kdpt_sort_paths, must
    reproduce thrust::remove_if(is_zero_bounce) and a stable sort on materialIdHit byte for byte — every field of
    every record in the same order (ragged tails, all-dead / all-alive packets, keys wider than one digit)."""
    rng = np.random.default_rng(n)
    sc = _scene("cornell_sphere_low_1.npz", 64, 64)
    with kdpt.Context(sc) as ctx:
        for variant in range(3):
            p = np.zeros(n, kdpt.PATH)
            p['origin'] = rng.normal(size=(n, 3)); p['direction'] = rng.normal(size=(n, 3)); p['color'] = rng.random((n, 3))
            p['sdepth'] = rng.random(n); p['isinside'] = rng.integers(0, 2, n)
            p['pixelIndex'] = rng.permutation(n)
            if variant == 0:
                p['remainingBounces'] = rng.integers(0, 3, n)                       # a third of the paths are finished
                p['materialIdHit'] = rng.integers(-1, 12, n)
            elif variant == 1:
                p['remainingBounces'] = (rng.random(n) < 0.03) * rng.integers(1, 8, n)   # almost everything finished, whole packets empty
                p['materialIdHit'] = rng.integers(-1, 70000, n)                     # keys that need three radix digits
            else:
                p['remainingBounces'] = 1 + rng.integers(0, 8, n)                   # nothing finished
                p['materialIdHit'] = 5
            got = ctx.compact_paths(p)
            exp = p[p['remainingBounces'] != 0]
            assert len(got) == len(exp) and got.tobytes() == exp.tobytes(), (n, variant)
            got = ctx.sort_paths(p)
            exp = p[np.argsort(p['materialIdHit'], kind='stable')]
            assert got.tobytes() == exp.tobytes(), (n, variant)


def test_ray_cache_and_brute_force_modes(oracle):
    """the two remaining switches of pathtrace(): rayCaching (pathtrace.cu:2448-2456 — every iteration starts from the camera
    rays of iteration 1) against the oracle with the same switch, and the KD-less intersect modes (ENABLEKD off / USEBBOX),
    whose images must be the bytes of the exact short-stack mode (same triangles, same nudge)"""
    sc = _scene("cornell8_stanford_bunny.npz", 96, 96)
    p = kdpt.default_params(mode=kdpt.SHORTSTACK, cacheRays=1)
    with kdpt.Context(sc) as ctx:
        ctx.clear_image(); st = ctx.render(1, 4, 1, p); cached = ctx.image()
        ctx.clear_image(); ctx.render(1, 4, 1, kdpt.default_params(mode=kdpt.SHORTSTACK)); plain = ctx.image()
        imgs = {}
        for m in (kdpt.SHORTSTACK, kdpt.BRUTE, kdpt.BRUTE_BBOX):
            ctx.clear_image(); ctx.render(1, 3, 1, kdpt.default_params(mode=m)); imgs[m] = ctx.image()
    assert not np.array_equal(cached, plain)
    oracle.lib.ko_set_ray_cache(1)
    try:
        o, total = _oracle_render(oracle, sc, p, [1, 2, 3, 4])
    finally:
        oracle.lib.ko_set_ray_cache(0)
    assert abs(int(st['ray_bounces']) - int(total)) <= max(2, int(1e-3 * total))
    assert (np.abs(cached - o).max(2) < 1e-3).mean() > 0.995
    assert imgs[kdpt.BRUTE].tobytes() == imgs[kdpt.SHORTSTACK].tobytes()
    assert imgs[kdpt.BRUTE_BBOX].tobytes() == imgs[kdpt.SHORTSTACK].tobytes()


def test_kdviz_mode(oracle):
    """KDPT_KDVIZ (VIZKD, pathTraceOneBounceKDbareBoxes pathtrace.cu:1738-1885): the scene's geoms, then every node box as a cube of
    the last material, no triangles. (a) The mode equals, bit for bit, an ordinary render of the scene rewritten that way by
    hand — geoms + kdpt_node_box_geom cubes, no tree; (b) that render equals the oracle's of the same rewritten scene (node
    cubes from the oracle's own ko_kdviz_geoms, whose test is pinned to the reference's boxIntersectionTestBox outputs) within
    the render loop's libm tolerance."""
    import dataclasses
    for name, res, iters in (("cornell8_stanford_bunny.npz", 96, 3), ("cornell_dragon_5.npz", 48, 2)):
        sc = _scene(name, res, res)
        p = kdpt.default_params(mode=kdpt.KDVIZ)
        with kdpt.Context(sc) as ctx:
            ctx.clear_image(); st = ctx.render(1, iters, 1, p); viz = ctx.image()
            ctx.clear_image(); ctx.render(1, iters, 1, kdpt.default_params(mode=kdpt.REF_T2)); plain = ctx.image()
            with pytest.raises(kdpt.KdptError):
                ctx.trace(np.zeros((4, 6), 'f4'), None, kdpt.KDVIZ)        # a render-loop mode only
        assert not np.array_equal(viz, plain)
        cubes = kdpt.node_box_geoms(sc.nodes, len(sc.materials) - 1)
        by_hand = dataclasses.replace(sc, nodes=np.zeros(0, kdpt.NODE), tris=np.zeros(0, kdpt.TRI), geoms=np.concatenate([sc.geoms, cubes]))
        q = kdpt.default_params(mode=kdpt.REF_T2)
        with kdpt.Context(by_hand) as ctx:
            ctx.clear_image(); st2 = ctx.render(1, iters, 1, q); hand = ctx.image()
        assert viz.tobytes() == hand.tobytes() and int(st['ray_bounces']) == int(st2['ray_bounces'])
        ocubes = np.zeros(len(sc.nodes), kdpt.GEOM)
        oracle.lib.ko_kdviz_geoms(P(sc.nodes), len(sc.nodes), len(sc.materials), P(ocubes))
        osc = dataclasses.replace(by_hand, geoms=np.concatenate([sc.geoms, ocubes]))
        o, total = _oracle_render(oracle, osc, q, list(range(1, iters + 1)))
        assert abs(int(st['ray_bounces']) - int(total)) <= max(2, int(2e-3 * total)), (st['ray_bounces'], total)
        rmse = float(np.sqrt(np.mean((viz.astype('f8') - o) ** 2)))
        assert rmse < 0.02 * max(float(o.mean()), 1e-3), rmse
        assert (np.abs(viz - o).max(2) < 1e-3).mean() > 0.99
