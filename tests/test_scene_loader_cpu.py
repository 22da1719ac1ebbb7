"""CPU-only: the product's host-side scene / OBJ loader and KD builder (csrc/kdpt_scene.cpp, pure host code) against
the arrays dumped from the UNMODIFIED reference loader (tests/golden/*.npz) — byte-for-byte — and against the oracle's
builder on random triangle soups."""
import os

import numpy as np
import pytest

import kdtreepathtraceroptimization_b200 as kdpt
from conftest import GOLDEN

SCENES = os.path.join(GOLDEN, "scenes")
CASES = [("cornell.txt", "sphere_low_1.obj", "cornell_sphere_low_1.npz", SCENES),
         ("cornell8.txt", "stanford_bunny.obj", "cornell8_stanford_bunny.npz", SCENES),
         ("cornell.txt", "dragon_5.obj", "cornell_dragon_5.npz", "/root/reference/scenes")]


@pytest.mark.parametrize("txt,obj,gold,base", CASES)
def test_loader_reproduces_the_reference_loader_byte_for_byte(txt, obj, gold, base):
    """Scene(txt) + loadObj(obj): materials, geoms (transform / inverse / inverse-transpose through glm's exact
    operation order), camera, per-shape OBJ materials, tinyobj's own decimal parser, KD build and flatten"""
    if not os.path.exists(os.path.join(base, obj)):
        pytest.skip(f"{obj} lives only in the reference checkout")
    s = kdpt.Scene.load(os.path.join(base, txt), os.path.join(base, obj))
    g = np.load(os.path.join(GOLDEN, gold))
    for name in ("nodes", "tris", "geoms", "materials"):
        assert getattr(s, name).tobytes() == g[name].tobytes(), name
    assert (s.offsets == g['offsets']).all()
    assert s.camera.tobytes() == g['camera'].tobytes()          # includes the NaN `right` of scene.cpp:221
    assert s.trace_depth == int(g['trace_depth']) and s.iterations == int(g['iterations'])


def test_builder_equals_oracle_builder_on_random_soups(oracle):
    rng = np.random.default_rng(3)
    for n, depth in ((0, 13), (1, 13), (2, 13), (3, 13), (500, 13), (20000, 13), (3000, 4), (3000, 30)):
        c = rng.uniform(-1, 1, (n, 3)); e1 = rng.uniform(-.1, .1, (n, 3)); e2 = rng.uniform(-.1, .1, (n, 3))
        t18 = np.concatenate([c, c + e1, c + e2, rng.normal(size=(n, 9))], 1).astype('f4')
        mtl = rng.integers(0, 3, n).astype('i4')
        s = kdpt.Scene.from_triangles(t18, mtl, depth)
        on, ot, src = oracle.build_tree(t18, mtl, depth)
        assert s.nodes.tobytes() == on.tobytes() and s.tris.tobytes() == ot.tobytes(), (n, depth)
        assert (s.tri_source == src).all()


def test_houdini_fixture_through_the_product_builder(tmp_path):
    """the reference's own golden vector (rnd/houdini/data -> dataout, split(30)) through the product's builder"""
    vals = np.array([float(l) for l in open(os.path.join(GOLDEN, "houdini_data.txt")).read().split()], 'f4').reshape(-1, 9)
    t18 = np.concatenate([vals, np.zeros_like(vals)], 1)
    s = kdpt.Scene.from_triangles(t18, None, 30)
    lines = ["%g %g %g %g %g %g" % (*n['mins'], *n['maxs']) for n in s.nodes]
    assert "\n".join(lines) + "\n" == open(os.path.join(GOLDEN, "houdini_dataout.txt")).read()


def test_loader_errors_and_normal_synthesis(tmp_path):
    with pytest.raises(kdpt.KdptError, match="cannot open"):
        kdpt.Scene.load(str(tmp_path / "nope.txt"))
    obj = tmp_path / "quad.obj"
    obj.write_text("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nf 1 2 3 4\n")
    txt = os.path.join(SCENES, "cornell.txt")
    with pytest.raises(kdpt.KdptError, match="normals"):
        kdpt.Scene.load(txt, str(obj))
    s = kdpt.Scene.load(txt, str(obj), synth_normals=True)
    assert len(s.tris) == 2 and len(s.nodes) == 1                  # fan triangulation, <= 2 triangles -> one leaf
    assert np.allclose([s.tris['nz1'], s.tris['nz2'], s.tris['nz3']], 1.0)
    assert len(s.materials) == 6 and s.offsets.tolist() == [5]     # default white material appended for the shape
    with pytest.raises(kdpt.KdptError, match="vertex"):
        bad = tmp_path / "bad.obj"
        bad.write_text("v 0 0 0\nvn 0 0 1\nf 1 2 3\n")
        kdpt.Scene.load(txt, str(bad))
    only = kdpt.Scene.load(txt)                                     # scene without a mesh
    assert len(only.nodes) == 0 and len(only.geoms) == 6


def _same_scene(a, b):
    for name in ("nodes", "tris", "geoms", "materials", "offsets", "camera", "tri_source"):
        assert getattr(a, name).tobytes() == getattr(b, name).tobytes(), name
    assert a.trace_depth == b.trace_depth and a.iterations == b.iterations


def test_binary_scene_cache_round_trip_and_invalidation(tmp_path):
    """kdpt_scene_load_cached: miss -> build + save, hit -> the same bytes without parsing; the key is the content of the
    scene file, the OBJ, its material library and the build options; damaged files are refused, not half-loaded"""
    import shutil
    work = tmp_path / "scene"; cache = tmp_path / "cache"
    work.mkdir(); cache.mkdir()
    for f in ("cornell8.txt", "stanford_bunny.obj", "stanford_bunny.mtl"):
        shutil.copy(os.path.join(SCENES, f), work / f)
    txt, obj = str(work / "cornell8.txt"), str(work / "stanford_bunny.obj")
    direct = kdpt.Scene.load(txt, obj)
    first = kdpt.Scene.load_cached(txt, obj, str(cache))
    assert not first.cache_hit
    files = sorted(os.listdir(cache))
    assert len(files) == 1 and files[0].startswith("kdpt-") and files[0].endswith(".scn")
    second = kdpt.Scene.load_cached(txt, obj, str(cache))
    assert second.cache_hit
    _same_scene(direct, first); _same_scene(direct, second)
    _same_scene(direct, kdpt.Scene.from_cache(str(cache / files[0])))
    # other build options, a touched material library or a touched scene file are other cache entries
    assert not kdpt.Scene.load_cached(txt, obj, str(cache), kd_depth=5).cache_hit
    with open(work / "stanford_bunny.mtl", "a") as f:
        f.write("\n# touched\n")
    assert not kdpt.Scene.load_cached(txt, obj, str(cache)).cache_hit
    with open(txt, "a") as f:
        f.write("\n")
    assert not kdpt.Scene.load_cached(txt, obj, str(cache)).cache_hit
    assert len(os.listdir(cache)) == 4
    # damage: truncation, a flipped payload byte, trailing garbage, a foreign file
    good = (cache / files[0]).read_bytes()
    for name, data in (("trunc", good[:len(good) // 2]), ("flip", good[:200] + bytes([good[200] ^ 1]) + good[201:]),
                       ("tail", good + b"x"), ("foreign", b"P6\n1 1\n255\n\0\0\0"), ("empty", b"")):
        p = tmp_path / (name + ".scn")
        p.write_bytes(data)
        with pytest.raises(kdpt.KdptError, match="not a valid scene cache"):
            kdpt.Scene.from_cache(str(p))
    with pytest.raises(kdpt.KdptError, match="cannot open"):
        kdpt.Scene.from_cache(str(tmp_path / "missing.scn"))
    # a damaged entry in the cache directory is rebuilt, not trusted
    (cache / files[0]).write_bytes(good[:100])
    os.remove(txt); shutil.copy(os.path.join(SCENES, "cornell8.txt"), txt)
    with open(work / "stanford_bunny.mtl", "w") as f:
        f.write(open(os.path.join(SCENES, "stanford_bunny.mtl")).read())
    again = kdpt.Scene.load_cached(txt, obj, str(cache))
    assert not again.cache_hit
    _same_scene(direct, again)
    assert (cache / files[0]).read_bytes() == good


def _reference_objs():
    d = "/root/reference/scenes"
    return sorted(f for f in os.listdir(d) if f.endswith(".obj")) if os.path.isdir(d) else []


@pytest.mark.parametrize("obj", _reference_objs() or ["(no reference checkout)"])
def test_loader_equals_the_reference_loader_on_every_bundled_obj(obj):
    """every .obj the reference ships (single- and multi-shape, with and without materials) through both loaders: nodes,
    triangles, geoms, materials, offsets and camera byte for byte. OBJs on which the reference itself reads out of bounds
    (fewer `vn` than `v`) are refused loudly by the product; OBJs that crash the reference loader are skipped."""
    import subprocess, sys
    from conftest import ROOT
    ref = "/root/reference/scenes"
    if not os.path.isdir(ref) or not os.path.exists(os.path.join(ROOT, "oracle/_ref/libkdref.so")):
        pytest.skip("needs /root/reference and oracle/_ref/libkdref.so")
    r = subprocess.run([sys.executable, os.path.join(GOLDEN, "compare_loader.py"), os.path.join(ref, "cornell.txt"), os.path.join(ref, obj)],
                       capture_output=True, text=True, timeout=300)
    if r.returncode != 0:
        pytest.skip("the reference loader itself fails on this file (exit %d)" % r.returncode)
    line = r.stdout.strip().splitlines()[-1]
    if "OURS FAILED" in line:
        assert "fewer normals than vertices" in line, line
    else:
        assert line.endswith("ALL EQUAL"), line


def _reference_scenes():
    d = "/root/reference/scenes"
    return sorted(f for f in os.listdir(d) if f.endswith(".txt")) if os.path.isdir(d) else []


@pytest.mark.parametrize("txt", _reference_scenes() or ["(no reference checkout)"])
def test_loader_equals_the_reference_loader_on_every_bundled_scene_file(txt):
    """every scenes/*.txt the reference ships (2-12 materials, cubes and spheres, all camera blocks) through both loaders"""
    import subprocess, sys
    from conftest import ROOT
    ref = "/root/reference/scenes"
    if not os.path.isdir(ref) or not os.path.exists(os.path.join(ROOT, "oracle/_ref/libkdref.so")):
        pytest.skip("needs /root/reference and oracle/_ref/libkdref.so")
    r = subprocess.run([sys.executable, os.path.join(GOLDEN, "compare_loader.py"), os.path.join(ref, txt), os.path.join(ref, "sphere_low_1.obj")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    assert r.stdout.strip().splitlines()[-1].endswith("ALL EQUAL"), r.stdout[-300:]
