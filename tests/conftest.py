import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
SCENE_FILES = ["cornell_sphere_low_1.npz", "cornell8_stanford_bunny.npz", "cornell_dragon_5.npz",
               "cornell_chair.npz", "cornell_hazelnut.npz"]      # chair: two OBJ shapes


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


KO_HIT = np.dtype([('hit', 'i4'), ('tri', 'i4'), ('t_min', 'f4'), ('z', 'f4'), ('u', 'f4'), ('v', 'f4'), ('normal', 'f4', 3),
                   ('point', 'f4', 3), ('last_bary', 'f4', 3), ('objMaterialIdx', 'i4'), ('matAccepted', 'i4')])
KO_CTR = np.dtype([('nodes', 'u8'), ('leaves', 'u8'), ('tris', 'u8'), ('rays', 'u8'), ('maxstack', 'u8')])
KO_T1, KO_T2, KO_BRUTE_T1, KO_BRUTE_T2, KO_STACK_T1, KO_STACK_T2, KO_DFS_T1, KO_DFS_T2 = range(8)


class Oracle:
    """ctypes view of oracle/liboracle.so — the CPU restatement, used only as the checker."""

    def __init__(self, lib):
        self.lib = lib
        if hasattr(lib, "ko_render_iteration"):
            lib.ko_render_iteration.restype = ctypes.c_long
            lib.ko_collect_bounce_rays.restype = ctypes.c_long
            lib.ko_render_iteration_counted.restype = ctypes.c_long

    def trace(self, mode, nodes, tris, rays, t_min_in=None, material_size=1, nthreads=8, counters=False):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 6)
        n = len(rays)
        out = np.zeros(n, KO_HIT)
        ctr = np.zeros(1, KO_CTR)
        tm = np.ascontiguousarray(t_min_in, dtype=np.float32) if t_min_in is not None else None
        self.lib.ko_trace_batch(mode, P(nodes), len(nodes), P(tris), len(tris), material_size, P(rays), n,
                                P(tm) if tm is not None else None, P(out), P(ctr), nthreads)
        return (out, ctr[0]) if counters else out

    def build_tree(self, tri18, mtl=None, maxdepth=13):
        from kdtreepathtraceroptimization_b200 import NODE, TRI
        tri18 = np.ascontiguousarray(tri18, dtype=np.float32).reshape(-1, 18)
        pn, pt, ps = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        n1, n2 = ctypes.c_int(), ctypes.c_int()
        m = np.ascontiguousarray(mtl, dtype=np.int32) if mtl is not None else None
        self.lib.ko_build_tree(P(tri18), P(m) if m is not None else None, len(tri18), maxdepth, ctypes.byref(pn),
                               ctypes.byref(n1), ctypes.byref(pt), ctypes.byref(n2), ctypes.byref(ps))
        nodes = np.ctypeslib.as_array(ctypes.cast(pn, ctypes.POINTER(ctypes.c_uint8)), (max(n1.value, 1) * 64,)).view(NODE)[:n1.value].copy()
        tris = np.ctypeslib.as_array(ctypes.cast(pt, ctypes.POINTER(ctypes.c_uint8)), (max(n2.value, 1) * 76,)).view(TRI)[:n2.value].copy()
        src = np.ctypeslib.as_array(ctypes.cast(ps, ctypes.POINTER(ctypes.c_int32)), (max(n2.value, 1),))[:n2.value].copy() if ps.value else np.zeros(0, 'i4')
        self.lib.ko_free.argtypes = [ctypes.c_void_p]
        for p in (pn, pt, ps):
            if p.value:
                self.lib.ko_free(p)
        return nodes, tris, src


def canonical_index(tris):
    """first TriBare index holding identical 76 bytes (what kdpt_hit.tri reports)"""
    _, first, inv = np.unique(np.ascontiguousarray(tris).view('V76'), return_index=True, return_inverse=True)
    return first[inv.reshape(-1)]


@pytest.fixture(scope="session")
def oracle():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
    return Oracle(ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so")))


@pytest.fixture(scope="session")
def hostsim():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "hostsim")])
    return ctypes.CDLL(os.path.join(ROOT, "tests", "hostsim", "_build", "libkdpt_hostsim.so"))


@pytest.fixture(scope="session", params=SCENE_FILES)
def golden(request):
    return np.load(os.path.join(GOLDEN, request.param))


@pytest.fixture(scope="session")
def golden_all():
    return {f: np.load(os.path.join(GOLDEN, f)) for f in SCENE_FILES}


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_hits_equal(kh, oh, tris, what=""):
    """kdpt_hit[] (product) vs ko_hit[] (oracle): bit-exact index, and t/z/u/v/normal compared bit-for-bit
    (tolerance 0; the north star allows 1e-5 relative, we hold the stricter bar)."""
    canon = canonical_index(tris)
    oc = np.where(oh['hit'] > 0, canon[np.maximum(oh['tri'], 0)], -1)
    bad = oc != kh['tri']
    assert not bad.any(), f"{what}: {bad.sum()} hit indices differ, first at {np.nonzero(bad)[0][:5]}"
    assert (bits(oh['t_min']) == bits(kh['t'])).all(), f"{what}: t differs"
    h = oh['hit'] > 0
    for k in ('z', 'u', 'v'):
        assert (bits(oh[k])[h] == bits(kh[k])[h]).all(), f"{what}: {k} differs"
    assert (bits(oh['normal'])[h] == bits(kh['normal'])[h]).all(), f"{what}: normal differs"


# ---- oracle render-loop helpers -----------------------------------------------------------------------------------
class KoCamera(ctypes.Structure):
    _fields_ = [("res", ctypes.c_int * 2), ("position", ctypes.c_float * 3), ("lookAt", ctypes.c_float * 3),
                ("view", ctypes.c_float * 3), ("up", ctypes.c_float * 3), ("right", ctypes.c_float * 3),
                ("fov", ctypes.c_float * 2), ("pixelLength", ctypes.c_float * 2)]


class KoScene(ctypes.Structure):
    _fields_ = [("nodes", ctypes.c_void_p), ("numNodes", ctypes.c_int), ("tris", ctypes.c_void_p), ("numTris", ctypes.c_int),
                ("geoms", ctypes.c_void_p), ("numGeoms", ctypes.c_int), ("mats", ctypes.c_void_p), ("numMats", ctypes.c_int),
                ("cam", KoCamera)]


class KoRenderParams(ctypes.Structure):
    _fields_ = [("focalLength", ctypes.c_float), ("dofAngle", ctypes.c_float), ("softness", ctypes.c_float),
                ("antialias", ctypes.c_int), ("enableSss", ctypes.c_int), ("compaction", ctypes.c_int),
                ("shortstack", ctypes.c_int), ("sortMode", ctypes.c_int), ("traceDepth", ctypes.c_int)]


KO_PATH = np.dtype([('origin', 'f4', 3), ('direction', 'f4', 3), ('isinside', 'u1'), ('_pad', 'u1', 3), ('sdepth', 'f4'),
                    ('color', 'f4', 3), ('pixelIndex', 'i4'), ('remainingBounces', 'i4'), ('materialIdHit', 'i4')])
KO_ISECT = np.dtype([('t', 'f4'), ('normal', 'f4', 3), ('materialId', 'i4')])
# kdpt_mode -> oracle traversal mode / ko_render_params.shortstack
KDPT_TO_KO = {0: KO_STACK_T1, 1: KO_STACK_T2, 2: KO_T1, 3: KO_T2}
KDPT_TO_KO_SHORTSTACK = {0: 2, 1: 3, 2: 0, 3: 1}


def ko_scene(scene):
    """scene: kdtreepathtraceroptimization_b200.Scene (arrays are kept alive by the caller)"""
    s = KoScene()
    s.nodes, s.numNodes = scene.nodes.ctypes.data, len(scene.nodes)
    s.tris, s.numTris = scene.tris.ctypes.data, len(scene.tris)
    s.geoms, s.numGeoms = (scene.geoms.ctypes.data if len(scene.geoms) else None), len(scene.geoms)
    s.mats, s.numMats = scene.materials.ctypes.data, len(scene.materials)
    ctypes.memmove(ctypes.byref(s.cam), scene.camera.ctypes.data, 84)
    return s


def ko_params(p, trace_depth):
    """kdpt CParams -> ko_render_params"""
    return KoRenderParams(p.focalLength, p.dofAngle, p.softness, p.antialias, p.enableSss, p.compaction,
                          KDPT_TO_KO_SHORTSTACK[p.mode], p.sortByMaterial, trace_depth)
