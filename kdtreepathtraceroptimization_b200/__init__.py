"""kdtreepathtraceroptimization_b200 — host-side mirror (Python, ctypes) of the C ABI in include/kdpt.h.

The product is the shared library ``libkdpt.so`` next to this file (hand-written sm_100a kernels behind a C ABI that
replaces the reference's ``pathtraceInit / pathtrace / pathtraceFree``, src/pathtrace.h:6-21). This module only binds
it for the tests and the benchmark: numpy dtypes of the reference's POD structs, a ``Context`` object and helpers to
load a flattened scene. There is no CPU implementation here: if the library is missing or no CUDA device is usable,
every call raises.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KDPT_LIB") or os.path.join(_HERE, "libkdpt.so")   # KDPT_LIB: tuning builds only

# ---- reference POD layouts (sizes asserted against the library's static_asserts via kdpt_abi_version) -----------
NODE = np.dtype([('axis', 'i4'), ('splitPos', 'f4'), ('mins', 'f4', 3), ('maxs', 'f4', 3), ('ID', 'i4'), ('parentID', 'i4'),
                 ('leftID', 'i4'), ('rightID', 'i4'), ('triIdStart', 'i4'), ('triIdSize', 'i4'), ('tmin', 'f4'), ('tmax', 'f4')])
TRI = np.dtype([(k, 'f4') for k in "x1 x2 x3 y1 y2 y3 z1 z2 z3 nx1 nx2 nx3 ny1 ny2 ny3 nz1 nz2 nz3".split()] + [('mtlIdx', 'i4')])
GEOM = np.dtype([('type', 'i4'), ('materialid', 'i4'), ('translation', 'f4', 3), ('rotation', 'f4', 3), ('scale', 'f4', 3),
                 ('transform', 'f4', 16), ('inverseTransform', 'f4', 16), ('invTranspose', 'f4', 16)])
MATERIAL = np.dtype([('color', 'f4', 3), ('specular_exponent', 'f4'), ('specular_color', 'f4', 3), ('hasReflective', 'f4'),
                     ('hasRefractive', 'f4'), ('indexOfRefraction', 'f4'), ('emittance', 'f4'), ('transmittance', 'f4', 3)])
CAMERA = np.dtype([('resolution', 'i4', 2), ('position', 'f4', 3), ('lookAt', 'f4', 3), ('view', 'f4', 3), ('up', 'f4', 3),
                   ('right', 'f4', 3), ('fov', 'f4', 2), ('pixelLength', 'f4', 2)])
PATH = np.dtype([('origin', 'f4', 3), ('direction', 'f4', 3), ('isinside', 'u1'), ('_pad', 'u1', 3), ('sdepth', 'f4'),
                 ('color', 'f4', 3), ('pixelIndex', 'i4'), ('remainingBounces', 'i4'), ('materialIdHit', 'i4')])
ISECT = np.dtype([('t', 'f4'), ('surfaceNormal', 'f4', 3), ('materialId', 'i4')])
STAGE_INTERSECT, STAGE_SHADE = 1, 2
RAY = np.dtype([('origin', 'f4', 3), ('direction', 'f4', 3)])
HIT = np.dtype([('t', 'f4'), ('z', 'f4'), ('u', 'f4'), ('v', 'f4'), ('tri', 'i4'), ('normal', 'f4', 3)])
assert (NODE.itemsize, TRI.itemsize, GEOM.itemsize, MATERIAL.itemsize, CAMERA.itemsize, RAY.itemsize, HIT.itemsize) == \
    (64, 76, 236, 56, 84, 24, 32)

FULLSTACK, SHORTSTACK, REF_T1, REF_T2, BRUTE, BRUTE_BBOX, KDVIZ = 0, 1, 2, 3, 4, 5, 6
MODES = {"fullstack": FULLSTACK, "shortstack": SHORTSTACK, "ref_t1": REF_T1, "ref_t2": REF_T2, "brute": BRUTE, "brute_bbox": BRUTE_BBOX, "kdviz": KDVIZ}


class KdptError(RuntimeError):
    pass


class _CCamera(ctypes.Structure):
    _fields_ = [("resolution", ctypes.c_int * 2), ("position", ctypes.c_float * 3), ("lookAt", ctypes.c_float * 3),
                ("view", ctypes.c_float * 3), ("up", ctypes.c_float * 3), ("right", ctypes.c_float * 3),
                ("fov", ctypes.c_float * 2), ("pixelLength", ctypes.c_float * 2)]


class _CScene(ctypes.Structure):
    _fields_ = [("nodes", ctypes.c_void_p), ("numNodes", ctypes.c_int), ("tris", ctypes.c_void_p), ("numTris", ctypes.c_int),
                ("geoms", ctypes.c_void_p), ("numGeoms", ctypes.c_int), ("materials", ctypes.c_void_p),
                ("numMaterials", ctypes.c_int), ("objMaterialOffsets", ctypes.c_void_p), ("numShapes", ctypes.c_int),
                ("camera", _CCamera), ("traceDepth", ctypes.c_int)]


class CParams(ctypes.Structure):
    """kdpt_params — the trailing arguments of the reference's pathtrace() (defaults main.cpp:35-60)."""
    _fields_ = [("focalLength", ctypes.c_float), ("dofAngle", ctypes.c_float), ("softness", ctypes.c_float),
                ("antialias", ctypes.c_int), ("enableSss", ctypes.c_int), ("compaction", ctypes.c_int),
                ("mode", ctypes.c_int), ("sortByMaterial", ctypes.c_int), ("testingmode", ctypes.c_int),
                ("laneQueue", ctypes.c_int), ("cacheRays", ctypes.c_int), ("reserved", ctypes.c_int * 1)]


class CBuildOpts(ctypes.Structure):
    _fields_ = [("refineLeafMax", ctypes.c_int), ("refineTraversalCost", ctypes.c_float), ("nodeForm", ctypes.c_int),
                ("reserved", ctypes.c_int * 5)]


class CRenderStats(ctypes.Structure):
    _fields_ = [("rayBounces", ctypes.c_uint64), ("launches", ctypes.c_uint64), ("intersectMs", ctypes.c_float),
                ("totalMs", ctypes.c_float), ("intersectLaunches", ctypes.c_uint64)]


def default_params(**kw) -> CParams:
    p = CParams(focalLength=6.0, dofAngle=0.0, softness=0.0, antialias=1, enableSss=0, compaction=1, mode=REF_T2,
                sortByMaterial=1)
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KdptError(f"unknown render parameter {k}")
        setattr(p, k, v)
    return p


_lib = None


def load_library() -> ctypes.CDLL:
    """Load libkdpt.so; raise (never fall back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KdptError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "or `make -C kdtreepathtraceroptimization_b200/csrc` — there is no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    lib.kdpt_last_error.restype = ctypes.c_char_p
    vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    lib.kdpt_create.argtypes = [ctypes.POINTER(_CScene), ci, ctypes.POINTER(vp)]
    lib.kdpt_create_opts.argtypes = [ctypes.POINTER(_CScene), ci, ctypes.POINTER(CBuildOpts), ctypes.POINTER(vp)]
    lib.kdpt_refined_info.argtypes = [vp] + [ctypes.POINTER(ci)] * 4
    lib.kdpt_destroy.argtypes = [vp]
    lib.kdpt_node_box_geom.argtypes = [vp, vp, ci, vp]
    lib.kdpt_tree_info.argtypes = [vp] + [ctypes.POINTER(ci)] * 4 + [ctypes.POINTER(ctypes.c_uint64)]
    lib.kdpt_trace.argtypes = [vp, vp, vp, ci, vp, ci]
    lib.kdpt_trace_device.argtypes = [vp, vp, vp, ci, vp, ci, vp]
    lib.kdpt_sync.argtypes = [vp]
    lib.kdpt_render.argtypes = [vp, ci, ci, ci, ctypes.POINTER(CParams), ctypes.POINTER(CRenderStats)]
    lib.kdpt_generate_rays.argtypes = [vp, ci, ctypes.POINTER(CParams), vp]
    lib.kdpt_bounce.argtypes = [vp, ci, ci, ctypes.POINTER(CParams), vp, ci, vp, ci]
    lib.kdpt_camera_first_frame.argtypes = [vp]
    lib.kdpt_camera_first_frame.restype = None
    lib.kdpt_scene_load.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ci, ci, ctypes.POINTER(vp)]
    lib.kdpt_scene_from_triangles.argtypes = [vp, vp, ci, ci, ctypes.POINTER(vp)]
    lib.kdpt_scene_from_triangles_gpu.argtypes = [vp, vp, ci, ci, ci, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(vp)]
    lib.kdpt_scene_save.argtypes = [vp, ctypes.c_char_p]
    lib.kdpt_scene_load_cache.argtypes = [ctypes.c_char_p, ctypes.POINTER(vp)]
    lib.kdpt_scene_load_cached.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ci, ci, ctypes.c_char_p, ctypes.POINTER(vp),
                                           ctypes.POINTER(ci)]
    lib.kdpt_scene_view.argtypes = [vp, ctypes.POINTER(_CScene), ctypes.POINTER(ci)]
    lib.kdpt_scene_triangle_sources.argtypes = [vp]
    lib.kdpt_scene_triangle_sources.restype = ctypes.POINTER(ctypes.c_int)
    lib.kdpt_scene_free.argtypes = [vp]
    lib.kdpt_scene_free.restype = None
    lib.kdpt_compact_paths.argtypes = [vp, vp, ci, ctypes.POINTER(ci)]
    lib.kdpt_sort_paths.argtypes = [vp, vp, ci]
    lib.kdpt_clear_image.argtypes = [vp]
    lib.kdpt_set_camera.argtypes = [vp, vp]
    lib.kdpt_set_stream.argtypes = [vp, vp]
    lib.kdpt_bind_image.argtypes = [vp, vp]
    lib.kdpt_counters.argtypes = [vp, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]
    lib.kdpt_get_image.argtypes = [vp, vp]
    lib.kdpt_set_image.argtypes = [vp, vp]
    lib.kdpt_image_device.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(ci)]
    lib.kdpt_allreduce_image.argtypes = [ctypes.POINTER(vp), ci]
    if lib.kdpt_abi_version() != 1:
        raise KdptError("libkdpt.so ABI version mismatch")
    _lib = lib
    return lib


def _check(rc: int, what: str):
    if rc != 0:
        raise KdptError(f"{what} failed ({rc}): {load_library().kdpt_last_error().decode(errors='replace')}")


def node_box_geoms(nodes: np.ndarray, materialid: int) -> np.ndarray:
    """kdpt_node_box_geom for every node: the cube geoms KDPT_KDVIZ draws (host-only helper)"""
    lib = load_library()
    out = np.zeros(len(nodes), GEOM)
    for i in range(len(nodes)):
        mn = np.ascontiguousarray(nodes['mins'][i], 'f4'); mx = np.ascontiguousarray(nodes['maxs'][i], 'f4')
        _check(lib.kdpt_node_box_geom(mn.ctypes.data, mx.ctypes.data, int(materialid), out[i:i + 1].ctypes.data), "kdpt_node_box_geom")
    return out


@dataclass
class Scene:
    """The flat arrays pathtraceInit reads from the reference's Scene (pathtrace.cu:203-258)."""
    nodes: np.ndarray
    tris: np.ndarray
    geoms: np.ndarray
    materials: np.ndarray
    offsets: np.ndarray
    camera: np.ndarray          # CAMERA[1]
    trace_depth: int = 8
    iterations: int = 0

    @staticmethod
    def from_npz(path: str) -> "Scene":
        z = np.load(path)
        return Scene(nodes=np.ascontiguousarray(z['nodes'].view(NODE) if z['nodes'].dtype != NODE else z['nodes']),
                     tris=np.ascontiguousarray(z['tris']), geoms=np.ascontiguousarray(z['geoms']),
                     materials=np.ascontiguousarray(z['materials']), offsets=np.ascontiguousarray(z['offsets'], dtype='i4'),
                     camera=np.ascontiguousarray(z['camera']), trace_depth=int(z['trace_depth']),
                     iterations=int(z['iterations']) if 'iterations' in z else 0)

    @staticmethod
    def _from_host(h) -> "Scene":
        lib = load_library()
        v = _CScene(); it = ctypes.c_int()
        _check(lib.kdpt_scene_view(h, ctypes.byref(v), ctypes.byref(it)), "kdpt_scene_view")

        def arr(ptr, n, dt):
            if not ptr or n == 0:
                return np.zeros(0, dt)
            return np.frombuffer(ctypes.string_at(ptr, n * dt.itemsize), dtype=dt).copy()
        cam = np.frombuffer(bytes(v.camera), dtype=CAMERA).copy()
        sc = Scene(nodes=arr(v.nodes, v.numNodes, NODE), tris=arr(v.tris, v.numTris, TRI), geoms=arr(v.geoms, v.numGeoms, GEOM),
                   materials=arr(v.materials, v.numMaterials, MATERIAL), offsets=arr(v.objMaterialOffsets, v.numShapes, np.dtype('i4')),
                   camera=cam, trace_depth=int(v.traceDepth), iterations=int(it.value))
        src = lib.kdpt_scene_triangle_sources(h)
        sc.tri_source = np.ctypeslib.as_array(src, (v.numTris,)).copy() if v.numTris else np.zeros(0, 'i4')
        lib.kdpt_scene_free(h)
        return sc

    @staticmethod
    def load(scene_txt: str, obj_path: str | None = None, mtl_dir: str | None = None, kd_depth: int = 13,
             synth_normals: bool = False) -> "Scene":
        """the reference's `Scene(txt)` + `loadObj(obj)` through the library's host loader (kdpt_scene_load)"""
        h = ctypes.c_void_p()
        _check(load_library().kdpt_scene_load(scene_txt.encode(), obj_path.encode() if obj_path else None,
                                              mtl_dir.encode() if mtl_dir else None, kd_depth, int(synth_normals),
                                              ctypes.byref(h)), "kdpt_scene_load")
        return Scene._from_host(h)

    @staticmethod
    def load_cached(scene_txt: str, obj_path: str | None, cache_dir: str, mtl_dir: str | None = None, kd_depth: int = 13,
                    synth_normals: bool = False) -> "Scene":
        """`load` through the content-addressed binary cache (kdpt_scene_load_cached); `.cache_hit` says which way it went"""
        h = ctypes.c_void_p(); hit = ctypes.c_int()
        _check(load_library().kdpt_scene_load_cached(scene_txt.encode(), obj_path.encode() if obj_path else None,
                                                     mtl_dir.encode() if mtl_dir else None, kd_depth, int(synth_normals),
                                                     cache_dir.encode(), ctypes.byref(h), ctypes.byref(hit)), "kdpt_scene_load_cached")
        sc = Scene._from_host(h)
        sc.cache_hit = bool(hit.value)
        return sc

    @staticmethod
    def from_cache(path: str) -> "Scene":
        h = ctypes.c_void_p()
        _check(load_library().kdpt_scene_load_cache(path.encode(), ctypes.byref(h)), "kdpt_scene_load_cache")
        return Scene._from_host(h)

    @staticmethod
    def from_triangles(tri18: np.ndarray, mtl: np.ndarray | None = None, kd_depth: int = 13) -> "Scene":
        """KD build + flatten of a triangle list (n x 18 float32) with the library's host builder"""
        tri18 = np.ascontiguousarray(tri18, dtype=np.float32).reshape(-1, 18)
        m = np.ascontiguousarray(mtl, dtype=np.int32) if mtl is not None else None
        h = ctypes.c_void_p()
        _check(load_library().kdpt_scene_from_triangles(_ptr(tri18), _ptr(m) if m is not None else None, len(tri18), kd_depth,
                                                        ctypes.byref(h)), "kdpt_scene_from_triangles")
        return Scene._from_host(h)

    @staticmethod
    def from_triangles_gpu(tri18: np.ndarray, mtl: np.ndarray | None = None, kd_depth: int = 13, device: int = 0) -> "Scene":
        """the same KD build on the GPU (kdpt_scene_from_triangles_gpu); `.build_device_ms` holds the device time"""
        tri18 = np.ascontiguousarray(tri18, dtype=np.float32).reshape(-1, 18)
        m = np.ascontiguousarray(mtl, dtype=np.int32) if mtl is not None else None
        h = ctypes.c_void_p(); ms = ctypes.c_float()
        _check(load_library().kdpt_scene_from_triangles_gpu(_ptr(tri18), _ptr(m) if m is not None else None, len(tri18), kd_depth,
                                                            device, ctypes.byref(ms), ctypes.byref(h)), "kdpt_scene_from_triangles_gpu")
        sc = Scene._from_host(h)
        sc.build_device_ms = float(ms.value)
        return sc

    def first_frame_camera(self) -> "Scene":
        """apply the reference's main.cpp camera set-up (kdpt_camera_first_frame) and return self"""
        cam = np.ascontiguousarray(self.camera.copy())
        load_library().kdpt_camera_first_frame(cam.ctypes.data_as(ctypes.c_void_p))
        self.camera = cam
        return self

    def with_resolution(self, w: int, h: int) -> "Scene":
        """override RES the way the reference's loader derives pixelLength from it (scene.cpp:216-223)"""
        cam = self.camera.copy()
        fovy = np.float32(cam['fov'][0][1])
        yscaled = np.float32(np.tan(np.float32(fovy * np.float32(np.float32(np.pi) / np.float32(180)))))
        xscaled = np.float32(np.float32(yscaled * np.float32(w)) / np.float32(h))
        cam['resolution'][0] = (w, h)
        cam['pixelLength'][0] = (np.float32(np.float32(2) * xscaled / np.float32(w)), np.float32(np.float32(2) * yscaled / np.float32(h)))
        self.camera = cam
        return self

    @staticmethod
    def tree_only(nodes: np.ndarray, tris: np.ndarray, num_materials: int = 1) -> "Scene":
        return Scene(nodes=np.ascontiguousarray(nodes), tris=np.ascontiguousarray(tris), geoms=np.zeros(0, GEOM),
                     materials=np.zeros(num_materials, MATERIAL), offsets=np.zeros(1, 'i4'), camera=np.zeros(1, CAMERA),
                     trace_depth=8)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p) if a.size else None


class Context:
    """kdpt_ctx: pathtraceInit on creation, pathtraceFree on close()."""

    def __init__(self, scene: Scene, device: int = 0, refine_leaf_max: int = -1, refine_traversal_cost: float = 0.0, node_form: int = 0):
        """refine_leaf_max: kdpt_build_opts.refineLeafMax (-1 default, 0 = the exact modes walk the reference tree);
        node_form: kdpt_build_opts.nodeForm (0 default, 1 = 16-byte nodes, 2 = two-level records)"""
        self._lib = load_library()
        self._h = ctypes.c_void_p()
        self.scene = scene
        for name, dt in (("nodes", NODE), ("tris", TRI), ("geoms", GEOM), ("materials", MATERIAL), ("camera", CAMERA)):
            a = getattr(scene, name)
            if a.dtype.itemsize != dt.itemsize or not a.flags['C_CONTIGUOUS']:
                raise KdptError(f"scene.{name}: expected contiguous records of {dt.itemsize} bytes")
        cs = _CScene()
        cs.nodes, cs.numNodes = _ptr(scene.nodes), len(scene.nodes)
        cs.tris, cs.numTris = _ptr(scene.tris), len(scene.tris)
        cs.geoms, cs.numGeoms = _ptr(scene.geoms), len(scene.geoms)
        cs.materials, cs.numMaterials = _ptr(scene.materials), len(scene.materials)
        cs.objMaterialOffsets, cs.numShapes = _ptr(scene.offsets), len(scene.offsets)
        ctypes.memmove(ctypes.byref(cs.camera), scene.camera.ctypes.data, CAMERA.itemsize)
        cs.traceDepth = int(scene.trace_depth)
        opts = CBuildOpts(int(refine_leaf_max), float(refine_traversal_cost), int(node_form))
        _check(self._lib.kdpt_create_opts(ctypes.byref(cs), int(device), ctypes.byref(opts), ctypes.byref(self._h)), "kdpt_create_opts")
        self.device = device

    def close(self):
        if self._h:
            self._lib.kdpt_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def tree_info(self) -> dict:
        v = [ctypes.c_int() for _ in range(4)]
        b = ctypes.c_uint64()
        _check(self._lib.kdpt_tree_info(self._h, *[ctypes.byref(x) for x in v], ctypes.byref(b)), "kdpt_tree_info")
        return dict(num_nodes=v[0].value, num_leaf_tris=v[1].value, max_depth=v[2].value, max_leaf=v[3].value,
                    device_bytes=b.value)

    def refined_info(self) -> dict:
        """the tree the exact modes walk (kdpt_refined_info)"""
        v = [ctypes.c_int() for _ in range(4)]
        _check(self._lib.kdpt_refined_info(self._h, *[ctypes.byref(x) for x in v]), "kdpt_refined_info")
        return dict(num_nodes=v[0].value, num_leaf_tris=v[1].value, max_depth=v[2].value, max_leaf=v[3].value)

    def trace(self, rays: np.ndarray, t_min_in: np.ndarray | None = None, mode: int = FULLSTACK) -> np.ndarray:
        """Host buffers in, host buffers out (H2D + kernel + D2H inside the call). rays: (n,6) float32 or RAY[n]."""
        rays = np.ascontiguousarray(rays)
        n = rays.shape[0]
        if rays.dtype != RAY:
            rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(n, 6)
        hits = np.zeros(n, HIT)
        tm = None
        if t_min_in is not None:
            tm = np.ascontiguousarray(t_min_in, dtype=np.float32)
            if tm.shape != (n,):
                raise KdptError("t_min_in must have one float per ray")
        _check(self._lib.kdpt_trace(self._h, _ptr(rays), _ptr(tm) if tm is not None else None, n, _ptr(hits), int(mode)),
               "kdpt_trace")
        return hits

    def trace_device(self, d_rays: int, d_t_min_in: int | None, n: int, d_hits: int, mode: int = FULLSTACK,
                     stream: int | None = None):
        """Device pointers (e.g. torch tensor .data_ptr()); asynchronous on `stream` (cudaStream_t) or the ctx stream."""
        _check(self._lib.kdpt_trace_device(self._h, d_rays, d_t_min_in, int(n), d_hits, int(mode), stream),
               "kdpt_trace_device")

    def sync(self):
        _check(self._lib.kdpt_sync(self._h), "kdpt_sync")

    def set_stream(self, stream: int | None):
        _check(self._lib.kdpt_set_stream(self._h, stream), "kdpt_set_stream")

    def bind_image(self, d_rgb: int | None):
        _check(self._lib.kdpt_bind_image(self._h, d_rgb), "kdpt_bind_image")

    def counters(self) -> tuple[int, int]:
        a, b = ctypes.c_uint64(), ctypes.c_uint64()
        _check(self._lib.kdpt_counters(self._h, ctypes.byref(a), ctypes.byref(b)), "kdpt_counters")
        return a.value, b.value

    def render_async(self, iter_begin: int, iter_count: int, iter_stride: int = 1, params: CParams | None = None):
        """enqueue only (stats == NULL): no host synchronisation"""
        p = params if params is not None else default_params()
        _check(self._lib.kdpt_render(self._h, int(iter_begin), int(iter_count), int(iter_stride), ctypes.byref(p), None),
               "kdpt_render")

    def render(self, iter_begin: int, iter_count: int, iter_stride: int = 1, params: CParams | None = None) -> dict:
        p = params if params is not None else default_params()
        st = CRenderStats()
        _check(self._lib.kdpt_render(self._h, int(iter_begin), int(iter_count), int(iter_stride), ctypes.byref(p),
                                     ctypes.byref(st)), "kdpt_render")
        return dict(ray_bounces=st.rayBounces, launches=st.launches, intersect_ms=st.intersectMs, total_ms=st.totalMs,
                    intersect_launches=st.intersectLaunches)

    def generate_rays(self, iteration: int, params: CParams | None = None) -> np.ndarray:
        p = params if params is not None else default_params()
        w, h = (int(x) for x in self.scene.camera['resolution'][0])
        out = np.zeros(w * h, PATH)
        _check(self._lib.kdpt_generate_rays(self._h, int(iteration), ctypes.byref(p), _ptr(out)), "kdpt_generate_rays")
        return out

    def bounce(self, iteration: int, depth: int, paths: np.ndarray, stages: int = STAGE_INTERSECT | STAGE_SHADE,
               isects: np.ndarray | None = None, params: CParams | None = None):
        """one bounce on caller-owned PathSegment records; returns (paths, intersections)"""
        p = params if params is not None else default_params()
        paths = np.ascontiguousarray(paths.copy())
        assert paths.dtype.itemsize == PATH.itemsize
        isects = np.zeros(len(paths), ISECT) if isects is None else np.ascontiguousarray(isects.copy())
        _check(self._lib.kdpt_bounce(self._h, int(iteration), int(depth), ctypes.byref(p), _ptr(paths), len(paths),
                                     _ptr(isects), int(stages)), "kdpt_bounce")
        return paths, isects

    def compact_paths(self, paths: np.ndarray) -> np.ndarray:
        """kdpt_compact_paths: the render loop's stream compaction (stable remove_if remainingBounces == 0) on PathSegment records"""
        paths = np.ascontiguousarray(paths.copy())
        assert paths.dtype.itemsize == PATH.itemsize
        m = ctypes.c_int()
        _check(self._lib.kdpt_compact_paths(self._h, _ptr(paths), len(paths), ctypes.byref(m)), "kdpt_compact_paths")
        return paths[:m.value]

    def sort_paths(self, paths: np.ndarray) -> np.ndarray:
        """kdpt_sort_paths: the render loop's stable radix sort by materialIdHit on PathSegment records"""
        paths = np.ascontiguousarray(paths.copy())
        assert paths.dtype.itemsize == PATH.itemsize
        _check(self._lib.kdpt_sort_paths(self._h, _ptr(paths), len(paths)), "kdpt_sort_paths")
        return paths

    def set_camera(self, camera: np.ndarray):
        """kdpt_set_camera: new camera (CAMERA[1], resolution included), radiance buffer cleared, tree untouched"""
        cam = np.ascontiguousarray(camera)
        _check(self._lib.kdpt_set_camera(self._h, cam.ctypes.data_as(ctypes.c_void_p)), "kdpt_set_camera")
        import copy
        self.scene = copy.copy(self.scene)          # the caller's Scene object keeps its own camera
        self.scene.camera = cam.copy()

    def clear_image(self):
        _check(self._lib.kdpt_clear_image(self._h), "kdpt_clear_image")

    def image(self) -> np.ndarray:
        w, h = (int(x) for x in self.scene.camera['resolution'][0])
        out = np.zeros((h, w, 3), np.float32)
        _check(self._lib.kdpt_get_image(self._h, _ptr(out)), "kdpt_get_image")
        return out

    def set_image(self, rgb: np.ndarray):
        """restore a saved accumulation buffer (checkpoint / resume)"""
        a = np.ascontiguousarray(rgb, dtype=np.float32)
        w, h = (int(x) for x in self.scene.camera['resolution'][0])
        if a.size != w * h * 3:
            raise KdptError("set_image: expected H x W x 3 floats")
        _check(self._lib.kdpt_set_image(self._h, _ptr(a)), "kdpt_set_image")

    def image_device(self) -> tuple[int, int]:
        p = ctypes.c_void_p()
        n = ctypes.c_int()
        _check(self._lib.kdpt_image_device(self._h, ctypes.byref(p), ctypes.byref(n)), "kdpt_image_device")
        return p.value, n.value


def allreduce_images(contexts: list[Context]):
    """kdpt_allreduce_image: sum the radiance buffers of several contexts of THIS process (one per GPU) over NVLink
    peer memory and leave the total in each (single-process multi-GPU; torchrun ranks use NCCL on the bound image)."""
    if not contexts:
        raise KdptError("allreduce_images: no contexts")
    arr = (ctypes.c_void_p * len(contexts))(*[c._h for c in contexts])
    _check(contexts[0]._lib.kdpt_allreduce_image(arr, len(contexts)), "kdpt_allreduce_image")
