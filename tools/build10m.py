import sys, ctypes
sys.path.insert(0, '.')
import kdtreepathtraceroptimization_b200 as kdpt
from kdtreepathtraceroptimization_b200 import meshes
lib = kdpt.load_library()
t18 = meshes.triangle_soup(10_000_000)
for rep in range(2):
    h = ctypes.c_void_p(); ms = ctypes.c_float()
    rc = lib.kdpt_scene_from_triangles_gpu(t18.ctypes.data_as(ctypes.c_void_p), None, len(t18), 13, 0, ctypes.byref(ms), ctypes.byref(h))
    lib.kdpt_scene_free(h)
    print("rc", rc, "device ms", ms.value, flush=True)
