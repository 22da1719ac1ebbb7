"""The drop-in binding of INTEGRATION.md, for real: oracle/dropin/pathtrace_kdpt.cpp (pathtraceInit / pathtrace /
pathtraceFree on the C ABI) compiled against the reference's own headers and linked with the reference's UNMODIFIED
Scene / KDtree / KDnode / objmesh / tinyobj objects (oracle/Makefile, target dropin_bin -> oracle/_ref/dropin_render; built
where /root/reference exists, the binary travels to the GPU box). The reference's loader and KD builder feed libkdpt.so
through the documented shim, one pathtrace() call per sample as runCuda() does, and the accumulated scene->state.image
must equal — bit for bit — what the library's own loader + Python mirror produce for the same scene."""
import json
import os
import subprocess

import numpy as np
import pytest

import kdtreepathtraceroptimization_b200 as kdpt
from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "oracle", "_ref", "dropin_render")
SCENES = os.path.join(GOLDEN, "scenes")


@pytest.mark.parametrize("txt,obj,shortstack", [("cornell8.txt", "stanford_bunny.obj", 1), ("cornell.txt", "sphere_low_1.obj", 0)])
def test_reference_host_code_plus_libkdpt_equals_the_library_alone(tmp_path, txt, obj, shortstack):
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/dropin_render not built (needs /root/reference at build time)")
    W, H, spp = 96, 64, 5
    out = tmp_path / "img.f32"
    out2 = tmp_path / "img_moved.f32"
    r = subprocess.run([EXE, os.path.join(SCENES, txt), os.path.join(SCENES, obj), SCENES + "/", str(W), str(H), str(spp), str(out),
                        str(shortstack), str(out2)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "kdpt:" not in r.stderr, r.stderr                       # the shim prints kdpt_last_error() on any failure
    info = json.loads(r.stdout.strip().splitlines()[-1])
    got = np.fromfile(out, '<f4').reshape(H, W, 3)
    sc = kdpt.Scene.load(os.path.join(SCENES, txt), os.path.join(SCENES, obj)).with_resolution(W, H).first_frame_camera()
    assert info["nodes"] == len(sc.nodes) and info["triangles"] == len(sc.tris) and info["geoms"] == len(sc.geoms)
    p = kdpt.default_params(mode=kdpt.REF_T2 if shortstack else kdpt.REF_T1, sortByMaterial=1)
    with kdpt.Context(sc) as ctx:
        ctx.clear_image()
        for it in range(1, spp + 1):
            ctx.render(it, 1, 1, p)
        want = ctx.image().reshape(H, W, 3)
        # the camera move of runCuda(): pathtraceFree + pathtraceInit with the same scene, which the shim turns into
        # kdpt_set_camera on the context it already holds
        cam = sc.camera.copy()
        cam['position'][0][0] = np.float32(cam['position'][0][0]) + np.float32(0.25)
        ctx.set_camera(cam)
        for it in range(1, spp + 1):
            ctx.render(it, 1, 1, p)
        moved = ctx.image().reshape(H, W, 3)
    moved_sc = kdpt.Scene.load(os.path.join(SCENES, txt), os.path.join(SCENES, obj)).with_resolution(W, H).first_frame_camera()
    moved_sc.camera = cam
    with kdpt.Context(moved_sc) as fresh:
        fresh.clear_image(); fresh.render(1, spp, 1, p)
        fresh_img = fresh.image().reshape(H, W, 3)
    assert want.max() > 0
    assert got.tobytes() == want.tobytes()
    got2 = np.fromfile(out2, '<f4').reshape(H, W, 3)
    assert moved.tobytes() == fresh_img.tobytes()          # kdpt_set_camera == a fresh context with that camera
    assert got2.tobytes() == moved.tobytes() and got2.tobytes() != got.tobytes()
