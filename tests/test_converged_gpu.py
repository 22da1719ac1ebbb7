"""Converged-image parity (north star: "converged images at fixed RNG seeds match to a stated RMSE"): BASELINE configs 2
and 3 at their full 800x800 resolution, 64 samples per pixel, the library's image against the CPU oracle's render of the
same 64 iterations (same seeds: the RNG is a pure function of iteration, path index and depth).

Stated bars, on the images divided by the sample count (radiance per sample), with the same number of ray-bounces to 0.1 %:
  configs[1] bunny   (diffuse + SSS, DEPTH 4): RMSE < 1e-3 and < 5 % of the mean radiance, >= 99 % of the pixels within 1e-3
  configs[2] dragon5 (glass dragon, DOF, DEPTH 8): RMSE < 8 % of the mean radiance
Device libm vs glibc in scatterRay (sin/cos/acos/pow) differ in the last ulp; a diffuse path absorbs that, a refraction
chain through the glass dragon amplifies it until the path leaves through another triangle — such a path then differs for
the rest of its bounces while the estimate stays unbiased. Measured on the B200 (round 2): bunny RMSE 1.9e-4 = 2.2 % of a
mean radiance of 0.009, 100 % of the pixels within 1e-3, identical ray-bounce count; dragon5 RMSE 9.3e-3 = 5.1 % of a mean
radiance of 0.18 (two independent 64-sample renders would differ by ~10x that), 26 % of the pixels within 1e-3, ray-bounce
counts 161 450 675 vs 161 451 276."""
import ctypes
import os

import numpy as np
import pytest

import bench
import kdtreepathtraceroptimization_b200 as kdpt
from conftest import P, ko_params, ko_scene

pytestmark = pytest.mark.gpu
SPP = 64


@pytest.mark.parametrize("workload,abs_bar,rel_bar,near_bar", [("bunny", 1e-3, 0.05, 0.99), ("dragon5", 1.0, 0.08, 0.15)])
def test_converged_image_matches_oracle(oracle, workload, abs_bar, rel_bar, near_bar):
    sc, params, desc = bench.load_scene(workload)
    w, h = (int(x) for x in sc.camera['resolution'][0])
    assert (w, h) == (800, 800)
    with kdpt.Context(sc) as ctx:
        ctx.clear_image()
        st = ctx.render(1, SPP, 1, params)
        g = ctx.image().astype('f8')
    o = np.zeros((h, w, 3), 'f4')
    ks, kp = ko_scene(sc), ko_params(params, sc.trace_depth)
    total = 0
    for it in range(1, SPP + 1):
        total += oracle.lib.ko_render_iteration(ctypes.byref(ks), ctypes.byref(kp), it, P(o), None, os.cpu_count() or 8)
    assert abs(int(st['ray_bounces']) - int(total)) <= 1e-3 * total
    g /= SPP
    o = o.astype('f8') / SPP
    mean = float(o.mean())
    rmse = float(np.sqrt(np.mean((g - o) ** 2)))
    near = float((np.abs(g - o).max(2) < 1e-3).mean())
    print(f"{workload}: {SPP} spp, RMSE = {rmse:.3e} ({rmse / mean:.4f} of the mean radiance {mean:.4f}), pixels within 1e-3 = {near:.4f}, "
          f"ray-bounces {st['ray_bounces']} vs {total}")
    assert rmse < abs_bar and rmse < rel_bar * mean, (rmse, mean)
    assert near > near_bar
