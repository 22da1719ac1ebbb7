"""small end-to-end case for compute-sanitizer (development tool)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kdtreepathtraceroptimization_b200 as kdpt
from kdtreepathtraceroptimization_b200 import meshes
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
for name in ("cornell8_stanford_bunny.npz", "cornell_dragon_5.npz"):
    g = np.load(os.path.join(G, name))
    sc = kdpt.Scene.from_npz(os.path.join(G, name)).with_resolution(48, 40).first_frame_camera()
    with kdpt.Context(sc) as ctx:
        for m in range(4):
            h = ctx.trace(g['rays'][:1500], g['t_min_in'][:1500], m)
        for kw in (dict(), dict(compaction=0), dict(sortByMaterial=2, mode=kdpt.SHORTSTACK), dict(mode=kdpt.FULLSTACK, enableSss=1, softness=0.2)):
            ctx.clear_image(); ctx.render(1, 3, 1, kdpt.default_params(**kw)); img = ctx.image()
        p = ctx.generate_rays(2); ctx.bounce(2, 0, p)
sc = kdpt.Scene.from_triangles(meshes.triangle_soup(60000))
with kdpt.Context(sc) as ctx:
    r = np.concatenate([np.random.default_rng(1).uniform(-1, 1, (3000, 3)), np.random.default_rng(2).normal(size=(3000, 3))], 1).astype('f4')
    for m in range(4):
        ctx.trace(r, None, m)
print("sanitize case done")
