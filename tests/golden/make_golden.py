"""Generate the golden fixtures in this directory from the UNMODIFIED reference (oracle/_ref/libkdref.so, built by
oracle/Makefile from /root/reference where it lies). Run in the authoring container only:

    python tests/golden/make_golden.py            # all scenes
    python tests/golden/make_golden.py one <scene.txt> <mesh.obj> <out.npz>

Each scene is produced in its own process because the reference's node-ID counter is a file-static that is never
reset (KDnode.cpp:2). What is stored per scene:
  nodes / tris / geoms / materials / offsets / camera / trace_depth  - the arrays pathtraceInit reads (reference loader)
  rays, t_min_in                                                      - seeded test rays (three classes)
  ref_T1_*, ref_T2_*                                                  - outputs of the reference's own host functions
                                                                        traverseKDbare / traverseKDbareShortHybrid
  prim_*                                                              - intersectAABBarrays / glm::intersectRayTriangle
                                                                        known answers
  rng_*, geom_*, scatter_*                                            - makeSeededRandomEngine draws, cube/sphere tests,
                                                                        scatterRay known answers
"""
import ctypes, os, subprocess, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
SCENES = [("cornell.txt", "sphere_low_1.obj", "cornell_sphere_low_1.npz", 4000),
          ("cornell8.txt", "stanford_bunny.obj", "cornell8_stanford_bunny.npz", 4000),
          ("cornell.txt", "dragon_5.obj", "cornell_dragon_5.npz", 3000),
          ("cornell.txt", "chair.obj", "cornell_chair.npz", 2000),          # two OBJ shapes
          ("cornell.txt", "hazelnut.obj", "cornell_hazelnut.npz", 2000)]

NODE = np.dtype([('axis', 'i4'), ('splitPos', 'f4'), ('mins', 'f4', 3), ('maxs', 'f4', 3), ('ID', 'i4'), ('parentID', 'i4'),
                 ('leftID', 'i4'), ('rightID', 'i4'), ('triIdStart', 'i4'), ('triIdSize', 'i4'), ('tmin', 'f4'), ('tmax', 'f4')])
TRI = np.dtype([(k, 'f4') for k in "x1 x2 x3 y1 y2 y3 z1 z2 z3 nx1 nx2 nx3 ny1 ny2 ny3 nz1 nz2 nz3".split()] + [('mtlIdx', 'i4')])
GEOM = np.dtype([('type', 'i4'), ('materialid', 'i4'), ('translation', 'f4', 3), ('rotation', 'f4', 3), ('scale', 'f4', 3),
                 ('transform', 'f4', 16), ('inverseTransform', 'f4', 16), ('invTranspose', 'f4', 16)])
MAT = np.dtype([('color', 'f4', 3), ('specular_exponent', 'f4'), ('specular_color', 'f4', 3), ('hasReflective', 'f4'),
                ('hasRefractive', 'f4'), ('indexOfRefraction', 'f4'), ('emittance', 'f4'), ('transmittance', 'f4', 3)])
CAM = np.dtype([('resolution', 'i4', 2), ('position', 'f4', 3), ('lookAt', 'f4', 3), ('view', 'f4', 3), ('up', 'f4', 3),
                ('right', 'f4', 3), ('fov', 'f4', 2), ('pixelLength', 'f4', 2)])
P = lambda a: a.ctypes.data_as(ctypes.c_void_p)


def make_rays(rng, lo, hi, n):
    c = (lo + hi) / 2
    rad = np.linalg.norm(hi - lo) / 2

    def dirs(k):
        d = rng.normal(size=(k, 3))
        return d / np.linalg.norm(d, axis=1, keepdims=True)
    oA = c + dirs(n) * rad * 2.5                      # outside the root box, aimed into it
    tgt = lo + rng.random((n, 3)) * (hi - lo)
    dA = tgt - oA
    dA /= np.linalg.norm(dA, axis=1, keepdims=True)
    oB = lo + rng.random((n, 3)) * (hi - lo)          # inside the root box (what secondary rays look like)
    dB = dirs(n)
    m = max(n // 10, 8)                               # axis-aligned directions: 1/0 = inf, 0*inf = NaN paths
    oC = lo + rng.random((m, 3)) * (hi - lo)
    dC = np.zeros((m, 3))
    dC[np.arange(m), rng.integers(0, 3, m)] = rng.choice([-1.0, 1.0], m)
    rays = np.concatenate([np.concatenate([oA, dA], 1), np.concatenate([oB, dB], 1), np.concatenate([oC, dC], 1)])
    return np.ascontiguousarray(rays, dtype=np.float32)


def one(scene, obj, out, nrays):
    R = ctypes.CDLL(os.path.join(ROOT, "oracle/_ref/libkdref.so"))
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)
    R.kdref_load(f"{REF}/scenes/{scene}".encode(), f"{REF}/scenes/{obj}".encode(), f"{REF}/scenes/".encode())
    os.dup2(saved, 1)
    s = (ctypes.c_int * 9)()
    R.kdref_sizes(s)
    nn, nt, ng, nm, nshape, depth, iters, rx, ry = list(s)
    ss = (ctypes.c_int * 8)()
    R.kdref_struct_sizes(ss)
    assert list(ss) == [NODE.itemsize, TRI.itemsize, GEOM.itemsize, MAT.itemsize, CAM.itemsize, 56, 20, 32], list(ss)
    nodes = np.zeros(nn, NODE); tris = np.zeros(nt, TRI); geoms = np.zeros(ng, GEOM); mats = np.zeros(nm, MAT)
    offs = np.zeros(max(nshape, 1), 'i4'); cam = np.zeros(1, CAM)
    R.kdref_copy(P(nodes), P(tris), P(geoms), P(mats), P(offs), P(cam))
    offs = offs[:nshape]
    # scene-file materials never initialise Material::transmittance (scene.cpp:236-271, Appendix A #8): the loader
    # leaves stack garbage there. The fixture pins it to zero for the scene-file materials (indices < offs[0]).
    nscene = int(offs[0]) if nshape > 0 else nm
    mats['transmittance'][:nscene] = 0
    rng = np.random.default_rng(12345)
    rays = make_rays(rng, nodes['mins'][0].astype('f8'), nodes['maxs'][0].astype('f8'), nrays)
    n = len(rays)
    rad = float(np.linalg.norm(nodes['maxs'][0] - nodes['mins'][0]))
    t_min_in = np.where(rng.random(n) < 0.5, np.float32(np.finfo('f4').max), rng.uniform(0, rad, n)).astype('f4')
    res = {}
    for mode, name in ((0, "T1"), (1, "T2")):
        for tname, tm in (("", None), ("_tmin", t_min_in)):
            hit = np.zeros(n, 'u1'); tmin = np.zeros(n, 'f4'); nrm = np.zeros((n, 3), 'f4'); bary = np.zeros((n, 3), 'f4')
            om = np.zeros(n, 'i4'); gi = np.zeros(n, 'i4')
            R.kdref_traverse(mode, P(nodes), nn, P(tris), nm, P(offs), P(rays), n, P(tm) if tm is not None else None,
                             P(hit), P(tmin), P(nrm), P(bary), P(om), P(gi))
            res[f"ref_{name}{tname}_hit"] = hit; res[f"ref_{name}{tname}_tmin"] = tmin; res[f"ref_{name}{tname}_normal"] = nrm
            res[f"ref_{name}{tname}_bary"] = bary; res[f"ref_{name}{tname}_objmat"] = om
    # primitive known answers: boxes = node boxes, triangles = leaf triangles, against the first rays
    k = min(n, 2000)
    bidx = rng.integers(0, nn, k)
    boxes = np.ascontiguousarray(np.concatenate([nodes['mins'][bidx], nodes['maxs'][bidx]], 1), dtype='f4')
    bh = np.zeros(k, 'u1'); bd = np.zeros(k, 'f4')
    R.kdref_aabb(P(rays), P(boxes), k, P(bh), P(bd))
    tidx = rng.integers(0, nt, k)
    T = tris[tidx]
    tri9 = np.ascontiguousarray(np.stack([T['x1'], T['y1'], T['z1'], T['x2'], T['y2'], T['z2'], T['x3'], T['y3'], T['z3']], 1), dtype='f4')
    # aim half of the rays at their triangle so that hits occur
    pr = rays[:k].copy()
    cen = (tri9[:, 0:3] + tri9[:, 3:6] + tri9[:, 6:9]) / 3
    aim = rng.random(k) < 0.6
    d = cen - pr[:, :3]
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-20)
    pr[aim, 3:] = d[aim]
    pr = np.ascontiguousarray(pr, dtype='f4')
    th = np.zeros(k, 'u1'); tb = np.zeros((k, 3), 'f4')
    R.kdref_raytri(P(pr), P(tri9), k, P(th), P(tb))
    res.update(prim_box=boxes, prim_box_hit=bh, prim_box_dist=bd, prim_tri=tri9, prim_tri_rays=pr, prim_tri_hit=th, prim_tri_bary=tb)
    # RNG known answers
    rq = np.array([[1, 0, 0], [1, 5, 1], [2, 639999, 3], [17, 123456, 7], [5000, 77, 2], [3, 1, 8]], 'i4')
    rdraw = np.zeros((len(rq), 12), 'f4'); rseed = np.zeros(len(rq), 'u4')
    for i, (it, ix, dp) in enumerate(rq):
        sd = ctypes.c_uint(0)
        R.kdref_rng(int(it), int(ix), int(dp), 12, P(rdraw[i]), ctypes.byref(sd))
        rseed[i] = sd.value
    res.update(rng_query=rq, rng_draws=rdraw, rng_seed=rseed)
    # analytic geoms: every geom against seeded rays from inside the room
    gk = 600
    go = np.stack([rng.uniform(-4, 4, gk), rng.uniform(0.5, 9.5, gk), rng.uniform(-4, 4, gk)], 1)
    gd = rng.normal(size=(gk, 3)); gd /= np.linalg.norm(gd, axis=1, keepdims=True)
    grays = np.ascontiguousarray(np.concatenate([go, gd], 1), dtype='f4')
    gt = np.zeros((ng, gk), 'f4'); gp = np.zeros((ng, gk, 3), 'f4'); gn = np.zeros((ng, gk, 3), 'f4'); gout = np.zeros((ng, gk), 'u1')
    for g in range(ng):
        R.kdref_geom_test(P(geoms[g:g + 1]), P(grays), gk, P(gt[g]), P(gp[g]), P(gn[g]), P(gout[g]))
    res.update(geom_rays=grays, geom_t=gt, geom_point=gp, geom_normal=gn, geom_outside=gout)
    # scatterRay: every material of the scene + synthetic ones, seeded normals
    extra = np.zeros(4, MAT)
    extra['color'] = [[.8, .7, .6], [.9, .9, .9], [.5, .6, .7], [.3, .9, .4]]
    extra['specular_color'] = [[1, 1, 1], [.9, .8, .7], [1, 1, 1], [1, 1, 1]]
    extra['hasReflective'] = [0.5, 1, 1, 0]; extra['hasRefractive'] = [0, 0, 1, 0]
    extra['indexOfRefraction'] = [0, 0, 1.5, 0]; extra['transmittance'] = [[0, 0, 0], [0, 0, 0], [0, 0, 0], [1, .7, .7]]
    smats = np.concatenate([mats, extra])
    sk = 300
    out_rays = []; out_cols = []; in_rays = []; in_pts = []; in_nrm = []
    for mi in range(len(smats)):
        for soft in (0.0, 0.3):
            r8 = np.zeros((sk, 8), 'f4')
            r8[:, :3] = rng.uniform(-3, 3, (sk, 3))
            dd = rng.normal(size=(sk, 3)); dd /= np.linalg.norm(dd, axis=1, keepdims=True)
            r8[:, 3:6] = dd
            r8[:, 6] = (rng.random(sk) < 0.3)
            r8[:, 7] = rng.uniform(0, 2, sk)
            nn_ = rng.normal(size=(sk, 3)); nn_ /= np.linalg.norm(nn_, axis=1, keepdims=True)
            nn_ = np.ascontiguousarray(nn_, 'f4')
            pts = np.ascontiguousarray(rng.uniform(-3, 3, (sk, 3)), 'f4')
            col = np.ones((sk, 3), 'f4')
            in_rays.append(r8.copy()); in_pts.append(pts); in_nrm.append(nn_)
            R.kdref_scatter.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int] * 3 + [ctypes.c_float, ctypes.c_int]
            R.kdref_scatter(P(r8), P(col), P(pts), P(nn_), P(smats[mi:mi + 1]), 7, 1000 * mi, 3, soft, sk)
            out_rays.append(r8); out_cols.append(col)
    res.update(scatter_mats=smats, scatter_in_rays=np.stack(in_rays), scatter_in_points=np.stack(in_pts),
               scatter_in_normals=np.stack(in_nrm), scatter_out_rays=np.stack(out_rays), scatter_out_colors=np.stack(out_cols))
    np.savez_compressed(out, nodes=nodes, tris=tris, geoms=geoms, materials=mats, offsets=offs, camera=cam,
                        trace_depth=np.int32(depth), iterations=np.int32(iters), rays=rays, t_min_in=t_min_in, **res)
    print(out, "nodes", nn, "tris", nt, "geoms", ng, "mats", nm, "rays", n, "T1 hits", int(res['ref_T1_hit'].sum()),
          "T2 hits", int(res['ref_T2_hit'].sum()), os.path.getsize(out) // 1024, "KiB")


def dragon_fixture():
    """scenes/dragon.obj (50 000 v / 100 000 f, no vn) as a compact array fixture: the seed mesh of the synthetic
    ~1.6 M-triangle dragon of BASELINE config 4 (kdtreepathtraceroptimization_b200.meshes.dense_dragon)."""
    v, f = [], []
    for l in open(f"{REF}/scenes/dragon.obj"):
        p = l.split()
        if not p:
            continue
        if p[0] == 'v':
            v.append([float(x) for x in p[1:4]])
        elif p[0] == 'f':
            idx = [int(t.split('/')[0]) - 1 for t in p[1:]]
            for k in range(1, len(idx) - 1):
                f.append([idx[0], idx[k], idx[k + 1]])
    np.savez_compressed(os.path.join(HERE, "dragon_100k.npz"), verts=np.array(v, 'f4'), faces=np.array(f, 'i4'))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        one(sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]))
    else:
        for scene, obj, out, n in SCENES:
            subprocess.check_call([sys.executable, __file__, "one", scene, obj, os.path.join(HERE, out), str(n)])
        # the reference's own golden vector (tree build): rnd/houdini/data -> rnd/houdini/dataout
        import shutil
        shutil.copy(f"{REF}/rnd/houdini/data", os.path.join(HERE, "houdini_data.txt"))
        shutil.copy(f"{REF}/rnd/houdini/dataout", os.path.join(HERE, "houdini_dataout.txt"))
        os.makedirs(os.path.join(HERE, "scenes"), exist_ok=True)
        for name in ("cornell.txt", "cornell8.txt", "sphere_low_1.obj", "sphere.mtl", "stanford_bunny.obj", "stanford_bunny.mtl", "dragon.mtl"):
            shutil.copy(f"{REF}/scenes/{name}", os.path.join(HERE, "scenes", name))
        dragon_fixture()
