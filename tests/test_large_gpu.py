"""GPU parity at the sizes of BASELINE configs 4 and 5 (trees far beyond the reference's own 4000-node limit, leaves of
hundreds to tens of thousands of triangles, so the chunk-culling and cooperative paths are the ones exercised).
Full-size checks are size-independent properties; a seeded sample is compared bit-for-bit with the oracle."""
import os

import numpy as np
import pytest

import kdtreepathtraceroptimization_b200 as kdpt
from kdtreepathtraceroptimization_b200 import meshes
from conftest import GOLDEN, KO_BRUTE_T1, KO_BRUTE_T2, KO_STACK_T1, KO_STACK_T2, KO_T1, KO_T2, assert_hits_equal, bits

pytestmark = pytest.mark.gpu


def _rays(sc, n, seed):
    rng = np.random.default_rng(seed)
    lo, hi = sc.nodes['mins'][0].astype('f8'), sc.nodes['maxs'][0].astype('f8')
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    inside = np.concatenate([lo + rng.random((n // 2, 3)) * (hi - lo), d[:n // 2]], 1)
    c = (lo + hi) / 2; rad = np.linalg.norm(hi - lo) / 2
    o = c + d[n // 2:] * rad * 2.0
    tgt = lo + rng.random((n - n // 2, 3)) * (hi - lo)
    dd = tgt - o; dd /= np.linalg.norm(dd, axis=1, keepdims=True)
    return np.concatenate([inside, np.concatenate([o, dd], 1)]).astype('f4')


def _check_scene(sc, oracle, nrays, nsample, min_hit, refine=-1, nbrute=None):
    """nsample rays are compared bit for bit with the oracle's stack traversal (proved equal to brute force on the CPU,
    tests/test_oracle_cpu.py), the first nbrute of each class also with brute force over every TriBare entry, and with the
    reference's own traversal semantics (T1 / T2)."""
    info = None
    rays = _rays(sc, nrays, 99)
    nbrute = nsample if nbrute is None else nbrute
    # the sample takes rays of both classes (inside / outside the root box)
    pick = np.concatenate([np.arange(nsample // 2), nrays // 2 + np.arange(nsample - nsample // 2)])
    pb = np.concatenate([np.arange(nbrute // 2), nrays // 2 + np.arange(nbrute - nbrute // 2)])
    with kdpt.Context(sc, refine_leaf_max=refine) as ctx:
        info = ctx.tree_info()
        a = ctx.trace(rays, None, kdpt.FULLSTACK)
        assert a.tobytes() == ctx.trace(rays, None, kdpt.FULLSTACK).tobytes()            # idempotent
        # batches of >= 65 536 rays are traced in leaf order (binRaysKernel), smaller ones as given: same bytes out
        m = min(nrays, 1 << 18)
        small = np.concatenate([ctx.trace(rays[k:k + 32768], None, kdpt.FULLSTACK) for k in range(0, m, 32768)])
        assert a[:m].tobytes() == small.tobytes()
        s = ctx.trace(rays, None, kdpt.SHORTSTACK)
        assert (a['tri'] != s['tri']).mean() < 2e-3                                     # differ only by the nudge (1e-5 vs 1e-4)
        r1 = ctx.trace(rays[pb], None, kdpt.REF_T1)
        r2 = ctx.trace(rays[pb], None, kdpt.REF_T2)
    h = a['tri'] >= 0
    assert h.mean() > min_hit
    # every hit re-verifies against its own triangle: the point lies on it
    T = sc.tris[a['tri'][h]]
    v0 = np.stack([T['x1'], T['y1'], T['z1']], 1).astype('f8')
    e1 = np.stack([T['x2'], T['y2'], T['z2']], 1) - v0; e2 = np.stack([T['x3'], T['y3'], T['z3']], 1) - v0
    p = rays[h, :3].astype('f8') + rays[h, 3:].astype('f8') * a['z'][h, None]
    q = v0 + e1 * a['u'][h, None] + e2 * a['v'][h, None]
    assert np.abs(p - q).max() < 1e-3
    assert (a['u'][h] >= 0).all() and (a['v'][h] >= 0).all() and (a['u'][h] + a['v'][h] <= 1).all() and (a['z'][h] >= 0).all()
    # sample: bit-exact against brute force over every TriBare entry / against the reference's traversal semantics
    nt = os.cpu_count() or 8
    assert_hits_equal(a[pick], oracle.trace(KO_STACK_T1, sc.nodes, sc.tris, rays[pick], None, 1, nthreads=nt), sc.tris, "fullstack vs stack oracle")
    assert_hits_equal(s[pick], oracle.trace(KO_STACK_T2, sc.nodes, sc.tris, rays[pick], None, 1, nthreads=nt), sc.tris, "shortstack vs stack oracle")
    assert_hits_equal(a[pb], oracle.trace(KO_BRUTE_T1, sc.nodes, sc.tris, rays[pb], None, 1, nthreads=nt), sc.tris, "fullstack")
    assert_hits_equal(s[pb], oracle.trace(KO_BRUTE_T2, sc.nodes, sc.tris, rays[pb], None, 1, nthreads=nt), sc.tris, "shortstack")
    assert_hits_equal(r1, oracle.trace(KO_T1, sc.nodes, sc.tris, rays[pb], None, 1, nthreads=nt), sc.tris, "ref_t1")
    assert_hits_equal(r2, oracle.trace(KO_T2, sc.nodes, sc.tris, rays[pb], None, 1, nthreads=nt), sc.tris, "ref_t2")
    return info


def test_config5_triangle_soup_2m_reference_tree(oracle):
    """config 5 recipe at 2 M triangles WITHOUT leaf refinement (full-depth tree: 32 767 nodes, ~170 triangles per leaf):
    the exact modes walk the reference tree itself, through the chunk-box culling and warp-cooperative leaf paths"""
    sc = kdpt.Scene.from_triangles(meshes.triangle_soup(2_000_000))
    info = _check_scene(sc, oracle, 1 << 20, 1500, 0.2, refine=0)
    assert info['num_nodes'] == 32767 and info['max_leaf'] >= 128


def test_config5_triangle_soup_10m(oracle):
    """config 5 at its full size: the 10 M-triangle soup bench.py --workload soup10m traces (refined tree, lean kernel)"""
    sc = kdpt.Scene.from_triangles(meshes.triangle_soup(10_000_000))
    info = _check_scene(sc, oracle, 1 << 21, 4096, 0.2, nbrute=256)
    assert info['num_nodes'] == 32767 and info['num_leaf_tris'] > 11_000_000


def test_config4_dense_dragon_1p6m(oracle):
    """config 4 mesh: scenes/dragon.obj subdivided twice (1.6 M triangles; the reference builder leaves a
    62 080-triangle leaf)"""
    t18 = meshes.dense_dragon(os.path.join(GOLDEN, "dragon_100k.npz"), 2, 0.6)
    tree = kdpt.Scene.from_triangles(t18)
    info = _check_scene(tree, oracle, 1 << 19, 4096, 0.1, nbrute=512)
    assert info['max_leaf'] > 10000


def test_config4_dense_dragon_1p6m_reference_tree(oracle):
    """the same mesh with refinement off: the 62 080-triangle leaf goes through the two-level chunk boxes"""
    t18 = meshes.dense_dragon(os.path.join(GOLDEN, "dragon_100k.npz"), 2, 0.6)
    tree = kdpt.Scene.from_triangles(t18)
    _check_scene(tree, oracle, 1 << 18, 600, 0.1, refine=0)
