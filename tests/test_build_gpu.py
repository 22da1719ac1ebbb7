"""KD build on the GPU (kdpt_scene_from_triangles_gpu) against the host builder, which is itself
byte-identical to the reference's KDtree/KDnode + flatten (tests/test_scene_loader_cpu.py): NodeBare[], TriBare[] and the
source-triangle map must be the same BYTES (integer / index work and copied floats: no tolerance)."""
import os

import numpy as np
import pytest

import kdtreepathtraceroptimization_b200 as kdpt
from kdtreepathtraceroptimization_b200 import meshes
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _same(t18, mtl=None, depth=13):
    a = kdpt.Scene.from_triangles(t18, mtl, depth)
    b = kdpt.Scene.from_triangles_gpu(t18, mtl, depth)
    assert len(a.nodes) == len(b.nodes) and len(a.tris) == len(b.tris)
    assert a.nodes.tobytes() == b.nodes.tobytes()
    assert a.tri_source.tobytes() == b.tri_source.tobytes()
    assert a.tris.tobytes() == b.tris.tobytes()
    return a, b


@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, 7, 33, 257, 2049, 5000])
def test_small_soups(n):
    _same(meshes.triangle_soup(n, seed=n + 1, edge=0.2).reshape(-1, 18))


@pytest.mark.parametrize("depth", [1, 2, 5, 13, 20])
def test_depths_and_material_indices(depth):
    t18 = meshes.triangle_soup(40_000, seed=7, edge=0.05)
    mtl = (np.arange(len(t18)) % 5).astype('i4')
    a, _ = _same(t18, mtl, depth)
    assert a.tris['mtlIdx'].max() == 4


def test_degenerate_inputs():
    # every triangle identical: one side keeps everything -> the root is the only node
    one = meshes.triangle_soup(1, seed=3, edge=0.3)
    a, _ = _same(np.repeat(one, 100, axis=0))
    assert len(a.nodes) == 1
    # huge triangles spanning the box (heavy duplication), mixed with small ones; zeros of both signs in the bounds
    rs = np.random.default_rng(5)
    big = meshes.triangle_soup(300, seed=11, edge=1.5)
    small = meshes.triangle_soup(3000, seed=12, edge=0.01)
    t = np.concatenate([big, small]); rs.shuffle(t)
    t[::7, 0] = 0.0; t[3::11, 0] = -0.0; t[5::13, 4] = -0.0
    _same(t)
    # axis-aligned flat mesh (all z equal): the z-levels never split
    flat = meshes.triangle_soup(2000, seed=13, edge=0.05); flat[:, [2, 5, 8]] = 0.25
    _same(flat)


def test_reference_meshes():
    """the 100 000-face scenes/dragon.obj (fixture) — the tree the reference's own loader builds"""
    t18 = meshes.dense_dragon(os.path.join(GOLDEN, "dragon_100k.npz"), 0, 1.0)
    a, _ = _same(t18)
    assert len(a.nodes) == 3597 and len(a.tris) == 141773            # SURVEY §8 tree sizes of dragon.obj


def test_soup_1m_and_traversal_on_the_gpu_built_tree():
    t18 = meshes.triangle_soup(1_000_000)
    a, b = _same(t18)
    assert len(a.nodes) == 32767
    rng = np.random.default_rng(1)
    d = rng.normal(size=(20000, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([rng.uniform(-1, 1, (20000, 3)), d], 1).astype('f4')
    with kdpt.Context(a) as ca, kdpt.Context(b) as cb:
        assert ca.trace(rays).tobytes() == cb.trace(rays).tobytes()
    assert b.build_device_ms > 0
