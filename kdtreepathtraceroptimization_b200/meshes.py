"""Deterministic synthetic meshes for the large BASELINE configs (the reference's own large assets are missing
upstream: `.MISSING_LARGE_BLOBS`). Plain numpy, host side, not on any timed path."""
from __future__ import annotations

import numpy as np


def triangle_soup(n: int, seed: int = 12345, edge: float | None = None) -> np.ndarray:
    """BASELINE config 5: n triangles, centres uniform in [-1,1]^3, two edge vectors uniform in [-r,r]^3 (r = 0.01 at
    1 M triangles, scaled with n^(-1/3) above so the median tree can still split), geometric normals. Returns (n,18)."""
    rs = np.random.default_rng(seed)
    r = edge if edge is not None else (0.01 * (1e6 / n) ** (1 / 3) if n > 1e6 else 0.01)
    c = rs.uniform(-1, 1, (n, 3)); e1 = rs.uniform(-r, r, (n, 3)); e2 = rs.uniform(-r, r, (n, 3))
    nrm = np.cross(e1, e2)
    nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-30)
    return np.concatenate([c, c + e1, c + e2, nrm, nrm, nrm], 1).astype(np.float32)


def subdivide(verts: np.ndarray, faces: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """1 -> 4 midpoint subdivision with shared edge midpoints"""
    e = np.sort(np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]]), axis=1)
    ue, inv = np.unique(e, axis=0, return_inverse=True)
    mid = ((verts[ue[:, 0]].astype(np.float64) + verts[ue[:, 1]]) / 2).astype(np.float32)
    nf = len(faces)
    m = inv.reshape(3, nf).T + len(verts)           # midpoint vertex ids of edges (01, 12, 20)
    a, b, c = faces[:, 0], faces[:, 1], faces[:, 2]
    f2 = np.concatenate([np.stack([a, m[:, 0], m[:, 2]], 1), np.stack([m[:, 0], b, m[:, 1]], 1),
                         np.stack([m[:, 2], m[:, 1], c], 1), np.stack([m[:, 0], m[:, 1], m[:, 2]], 1)])
    return np.concatenate([verts, mid]), f2.astype(np.int32)


def vertex_normals(verts: np.ndarray, faces: np.ndarray) -> np.ndarray:
    v = verts.astype(np.float64)
    n = np.cross(v[faces[:, 1]] - v[faces[:, 0]], v[faces[:, 2]] - v[faces[:, 0]])
    acc = np.zeros_like(v)
    for k in range(3):
        np.add.at(acc, faces[:, k], n)
    l = np.linalg.norm(acc, axis=1, keepdims=True)
    return (acc / np.maximum(l, 1e-30)).astype(np.float32)


def dense_dragon(fixture_npz: str, levels: int = 2, scale: float = 1.0, translate=(0.0, 0.0, 0.0)) -> np.ndarray:
    """BASELINE config 4: the 100 000-face scenes/dragon.obj (tests/golden/dragon_100k.npz) subdivided `levels` times
    (2 -> 1.6 M triangles) with area-weighted vertex normals. Returns (n,18) float32 in OBJ face order."""
    z = np.load(fixture_npz)
    v, f = z['verts'], z['faces']
    for _ in range(levels):
        v, f = subdivide(v, f)
    v = (v * np.float32(scale) + np.asarray(translate, np.float32)).astype(np.float32)
    n = vertex_normals(v, f)
    return np.concatenate([v[f].reshape(-1, 9), n[f].reshape(-1, 9)], 1).astype(np.float32)
