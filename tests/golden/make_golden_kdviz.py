"""Golden vectors for the KD-viz mode's node-box test, generated from the UNMODIFIED reference (oracle/_ref/libkdref.so:
boxIntersectionTestBox, intersections.h:51-104, executed on the host). Runs only where /root/reference was compiled
(the authoring container); the output tests/golden/kdviz_boxes.npz is committed.

    python tests/golden/make_golden_kdviz.py
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def main():
    R = ctypes.CDLL(os.path.join(ROOT, "oracle/_ref/libkdref.so"))
    rng = np.random.default_rng(2024)
    mins, maxs, rays = [], [], []
    # node boxes of the committed trees (what the mode draws) ...
    for f in ("cornell_dragon_5.npz", "cornell8_stanford_bunny.npz", "cornell_sphere_low_1.npz"):
        g = np.load(os.path.join(HERE, f))
        n = len(g['nodes'])
        idx = rng.integers(0, n, 1500)
        lo, hi = g['nodes']['mins'][idx].astype('f8'), g['nodes']['maxs'][idx].astype('f8')
        root_lo, root_hi = g['nodes']['mins'][0].astype('f8'), g['nodes']['maxs'][0].astype('f8')
        k = len(idx)
        # origins around and inside the root box, half of the rays aimed at a point of their box
        o = root_lo - 2 + rng.random((k, 3)) * (root_hi - root_lo + 4)
        inside = rng.random(k) < 0.2
        o[inside] = (lo + rng.random((k, 3)) * (hi - lo))[inside]
        d = rng.normal(size=(k, 3))
        aim = rng.random(k) < 0.6
        tgt = lo + rng.random((k, 3)) * (hi - lo)
        d[aim] = (tgt - o)[aim]
        d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-30)
        d *= rng.uniform(0.5, 2.0, (k, 1))                                       # directions are not normalised in general
        mins.append(lo); maxs.append(hi); rays.append(np.concatenate([o, d], 1))
    # ... and degenerate ones: flat boxes (a zero extent: singular matrix), axis-parallel rays, origin on a face
    lo = np.array([[0, 0, 0], [0, 0, 0], [-1, -1, -1], [-1, -1, -1], [-1, -1, -1], [2, 3, 4]], 'f8')
    hi = np.array([[1, 1, 0], [0, 1, 1], [1, 1, 1], [1, 1, 1], [1, 1, 1], [2.5, 3.5, 4.5]], 'f8')
    r = np.array([[0.5, 0.5, -3, 0, 0, 1], [-3, 0.5, 0.5, 1, 0, 0], [0, 0, -5, 0, 0, 1], [1, 0, 0, 1, 0, 0], [0, 0, 0, 0, 1, 0],
                  [0, 0, 0, 2.25, 3.25, 4.25]], 'f8')
    mins.append(lo); maxs.append(hi); rays.append(r)
    mins = np.ascontiguousarray(np.concatenate(mins), 'f4'); maxs = np.ascontiguousarray(np.concatenate(maxs), 'f4')
    rays = np.ascontiguousarray(np.concatenate(rays), 'f4')
    n = len(rays)
    t = np.zeros(n, 'f4'); pt = np.zeros((n, 3), 'f4'); nr = np.zeros((n, 3), 'f4'); outside = np.zeros(n, 'u1')
    R.kdref_box_test_box(P(mins), P(maxs), P(rays), n, P(t), P(pt), P(nr), P(outside))
    out = os.path.join(HERE, "kdviz_boxes.npz")
    np.savez_compressed(out, mins=mins, maxs=maxs, rays=rays, t=t, point=pt, normal=nr, outside=outside)
    print(out, n, "pairs,", int((t > 0).sum()), "hits,", int(np.isnan(t).sum()), "NaN,", os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    main()
