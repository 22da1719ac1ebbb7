"""CPU-only: the product's host logic (flatten, C-ABI surface) and — through tests/hostsim, a g++ build of the very
same per-ray device functions the kernels call — the traversal logic against the oracle. No GPU compute here."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import kdtreepathtraceroptimization_b200 as kdpt
from conftest import (KO_BRUTE_T1, KO_BRUTE_T2, KO_STACK_T1, KO_T1, KO_T2, P, ROOT, assert_hits_equal, bits)


def _sim(hostsim, g, mode, tmin=None, nmat=6, rays=None, refine=-1):
    """refine: leaf size of the exact modes' refined tree (-1 = the product's default, 0 = walk the reference tree)"""
    hostsim.hostsim_set_refine(refine)
    rays = g['rays'] if rays is None else rays
    n = len(rays)
    out = np.zeros(n, kdpt.HIT)
    lb = np.zeros((n, 3), 'f4')
    ctr = np.zeros(5, 'u8')
    rc = hostsim.hostsim_trace(P(g['nodes']), len(g['nodes']), P(g['tris']), len(g['tris']), nmat, P(rays),
                               P(tmin) if tmin is not None else None, n, P(out), mode, P(ctr), P(lb))
    assert rc == 0
    return out, lb, ctr


@pytest.mark.parametrize("mode,omode", [(kdpt.FULLSTACK, KO_BRUTE_T1), (kdpt.SHORTSTACK, KO_BRUTE_T2), (100, KO_BRUTE_T2)])
def test_exact_modes_equal_brute_force(hostsim, oracle, golden, mode, omode):
    for tm in (None, golden['t_min_in']):
        o = oracle.trace(omode, golden['nodes'], golden['tris'], golden['rays'], tm, 6)
        for refine in (-1, 0, 1, 16):       # default refinement, the reference tree itself, finest, coarse
            k, _, _ = _sim(hostsim, golden, mode, tm, refine=refine)
            assert_hits_equal(k, o, golden['tris'], f"mode {mode} refine {refine}")


LEAN_FULL, LEAN_RESTART, LEAN_WIDE_FULL, LEAN_WIDE_RESTART = 200, 201, 202, 203


@pytest.mark.parametrize("mode,omode", [(LEAN_FULL, KO_BRUTE_T1), (LEAN_RESTART, KO_BRUTE_T2), (LEAN_WIDE_FULL, KO_BRUTE_T1),
                                        (LEAN_WIDE_RESTART, KO_BRUTE_T2)])
def test_lean_loop_equals_brute_force(hostsim, oracle, golden, mode, omode):
    """the per-lane loop of the lean kernels (kdpt_lean.cuh: leanBegin / leanRound, compiled for the host) on the refined tree, on its
    16-byte nodes and on the two-level records, full stack and 8-entry ring with kd-restart"""
    for tm in (None, golden['t_min_in']):
        o = oracle.trace(omode, golden['nodes'], golden['tris'], golden['rays'], tm, 6)
        for refine in (-1, 1, 3, 16):
            k, _, ctr = _sim(hostsim, golden, mode, tm, refine=refine)
            assert_hits_equal(k, o, golden['tris'], f"lean mode {mode} refine {refine}")
    # the two node forms visit the same leaves in the same order: the same number of rounds (full stack)
    if mode == LEAN_WIDE_FULL:
        _, _, a = _sim(hostsim, golden, LEAN_FULL)
        _, _, b = _sim(hostsim, golden, LEAN_WIDE_FULL)
        assert int(a[1]) == int(b[1]) > 0


@pytest.mark.parametrize("mode,omode,name", [(kdpt.REF_T1, KO_T1, "T1"), (kdpt.REF_T2, KO_T2, "T2")])
def test_ref_modes_equal_reference(hostsim, oracle, golden, mode, omode, name):
    for sfx, tm in (("", None), ("_tmin", golden['t_min_in'])):
        k, lb, _ = _sim(hostsim, golden, mode, tm)
        o = oracle.trace(omode, golden['nodes'], golden['tris'], golden['rays'], tm, 6)
        assert_hits_equal(k, o, golden['tris'], name)
        # and directly against the reference's own outputs
        assert ((k['tri'] >= 0) == (golden[f'ref_{name}{sfx}_hit'] > 0)).all()
        assert (bits(k['t']) == bits(golden[f'ref_{name}{sfx}_tmin'])).all()
        # bary.z (the value the reference prunes on) — x/y of rejected triangles are not reproduced (never read)
        assert (bits(lb[:, 2]) == bits(golden[f'ref_{name}{sfx}_bary'][:, 2])).all()


def test_counters_match_oracle_stack_traversal(hostsim, oracle, golden):
    """the algorithmic-bytes bookkeeping (V nodes, T triangle tests per ray) is the same on both sides"""
    _, _, ctr = _sim(hostsim, golden, kdpt.FULLSTACK, refine=0)
    _, oc = oracle.trace(KO_STACK_T1, golden['nodes'], golden['tris'], golden['rays'], None, 6, counters=True)
    assert int(ctr[0]) == int(oc['nodes']) and int(ctr[1]) == int(oc['leaves'])
    # triangles actually tested: never more than the algorithmic count, fewer where chunk boxes cull fat leaves
    assert int(ctr[2]) <= int(oc['tris'])
    # the refined tree trades node steps for triangle tests
    _, _, cx = _sim(hostsim, golden, kdpt.FULLSTACK)
    assert int(cx[0]) >= int(ctr[0]) and int(cx[2]) <= int(ctr[2])


def test_refined_tree_shape(hostsim, golden_all):
    """every fat reference leaf is split until the surface-area heuristic stops paying; depth stays within the per-ray
    stacks (kMaxDepth = 32)"""
    for name, g in golden_all.items():
        info = np.zeros(4, 'i8')
        assert hostsim.hostsim_refined_info(P(g['nodes']), len(g['nodes']), P(g['tris']), len(g['tris']), -1, P(info)) == 0
        nodes, entries, depth, leaf = (int(x) for x in info)
        assert nodes >= len(g['nodes']) and depth <= 31
        assert leaf <= max(24, int(g['nodes']['triIdSize'].max()) // 8), (name, leaf)
        assert entries <= 8 * len(g['tris'])


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_lean_loop_on_random_soups(hostsim, oracle, seed):
    """clustered random triangles with exact duplicates, zero-area and axis-aligned ones (deep refinement, empty cells, leaves the
    refinement cannot split): the lean loop in all four forms and the packet form vs brute force, rays from inside and outside"""
    rng = np.random.default_rng(seed)
    n = 2500
    centres = rng.uniform(-1, 1, (12, 3))
    c = centres[rng.integers(0, 12, n)] + rng.normal(scale=0.08, size=(n, 3))
    e1 = rng.normal(scale=0.03, size=(n, 3)); e2 = rng.normal(scale=0.03, size=(n, 3))
    e2[:40] = e1[:40] * 2                                   # zero area
    e1[40:80, 2] = 0; e2[40:80, 2] = 0                      # axis-aligned (flat boxes before padding)
    v = np.concatenate([c, c + e1, c + e2], 1).astype('f4')
    v[100:160] = v[200:260]                                 # exact duplicates (canonical index = the smaller one)
    v[300:500] = v[300]                                     # 200 copies of one triangle: a leaf no plane can split
    nr = rng.normal(size=(n, 9)).astype('f4')
    nr[100:160] = nr[200:260]; nr[300:500] = nr[300]
    nodes, tris, _ = oracle.build_tree(np.concatenate([v, nr], 1), None, 13)
    m = 3000
    d = rng.normal(size=(m, 3))
    o = np.where(rng.random((m, 1)) < 0.5, rng.uniform(-1.2, 1.2, (m, 3)), 3 * d / np.linalg.norm(d, axis=1, keepdims=True))
    tgt = centres[rng.integers(0, 12, m)] + rng.normal(scale=0.1, size=(m, 3))
    d = np.where(rng.random((m, 1)) < 0.7, tgt - o, d)
    rays = np.concatenate([o, d], 1).astype('f4')
    g = dict(nodes=nodes, tris=tris, rays=rays)
    ref = {om: oracle.trace(om, nodes, tris, rays, None, 1) for om in (KO_BRUTE_T1, KO_BRUTE_T2)}
    assert (ref[KO_BRUTE_T1]['hit'] > 0).sum() > 300
    for mode, om in ((kdpt.FULLSTACK, KO_BRUTE_T1), (kdpt.SHORTSTACK, KO_BRUTE_T2), (LEAN_FULL, KO_BRUTE_T1), (LEAN_RESTART, KO_BRUTE_T2),
                     (LEAN_WIDE_FULL, KO_BRUTE_T1), (LEAN_WIDE_RESTART, KO_BRUTE_T2)):
        for refine in (-1, 2):
            k, _, _ = _sim(hostsim, g, mode, nmat=1, refine=refine)
            assert_hits_equal(k, ref[om], tris, f"seed {seed} mode {mode} refine {refine}")


def test_two_level_records_encode_the_same_tree(hostsim, golden_all):
    """kdpt_flatten.cpp:buildWide against the 16-byte nodes it was made from: walking the records reaches every leaf of the refined
    tree exactly once, with the same planes and axes on the way and the same leaf ranges (inline or by record index)"""
    hostsim.hostsim_wide_check.restype = ctypes.c_long
    for name, g in golden_all.items():
        for leaf in (-1, 1, 16):
            info = np.zeros(3, 'i8')
            got = hostsim.hostsim_wide_check(P(g['nodes']), len(g['nodes']), P(g['tris']), len(g['tris']), leaf, P(info))
            nodes, records, leaves = (int(x) for x in info)
            assert got == leaves > 0, (name, leaf, got, leaves)
            assert records <= (nodes - leaves + 1) // 2 + 1 + leaves // 2, (name, nodes, records)     # about one record per two interior nodes


def test_degenerate_rays(hostsim, oracle, golden_all):
    g = golden_all["cornell8_stanford_bunny.npz"]
    lo, hi = g['nodes']['mins'][0], g['nodes']['maxs'][0]
    c = (lo + hi) / 2
    rays = np.array([[*c, 0, 0, 0], [*c, np.nan, 0, 1], [*c, 1e-30, 0, 0], [*c, 0, 1e30, 0], [*lo, 1, 1, 1], [*hi, -1, -1, -1],
                     [c[0], c[1], lo[2] - 5, 0, 0, 1], [c[0], c[1], lo[2] - 5, 0, 0, -1], [*c, 0, -0.0, 1]], 'f4')
    for mode, om in ((kdpt.FULLSTACK, KO_BRUTE_T1), (kdpt.SHORTSTACK, KO_BRUTE_T2), (kdpt.REF_T1, KO_T1), (kdpt.REF_T2, KO_T2),
                     (LEAN_FULL, KO_BRUTE_T1), (LEAN_RESTART, KO_BRUTE_T2), (LEAN_WIDE_FULL, KO_BRUTE_T1), (LEAN_WIDE_RESTART, KO_BRUTE_T2)):
        k, _, _ = _sim(hostsim, g, mode, rays=rays)
        o = oracle.trace(om, g['nodes'], g['tris'], rays, None, 6)
        assert_hits_equal(k, o, g['tris'], f"degenerate mode {mode}")


def test_tiny_trees(hostsim, oracle):
    """0, 1, 2 and 3 triangles: single-leaf trees (KDnode.cpp:164) and the T2 tag quirk on a root leaf"""
    rng = np.random.default_rng(5)
    for n in (1, 2, 3, 7):
        v = rng.uniform(-1, 1, (n, 9)).astype('f4')
        nr = rng.normal(size=(n, 9)).astype('f4')
        nodes, tris, _ = oracle.build_tree(np.concatenate([v, nr], 1), None, 13)
        o = rng.uniform(-2, 2, (500, 3)); d = rng.normal(size=(500, 3))
        rays = np.concatenate([o, d], 1).astype('f4')
        g = dict(nodes=nodes, tris=tris, rays=rays)
        for mode, om in ((kdpt.FULLSTACK, KO_BRUTE_T1), (kdpt.SHORTSTACK, KO_BRUTE_T2), (kdpt.REF_T1, KO_T1), (kdpt.REF_T2, KO_T2),
                         (LEAN_FULL, KO_BRUTE_T1), (LEAN_RESTART, KO_BRUTE_T2), (LEAN_WIDE_FULL, KO_BRUTE_T1), (LEAN_WIDE_RESTART, KO_BRUTE_T2)):
            # (the lean loop on a single-leaf tree: the root itself is a leaf reference in the two-level form)
            for refine in ((-1, 1) if mode >= 200 else (-1,)):
                k, _, _ = _sim(hostsim, g, mode, nmat=1, refine=refine)
                assert_hits_equal(k, oracle.trace(om, nodes, tris, rays, None, 1), tris, f"n={n} mode {mode} refine {refine}")


def test_flatten_rejects_malformed_trees(hostsim, golden_all):
    g = golden_all["cornell_sphere_low_1.npz"]
    rays = g['rays'][:4]
    out = np.zeros(4, kdpt.HIT)

    def run(nodes, tris):
        return hostsim.hostsim_trace(P(nodes), len(nodes), P(tris), len(tris), 6, P(rays), None, 4, P(out), 0, None, None)
    assert run(g['nodes'], g['tris']) == 0
    bad = g['nodes'].copy(); bad['ID'][3] = 77
    assert run(bad, g['tris']) == -1
    bad = g['nodes'].copy(); bad['leftID'][0] = 10 ** 6
    assert run(bad, g['tris']) == -1
    bad = g['nodes'].copy()
    leaf = np.nonzero(bad['triIdSize'] > 0)[0][0]; bad['triIdStart'][leaf] = len(g['tris'])
    assert run(bad, g['tris']) == -1
    bad = g['nodes'].copy(); bad['parentID'][0] = 3
    assert run(bad, g['tris']) == -1
    # a node listed as a child by two parents (a DAG: child-link paths longer than the parent-link depth the stacks are sized by)
    bad = g['nodes'].copy()
    inner = [i for i in range(len(bad)) if bad['leftID'][i] != -1]
    a, b = inner[1], inner[2]
    bad['leftID'][b] = bad['leftID'][a]
    assert run(bad, g['tris']) == -1
    # both children the same node
    bad = g['nodes'].copy(); bad['rightID'][a] = bad['leftID'][a]
    assert run(bad, g['tris']) == -1
    # two leaves claiming overlapping triangle ranges
    bad = g['nodes'].copy()
    leaves = np.nonzero(bad['triIdSize'] > 1)[0]
    bad['triIdStart'][leaves[1]] = bad['triIdStart'][leaves[0]] + 1
    assert run(bad, g['tris']) == -1


def test_kdviz_node_geoms_equal_the_oracles(oracle, golden_all):
    """kdpt_node_box_geom (what KDPT_KDVIZ draws for a node) builds the matrices of boxIntersectionTestBox exactly like the
    oracle, whose cube test on them is pinned to the reference's outputs (test_kdviz_node_box_test_matches_reference): transform
    and inverse byte for byte on every node of every fixture tree. Host-only, no GPU."""
    for name, g in golden_all.items():
        nodes = np.ascontiguousarray(g['nodes'])
        mine = kdpt.node_box_geoms(nodes, 5)
        ref = np.zeros(len(nodes), kdpt.GEOM)
        oracle.lib.ko_kdviz_geoms(P(nodes), len(nodes), 6, P(ref))
        for k in ('type', 'materialid', 'transform', 'inverseTransform', 'translation', 'scale'):
            assert mine[k].tobytes() == ref[k].tobytes(), (name, k)


def test_c_abi_exports_every_declared_symbol():
    """libkdpt.so loads on a machine without a GPU and exports what include/kdpt.h declares; compute calls fail
    loudly (no CPU fallback)."""
    hdr = open(os.path.join(ROOT, "include", "kdpt.h")).read()
    names = sorted(set(re.findall(r"\b(kdpt_[a-z_]+)\s*\(", hdr)))
    assert len(names) >= 12
    lib = kdpt.load_library()
    for n in names:
        assert hasattr(lib, n), n
    out = subprocess.check_output(["nm", "-D", "--defined-only", kdpt.LIB_PATH]).decode()
    for n in names:
        assert re.search(rf"\bT {n}\b", out), n
    assert lib.kdpt_abi_version() == 1


def test_no_gpu_means_loud_failure(golden_all):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(kdpt.KdptError, match="no CUDA device"):
        kdpt.Context(kdpt.Scene.from_npz(os.path.join(ROOT, "tests/golden/cornell_sphere_low_1.npz")))


def test_gpu_kd_build_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(kdpt.KdptError, match="no CUDA device"):
        kdpt.Scene.from_triangles_gpu(np.zeros((4, 18), 'f4'))


def test_canonical_index_of_a_large_tree_with_duplicate_triangles(hostsim):
    """flatten's canonical index (first TriBare entry with identical 76 bytes) on a tree big enough for its partitioned,
    multi-threaded path, with a tenth of the input triangles duplicated byte for byte under other source indices"""
    from kdtreepathtraceroptimization_b200 import meshes
    from conftest import P, canonical_index
    t18 = meshes.triangle_soup(200_000, seed=5, edge=0.02)
    t18[-20_000:] = t18[:20_000]
    sc = kdpt.Scene.from_triangles(t18)
    assert len(sc.tris) >= 1 << 16
    got = np.zeros(len(sc.tris), 'i4')
    assert hostsim.hostsim_canonical(P(sc.nodes), len(sc.nodes), P(sc.tris), len(sc.tris), P(got)) == 0
    want = canonical_index(sc.tris)
    assert (want != np.arange(len(want))).sum() > 20_000         # duplicates across leaves and across sources
    assert (got == want).all()


def test_integration_doc_matches_the_compiled_binding():
    """INTEGRATION.md's shim is the file that oracle/Makefile compiles against the reference (oracle/dropin/)"""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = doc[doc.index("```cpp") + 6:]
    block = block[:block.index("```")]
    src = open(os.path.join(ROOT, "oracle", "dropin", "pathtrace_kdpt.cpp")).read()

    def code(text):
        lines = [l.rstrip() for l in text.splitlines()]
        return [l for l in lines if l.strip() and not l.lstrip().startswith("// ")]
    assert code(block) == code(src)


def test_product_never_touches_the_oracle():
    """no file of the product (package, csrc, include) includes, links, loads or imports anything under oracle/
    (comments may cite it)"""
    pat = re.compile(r"#\s*include[^\n]*oracle|liboracle|libkdref|CDLL\([^\n]*oracle|import[^\n]*oracle|-l\s*oracle")
    for base in ("kdtreepathtraceroptimization_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            if "build" in dp or "__pycache__" in dp:
                continue
            for f in fs:
                if f.endswith((".so", ".o", ".log", ".pyc")):
                    continue
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert not pat.search(txt), os.path.join(dp, f)
    out = subprocess.check_output(["nm", "-D", kdpt.LIB_PATH]).decode()
    assert "ko_" not in out and "kdref_" not in out


def test_headless_cli_builds_and_fails_loudly_without_a_gpu():
    import torch
    exe = os.path.join(ROOT, "kdtreepathtraceroptimization_b200", "kdpt_render")
    assert os.path.exists(exe)
    assert subprocess.run([exe, "--help"], capture_output=True).returncode == 0
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    scenes = os.path.join(ROOT, "tests", "golden", "scenes")
    r = subprocess.run([exe, os.path.join(scenes, "cornell.txt"), os.path.join(scenes, "sphere_low_1.obj"), "--spp", "1"],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "no CUDA device" in r.stderr


def _check_contract(d, reference):
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "e2e", "cpu_baseline"):
        assert k in d, k
    assert d["unit"] == "Mrays/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and isinstance(d["config"].get("workload"), str)
    assert "model" not in d["config"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in d["cpu_baseline"], k
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    if reference:
        assert d["impl"] == "reference" and d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
        assert d["cpu_baseline"]["value"] == d["value"]
    else:
        assert d["gpu_launches"] > 0 and d["check"]["ok"] is True
        r = d["roofline"]
        for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
            assert k in r, k
        assert r["bound"] in ("hbm", "l2") and r["unit"] == "GB/s" and 0 < r["frac"] <= 1.0
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        c = d["clocks"]
        assert c["sm_mhz"] > 0 and c["sm_max_mhz"] >= c["sm_mhz"] and isinstance(c["reasons"], list)
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0


def test_bench_line_contract():
    """the JSON line of bench.py: the reference arm run live on its smallest workload (the reference's own host functions, no GPU, and
    no libkdpt.so in that process), the product arm from the line the last GPU call of the round recorded"""
    import json
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "sphere", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    _check_contract(json.loads(out.stdout.strip().splitlines()[-1]), reference=True)
    lines = [l for l in open(os.path.join(ROOT, "profiles", "r2_bench_lines_final.txt")) if l.startswith("{")]
    assert len(lines) >= 3
    checked = 0
    for l in lines:
        d = json.loads(l)
        if "1.6 M tris" in d["config"]["workload"]:         # the driver's commands (the 10 M soup line was taken without its CPU sample)
            _check_contract(d, reference=d.get("impl") == "reference")
            checked += 1
    assert checked >= 3
    assert "1.6 M tris" in json.loads(lines[-1])["config"]["workload"] and "1920x1080" in json.loads(lines[-1])["config"]["workload"]
