"""development helper (CPU only): rounds, empty-leaf rounds, triangle tests and node steps per ray of the lean kernels' per-lane loop
(kdpt_lean.cuh compiled for the host with -DKDPT_LEAN_COUNT, tests/hostsim) on a golden fixture's rays, for the 16-byte-node and the
two-level form:  python tools/lean_counts.py cornell_dragon_5.npz   (numbers quoted in profiles/r2_lean_variants.txt, calls 19/20)"""
import sys, os, ctypes, subprocess, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import kdtreepathtraceroptimization_b200 as kdpt
from conftest import P
CS='/root/repo/kdtreepathtraceroptimization_b200/csrc'
subprocess.check_call(f"g++ -std=c++17 -O2 -fPIC -ffp-contract=off -x c++ -pthread -shared -DKDPT_LEAN_COUNT -DHOSTSIM_COUNT -o /tmp/hs_cnt.so /root/repo/tests/hostsim/hostsim.cpp {CS}/kdpt_flatten.cpp", shell=True)
hs=ctypes.CDLL('/tmp/hs_cnt.so')
g=np.load('/root/repo/tests/golden/'+sys.argv[1])
rays=g['rays']
for mode in (200,202):
    out=np.zeros(len(rays),kdpt.HIT); ctr=np.zeros(5,'u8'); lb=np.zeros((len(rays),3),'f4')
    c=(ctypes.c_ulonglong*4)()
    hs.hostsim_lean_counts(c,1)
    rc=hs.hostsim_trace(P(g['nodes']),len(g['nodes']),P(g['tris']),len(g['tris']),6,P(rays),None,len(rays),P(out),mode,P(ctr),P(lb))
    hs.hostsim_lean_counts(c,0)
    n=len(rays)
    print(mode, rc, "rays",n,"rounds/ray %.2f empty %.2f tris %.2f steps %.2f"%(c[0]/n,c[1]/n,c[2]/n,c[3]/n))
