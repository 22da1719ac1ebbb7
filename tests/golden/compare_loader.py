"""Run by tests/test_scene_loader_cpu.py (one process per OBJ: the reference keeps its scene in globals): loads SCENE.txt +
MESH.obj through the UNMODIFIED reference loader (oracle/_ref/libkdref.so) and through the product's loader and compares
every array byte for byte. Only usable where /root/reference exists. Prints one line: "<obj> ... ALL EQUAL" on success."""
import ctypes, os, sys, numpy as np
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0,ROOT)
import kdtreepathtraceroptimization_b200 as kdpt
txt, obj = sys.argv[1], sys.argv[2]
R=ctypes.CDLL(os.path.join(ROOT,'oracle/_ref/libkdref.so'))
P=lambda a:a.ctypes.data_as(ctypes.c_void_p)
devnull=os.open(os.devnull,os.O_WRONLY); saved=os.dup(1); os.dup2(devnull,1)
R.kdref_load(txt.encode(), obj.encode(), (os.path.dirname(obj)+'/').encode())
os.dup2(saved,1)
s=(ctypes.c_int*9)(); R.kdref_sizes(s)
nn,nt,ng,nm,nshape,depth,iters,rx,ry=list(s)
nodes=np.zeros(nn,kdpt.NODE); tris=np.zeros(nt,kdpt.TRI); geoms=np.zeros(ng,kdpt.GEOM); mats=np.zeros(nm,kdpt.MATERIAL)
offs=np.zeros(max(nshape,1),'i4'); cam=np.zeros(1,kdpt.CAMERA)
R.kdref_copy(P(nodes),P(tris),P(geoms),P(mats),P(offs),P(cam))
offs=offs[:nshape]
nscene=int(offs[0]) if nshape>0 else nm
mats['transmittance'][:nscene]=0
try:
    a=kdpt.Scene.load(txt,obj)
except Exception as e:
    print(os.path.basename(obj),'OURS FAILED:',e,'| ref: nodes',nn,'tris',nt,'shapes',nshape); sys.exit(0)
res={k:(getattr(a,k).tobytes()==v.tobytes()) for k,v in (('nodes',nodes),('tris',tris),('geoms',geoms),('materials',mats),('offsets',offs),('camera',cam))}
print(os.path.basename(obj),'nodes',nn,'tris',nt,'shapes',nshape,'mats',nm,res if not all(res.values()) else 'ALL EQUAL')
