"""Builds and runs scripts/experiments/sorted_traversal.cpp (research code): node visits / triangle tests per
ray of the product's traversal order against distance-sorted variants, and the share of light-sample shadow
rays whose visibility is never used, on the bench scene's ray population (primary + 2 bounces + shadow rays).
CPU only. Results quoted in DESIGN.md §2."""
import ctypes as C
import os
import subprocess
import sys
import tempfile
import warnings

warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "scripts")):
    sys.path.insert(0, p)
CSRC = os.path.join(ROOT, "chameleonrt_b200", "csrc")
LIB = os.path.join(tempfile.gettempdir(), "libcrt_sorted_traversal_exp.so")
subprocess.check_call(["make", "-s", "-C", CSRC, "host_scene.o", "bvh8_build.o"])
subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-pthread", "-march=x86-64-v3", "-ffp-contract=off", "-I" + CSRC,
                       "-I" + os.path.join(ROOT, "include"), "-shared", "-o", LIB,
                       os.path.join(ROOT, "scripts", "experiments", "sorted_traversal.cpp"),
                       os.path.join(CSRC, "host_scene.o"), os.path.join(CSRC, "bvh8_build.o")])
import numpy as np
import helpers
real = C.CDLL
helpers.C.CDLL = lambda path,*a,**k: real(LIB if "hostcheck" in path else path,*a,**k)
from chameleonrt_b200 import scenes
from oracle import OracleBackend
from oracle.oracle import primary_rays
from bvh_quality import cosine_bounce, shadow_rays
import os
_sc = os.environ.get("EXP_SCENE", "sponza_like")
scene, cam = scenes.san_miguel_like(spp=1, scale=0.1, tex_size=64) if _sc == "san_miguel_like" else getattr(scenes, _sc)(spp=1)
c = helpers.camera_for(cam)
hc = helpers.HostCheck(scene)
lib = hc.lib
lib.crt_hostcheck_trace_exp.argtypes=[C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
o = OracleBackend(fast=True); o.initialize(8,8); o.set_scene(scene)
rng = np.random.default_rng(7)
rays = primary_rays(480, 270, c.eye(), c.dir(), c.up(), cam["fov_y"])
gens=[]; pts=[]
for g in range(3):
    hits, normals = o.trace_closest(rays, True); gens.append(rays); rays, p = cosine_bounce(rays, hits, normals, rng); pts.append(p)
sh = shadow_rays(np.concatenate(pts), scene.lights[0], rng)
allr = np.concatenate(gens)
def exp(r, any_hit, mode):
    hits = np.zeros((len(r),4),np.float32); cnt = np.zeros((len(r),3),np.uint32)
    lib.crt_hostcheck_trace_exp(hc.h, r.ctypes.data, len(r), any_hit, mode, hits.ctypes.data, cnt.ctypes.data); return hits, cnt
hb,_,cb = hc.trace(allr, counters=True)
print("product closest: %.2f nodes %.2f tris" % (cb[:,0].mean(), cb[:,1].mean()))
for mode in (1,3,0,2):
    h, c_ = exp(allr, 0, mode)
    print("mode", mode, "closest: %.2f nodes %.2f tris maxstack %d  identical hits: %s" % (c_[:,0].mean(), c_[:,1].mean(), c_[:,2].max(), np.array_equal(h.view(np.uint32), hb.view(np.uint32))))
h4, c4 = exp(allr, 0, 4)
print("sorted traversal: %.2f of %.2f node visits per ray find no child at all (box of the node hit, all 8 children missed)" % (c4[:,2].mean(), c4[:,0].mean()))
ha,_,ca = hc.trace(sh, any_hit=True, counters=True)
print("product any: %.2f nodes %.2f tris" % (ca[:,0].mean(), ca[:,1].mean()))
for mode in (1,3,0,2):
    h, c_ = exp(sh, 1, mode)
    occ = (h[:,3].view(np.uint32)!=0xFFFFFFFF); occp = (ha[:,3].view(np.uint32)!=0xFFFFFFFF)
    print("mode", mode, "any: %.2f nodes %.2f tris  same occlusion: %s" % (c_[:,0].mean(), c_[:,1].mean(), np.array_equal(occ, occp)))
# fraction of light-sample shadow rays whose visibility is not used (bsdf_pdf == 0 or light_pdf < eps)
P = np.concatenate(pts)
# recompute normals for those points: regenerate generations
rng2 = np.random.default_rng(7)
rays = primary_rays(480, 270, c.eye(), c.dir(), c.up(), cam["fov_y"])
Ns=[]; 
for g in range(3):
    hits, normals = o.trace_closest(rays, True)
    hit = hits[:,3].view(np.uint32)!=0xFFFFFFFF
    n = normals[hit]; d = rays[hit][:,4:7]
    n = np.where((np.sum(n*d,axis=1)>0)[:,None], -n, n)
    Ns.append(n)
    rays, p = cosine_bounce(rays, hits, normals, rng2)
N = np.concatenate(Ns)
L = sh[:,4:7]
light = scene.lights[0]
ln = np.array(light.normal, np.float64)[:3]
front = np.sum(N*L,axis=1) > 0
lfacing = np.sum(-L*ln,axis=1) >= 1e-4
used = front & lfacing
print("light-sample shadow rays: %d; surface faces light %.3f; light faces surface %.3f; visibility used %.3f" % (len(sh), front.mean(), lfacing.mean(), used.mean()))
ha,_,ca = hc.trace(sh, any_hit=True, counters=True)
print("node visits in unused rays: %.3f of all any-hit node visits" % (ca[~used,0].sum()/ca[:,0].sum()))
occ = ha[:, 3].view(np.uint32) != 0xFFFFFFFF
print("any-hit: %.1f %% of shadow rays occluded; nodes per occluded ray %.2f, per unoccluded ray %.2f; share of any-hit node visits spent on occluded rays %.2f"
      % (100 * occ.mean(), ca[occ, 0].mean(), ca[~occ, 0].mean(), ca[occ, 0].sum() / ca[:, 0].sum()))
for mode, name in ((10, "farthest child first"), (11, "largest child first"), (12, "longest segment first"), (13, "segment x area first")):
    h, c_ = exp(sh, 1, mode)
    o2 = (h[:, 3].view(np.uint32) != 0xFFFFFFFF)
    print("any-hit order: %-24s %.2f nodes %.2f tris  same occlusion: %s" % (name, c_[:, 0].mean(), c_[:, 1].mean(), np.array_equal(o2, occ)))
