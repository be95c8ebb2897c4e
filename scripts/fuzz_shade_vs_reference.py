"""Randomised comparison of the product's DEVICE shading source compiled for the host
(libcrt_shade_hostcheck.so = chameleonrt_b200/csrc/shade_math.cuh) with the reference's own functions
(oracle/_ref/libcrt_embree.so = /root/reference/backends/embree/disney_bsdf.ih, lights.ih): random Disney
materials over the whole parameter cube, random normals and directions in BOTH hemispheres (grazing ones
included), random rng states. Expected: pdf, sampled direction and rng state bit-identical; BSDF value identical up
to the multiplied-out Schlick weight. CPU only.   python scripts/fuzz_shade_vs_reference.py [n] [seed]"""
import ctypes as C
import os
import sys
import warnings

import numpy as np

warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_embree import load_ref_embree_lib  # noqa: E402


def main(n, seed):
    ref = load_ref_embree_lib()
    dev = C.CDLL(os.path.join(ROOT, "chameleonrt_b200", "csrc", "libcrt_shade_hostcheck.so"))
    vp = C.c_void_p
    dev.shadekat_disney_eval.argtypes = [vp] * 5
    dev.shadekat_disney_sample.argtypes = [vp] * 5
    rng = np.random.default_rng(seed)
    worst_f = 0.0
    n_f_diff = n_pdf_diff = n_dir_diff = 0
    for _ in range(n):
        m = rng.random(16).astype(np.float32)
        m[12] = np.float32(rng.uniform(1.0, 2.5))                     # ior
        if rng.random() < 0.5:
            m[13] = 0.0                                                 # opaque
        if rng.random() < 0.5:
            m[7] = 0.0                                                  # isotropic
        for k in (3, 5, 10, 11):
            if rng.random() < 0.1:
                m[k] = np.float32(rng.choice([0.0, 1.0]))              # parameter extremes
        nrm = rng.normal(size=3)
        nrm = (nrm / np.linalg.norm(nrm)).astype(np.float32)
        wo = rng.normal(size=3)
        wi = rng.normal(size=3)
        if rng.random() < 0.2:                                          # grazing
            wo = wo - np.dot(wo, nrm) * nrm * 0.999
        wo = (wo / np.linalg.norm(wo)).astype(np.float32)
        wi = (wi / np.linalg.norm(wi)).astype(np.float32)
        a, b = np.zeros(4, np.float32), np.zeros(4, np.float32)
        ref.refispc_kat_disney_eval(m.ctypes.data, nrm.ctypes.data, wo.ctypes.data, wi.ctypes.data, a.ctypes.data)
        dev.shadekat_disney_eval(m.ctypes.data, nrm.ctypes.data, wo.ctypes.data, wi.ctypes.data, b.ctypes.data)
        if a.view(np.uint32)[3] != b.view(np.uint32)[3] and not (np.isnan(a[3]) and np.isnan(b[3])):
            n_pdf_diff += 1
        for k in range(3):
            if np.isnan(a[k]) or np.isnan(b[k]) or np.isinf(a[k]) or np.isinf(b[k]):
                if not ((np.isnan(a[k]) and np.isnan(b[k])) or a[k] == b[k]):
                    n_f_diff += 1
                    worst_f = np.inf
                continue
            if a[k] != b[k]:
                worst_f = max(worst_f, abs(float(a[k]) - float(b[k])) / max(abs(float(a[k])), 1e-30))
        sa, sb = np.zeros(8, np.float32), np.zeros(8, np.float32)
        st = int(rng.integers(1, 2 ** 32 - 1))
        ra, rb = C.c_uint32(st), C.c_uint32(st)
        ref.refispc_kat_disney_sample(m.ctypes.data, nrm.ctypes.data, wo.ctypes.data, C.addressof(ra), sa.ctypes.data)
        dev.shadekat_disney_sample(m.ctypes.data, nrm.ctypes.data, wo.ctypes.data, C.addressof(rb), sb.ctypes.data)
        same_dir = np.array_equal(sa[3:7].view(np.uint32), sb[3:7].view(np.uint32)) or bool((np.isnan(sa[3:7]) == np.isnan(sb[3:7])).all() and
                                                                                          np.allclose(np.nan_to_num(sa[3:7]), np.nan_to_num(sb[3:7]), rtol=0, atol=0))
        if not same_dir or ra.value != rb.value:
            n_dir_diff += 1
    print(f"{n} random (material, n, w_o, w_i, rng) tuples: pdf differs in {n_pdf_diff}, sampled (pdf, w_i, rng state) in {n_dir_diff}, "
          f"non-finite mismatch in {n_f_diff}; worst relative difference of the BSDF value {worst_f:.3e}")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 20000, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
