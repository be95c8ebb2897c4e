"""Scene-load throughput: the reference's Scene constructor (oracle/_ref/libcrt_refscene.so: tinyobjloader /
tinygltf / the .crts reader, compiled from /root/reference) against the native loader (include/crt_scene_io.h) on the same files,
page cache warm, best of N. Writes a markdown table to stdout. Host-only; run in the build container:

    python scripts/bench_scene_load.py [--scale 1.0] [--repeats 3]
"""
import argparse
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def burn(n):
    x = 0
    for i in range(n):
        x += i * i
    return x


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0, help="size of the generated scenes (1.0: ~0.7-2.8 M triangles)")
    ap.add_argument("--repeats", type=int, default=3)
    args = ap.parse_args()
    import test_scene_io as t
    from chameleonrt_b200 import scene_io
    from chameleonrt_b200.crts_io import write_crts
    from chameleonrt_b200.gltf_io import write_gltf
    from chameleonrt_b200.obj_io import write_obj
    from chameleonrt_b200.scenes import rungholt_like, san_miguel_like, sponza_like

    d = tempfile.mkdtemp(prefix="crt_load_")
    files = []
    city = rungholt_like(scale=0.12 * args.scale)[0]
    files.append(("OBJ, voxel city", write_obj(city, os.path.join(d, "city.obj"))))
    files.append(("OBJ, textured atrium", write_obj(sponza_like(detail=2.0 * args.scale, tex_size=512)[0], os.path.join(d, "sponza.obj"))))
    miguel = san_miguel_like(spp=1, scale=0.6 * args.scale, tex_size=512)[0]
    files.append(("glTF + .bin + PNG, instanced courtyard", write_gltf(miguel, os.path.join(d, "miguel.gltf"))))
    files.append((".crts, instanced courtyard", write_crts(miguel, os.path.join(d, "miguel.crts"))))
    files.append((".crts, voxel city", write_crts(city, os.path.join(d, "city.crts"))))
    import multiprocessing as mp

    t0 = time.perf_counter()
    burn(3_000_000)
    one = time.perf_counter() - t0
    t0 = time.perf_counter()
    with mp.Pool(4) as pool:
        pool.map(burn, [3_000_000] * 4)
    print(f"cores this container really gives: 4 busy processes took {(time.perf_counter() - t0) / one:.1f}x the time of one\n")
    print(f"| scene | file | unique triangles | reference loader | native loader ({os.cpu_count()} threads) | native, 1 thread | speed-up |")
    print("|---|---|---|---|---|---|---|")
    for label, path in files:
        stem, ext = os.path.splitext(os.path.basename(path))
        companions = {".obj": (".obj", ".mtl", ".png"), ".gltf": (".gltf", ".bin", ".png"), ".crts": (".crts",)}[ext]
        size = sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d) if f.startswith(stem) and f.endswith(companions))
        ref_s = min(t._reference_arrays(path)["seconds"] for _ in range(args.repeats))
        best = {}
        for threads in (0, 1):
            times = []
            for _ in range(args.repeats):
                t0 = time.perf_counter()
                loaded = scene_io.load_scene(path, threads)
                times.append(time.perf_counter() - t0)
                s = loaded.c_scene.contents
                tris = sum(s.meshes[m].geometries[g].num_tris for m in range(s.num_meshes) for g in range(s.meshes[m].num_geometries))
                del loaded
            best[threads] = min(times)
        print(f"| {label} | {size / 1e6:.0f} MB | {tris:,} | {ref_s * 1e3:.0f} ms | {best[0] * 1e3:.1f} ms | {best[1] * 1e3:.1f} ms | {ref_s / best[0]:.0f}x |")


if __name__ == "__main__":
    main()
