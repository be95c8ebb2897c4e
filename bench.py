#!/usr/bin/env python
"""bench.py — MRays/s and ms/frame of the per-frame render path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W          # the CUDA wavefront backend
  python bench.py --impl reference --gpus N ...           # the CPU path (oracle), rank 0 only

A "step" is one frame: one pass of the hot path (raygen -> [closest-hit traversal -> shade ->
any-hit traversal -> NEE resolve] x depth -> resolve/tonemap) over the whole image at the
configured samples per pixel. Workload = BASELINE.json configs[1]: Sponza-class OBJ scene
(seeded procedural stand-in, the asset is not on the box), 1280x720, 4 spp, max depth 8.

N > 1 (torchrun, one process per GPU): the image's 64x64 tiles are sharded round-robin across
ranks (scene + BVH replicated), every rank renders its tiles with no communication, and the
accumulated tiles are gathered to rank 0 over NCCL/NVLink at frame end, inside the timed
region. Total work is fixed, so scaling is "strong".

One JSON line on stdout (rank 0). Keys follow the driver contract; `roofline` describes
k_traverse, `cpu_baseline` the CPU oracle timed on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# BASELINE.json configs. The default (what the driver runs) is configs[1] = "c2"; c3 / c4 are the
# HBM-resident stress cases and are run by hand (results under profiles/).
WORKLOADS = {
    # configs[0] is the reference's own CPU-runnable plumbing case; here for completeness (a 34-triangle scene says
    # nothing about a GPU)
    "c1": dict(name="C1 cornell_box OBJ-class scene (34 tris), 512x512, 1 spp, max depth 5", gen="cornell_box", kw={}, w=512, h=512,
               spp=1, depth=5),
    "c2": dict(name="C2 sponza_like OBJ-class scene (263,792 tris, 25 materials, 8 sRGB 1024^2 textures), "
                    "1280x720, 4 spp, max depth 8", gen="sponza_like", kw={}, w=1280, h=720, spp=4, depth=8),
    "c3": dict(name="C3 san_miguel_like glTF-class scene (10.5 M instanced tris, 961 instances, 100 materials, "
                    "11 textures), 1920x1080, 8 spp, max depth 8", gen="san_miguel_like", kw={}, w=1920, h=1080, spp=8, depth=8),
    "c4": dict(name="C4 rungholt_like OBJ-class voxel city (6.7 M tris, 80 untextured materials), "
                    "1920x1080, 4 spp, max depth 8", gen="rungholt_like", kw={"scale": 1.24}, w=1920, h=1080, spp=4, depth=8),
    "c5": dict(name="C5 san_miguel_like glTF-class scene (10.5 M instanced tris), 3840x2160, 8 spp per step (64 spp accumulated "
                    "after 8 steps), max depth 8; meant for --gpus 8", gen="san_miguel_like", kw={}, w=3840, h=2160, spp=8, depth=8),
    # not a benchmark: a frame small enough for the CPU emulation of the renderer, used by tests/test_bench_contract.py
    # to dry-run this file's GPU arm where no GPU exists
    "dev": dict(name="DEV cornell_box, 128x48, 1 spp, max depth 5 (dry run, not a benchmark)", gen="cornell_box", kw={}, w=128, h=48,
                spp=1, depth=5),
}
WORKLOAD = WORKLOADS["c2"]["name"]
WIDTH, HEIGHT, SPP, MAX_DEPTH = 1280, 720, 4, 8
S_NODE, S_TRI, S_RAY, S_HIT = 80, 48, 32, 16  # algorithmic bytes, SURVEY.md §8d / DESIGN.md §5


def select_workload(key: str) -> None:
    global WORKLOAD, WIDTH, HEIGHT, SPP, MAX_DEPTH, _GEN
    w = WORKLOADS[key]
    WORKLOAD, WIDTH, HEIGHT, SPP, MAX_DEPTH = w["name"], w["w"], w["h"], w["spp"], w["depth"]
    _GEN = (w["gen"], w["kw"])


_GEN = ("sponza_like", {})


def make_workload():
    from chameleonrt_b200 import ArcballCamera
    from chameleonrt_b200 import scenes

    scene, cam = getattr(scenes, _GEN[0])(spp=SPP, **_GEN[1])
    camera = ArcballCamera(cam["eye"], cam["center"], cam["up"])
    return scene, (camera.eye(), camera.dir(), camera.up(), cam["fov_y"])


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi clocks and throttle reasons while the timed region runs."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self._stop = threading.Event()
        self._thread = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.QUERY}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([c.strip() for c in line.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thread.join(timeout=10)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_backend():
    """The CPU implementation of the path that is timed beside the GPU. Where the reference's own Embree
    backend was compiled (oracle/_ref/libcrt_embree_fast.so: /root/reference/backends/embree sources, ISPC
    kernels as scalar C++, Embree/TBB replaced by stand-ins; oracle/ref_build/Makefile) that is used and the
    kind is "reference"; otherwise the oracle port. Both render bit-identical frames in their strict builds
    (tests/test_reference_embree.py); these are the -O3 builds of each."""
    from oracle import ref_embree

    if ref_embree.available(fast=True) and os.environ.get("CRT_BENCH_CPU", "reference") != "port":
        return (ref_embree.RefEmbreeBackend(max_depth=MAX_DEPTH, fast=True), "reference",
                "the reference's backends/embree sources (render_embree.cpp + render_embree.ispc/.ih compiled as "
                "scalar C++, -O3 x86-64-v3; Embree replaced by an own 4-wide BVH, TBB by std::thread), all host threads")
    from oracle import OracleBackend

    return (OracleBackend(max_depth=MAX_DEPTH, fast=True), "port",
            "CPU oracle = Embree/ISPC backend restated with an own BVH2, -O3 x86-64-v3, all host threads")


def time_oracle(scene, view, budget_s: float, frames_cap: int):
    """Times the CPU implementation (all host threads) on the workload; returns per-frame list."""
    cpu, kind, desc = cpu_reference_backend()
    cpu.initialize(WIDTH, HEIGHT)
    cpu.set_scene(scene)
    results = []
    t0 = time.time()
    f = 0
    while True:
        st = cpu.render(*view, f == 0, True)
        results.append((st.render_time, st.num_rays))
        f += 1
        if f >= frames_cap or (time.time() - t0) > budget_s:
            break
    return results, kind, desc


def run_reference_arm(args):
    """--impl reference: the CPU implementation of the path on this box's host cores, with every host
    thread: the reference's own Embree backend sources where they were compiled (oracle/_ref, prebuilt for
    the GPU box), else the oracle port (cpu_reference_backend)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scene, view = make_workload()
    budget_s = float(os.environ.get("CRT_BENCH_REF_BUDGET", "240"))
    cpu, kind, desc = cpu_reference_backend()
    cpu.initialize(WIDTH, HEIGHT)
    cpu.set_scene(scene)
    # probe: one full frame. Each step is a full frame of the workload when K + W of them fit the
    # budget on this box's cores; otherwise a step keeps every pixel and takes fewer samples per pixel
    # (same rays, same scene, same depth: MRays/s is unchanged in meaning, the sample is smaller).
    t0 = time.time()
    cpu.render(*view, True, True)
    probe_s = time.time() - t0
    spp_step = SPP
    est = probe_s * (args.steps + args.warmup)
    if est > budget_s and SPP > 1:
        spp_step = max(1, int(SPP * budget_s / est))
        scene.samples_per_pixel = spp_step
        # a fresh backend: the reference's set_scene appends to its material/texture tables
        # (render_embree.cpp:106-131), it is meant to be called once per renderer
        cpu, kind, desc = cpu_reference_backend()
        cpu.initialize(WIDTH, HEIGHT)
        cpu.set_scene(scene)
    f = 0
    for _ in range(args.warmup):
        cpu.render(*view, f == 0, True)
        f += 1
    rays, ms = 0, 0.0
    for _ in range(args.steps):
        st = cpu.render(*view, f == 0, True)
        f += 1
        rays += st.num_rays
        ms += st.render_time
    value = rays / (ms * 1e3)
    cores = cpu_cores()
    line = {
        "impl": "reference", "metric": "MRays/s", "value": value, "unit": "MRays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "width": WIDTH, "height": HEIGHT, "spp": SPP, "max_depth": MAX_DEPTH},
        "cpu_baseline": {"value": value, "unit": "MRays/s", "cores": cores, "kind": kind,
                         "sample": f"{args.steps} frames of the workload at {spp_step} of {SPP} spp per step after 1 "
                                   f"probe + {args.warmup} warm-up frames; {desc}"},
        "e2e": {"value": value, "unit": "MRays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_set_scene_probe() -> None:
    """Sub-process of the N=1 run (`--probe-set-scene`): crtc_set_scene of the bench scene with the BVH8 built on the
    host (binned SAH, the default and what the timed region renders over) and on the device (bvh8_device.cuh, SURVEY.md
    §8(f) rank 1), wall-clock around the call, then two frames over each tree: they must be bit-identical, and the
    instrumented counters tell what the faster build costs in traversal work. Its own process, so that a fault in
    the newer device builder cannot take the bench line with it."""
    import time

    import numpy as np

    from chameleonrt_b200 import RenderCUDA

    scene, view = make_workload()
    out, frames = {}, {}
    for builder in ("host", "device"):
        r = RenderCUDA(0, max_depth=MAX_DEPTH, bvh_builder=builder, count_traversal=True, any_far_first=0)
        r.initialize(WIDTH, HEIGHT)
        t0 = time.perf_counter()
        r.set_scene(scene)
        wall = (time.perf_counter() - t0) * 1e3
        for f in range(2):
            r.render(*view, f == 0, False)
        info, c, st = r.scene_info(), r.counters(), r.stage_times()
        out[builder] = {"set_scene_ms": wall, "bvh_build_ms": info["bvh_build_ms"],
                        "phases_ms": {k: round(info[k + "_ms"], 3) for k in ("flatten", "sort", "tree", "emit", "pack")},
                        "ploc_rounds": int(info["ploc_rounds"]), "bvh8_nodes": int(info["bvh8_nodes"]),
                        "bvh8_depth": int(info["bvh8_depth"]),
                        "closest_nodes_per_ray": c["closest_nodes_visited"] / max(1, c["closest_rays"]),
                        "any_hit_nodes_per_ray": c["any_nodes_visited"] / max(1, c["occlusion_rays"]),
                        "traverse_ms_instrumented": st["traverse"] + st["traverse_primary"]}
        frames[builder] = r.read_accum()
        del r
    out["triangles"] = int(info["triangles"])
    out["frames_bit_identical"] = bool(np.array_equal(frames["host"].view(np.uint32), frames["device"].view(np.uint32)))
    emit(out)


def set_scene_probe(workload_key: str):
    """Runs run_set_scene_probe() in a child process; returns its dict, or {"error": ...}."""
    import subprocess

    try:
        # (CRT_BENCH_PROBE_SCRIPT: tests/test_bench_contract.py dry-runs this through its own launcher)
        script = os.environ.get("CRT_BENCH_PROBE_SCRIPT", os.path.abspath(__file__))
        p = subprocess.run([sys.executable, script, "--probe-set-scene", "--workload", workload_key],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=int(os.environ.get("CRT_BENCH_PROBE_TIMEOUT", "240")))
        if p.returncode != 0:
            return {"error": f"exit code {p.returncode}: " + p.stderr.decode(errors="replace").strip().splitlines()[-1][:300]}
        return json.loads(p.stdout.decode().strip().splitlines()[-1])
    except Exception as e:  # timeout, unparsable output
        return {"error": f"{type(e).__name__}: {e}"[:300]}


def emit(line: dict) -> None:
    """The contract is ONE JSON line on stdout. Libraries (NCCL prints its version banner) write
    to fd 1 too, so main() points fd 1 at stderr for the whole run and the result goes to the
    saved real stdout."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


_REAL_STDOUT = None


def main():
    global _REAL_STDOUT
    # watchdog: a run that does not finish is killed with a traceback instead of hanging its caller
    import faulthandler

    faulthandler.dump_traceback_later(int(os.environ.get("CRT_BENCH_WATCHDOG", "600")), repeat=False, file=sys.stderr,
                                      exit=True)
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-mode", action="store_true",
                    help="for runs under ncu: skip the instrumented counting pass and the CPU baseline")
    ap.add_argument("--probe-set-scene", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cuda" else max(args.warmup, 1)
    select_workload(args.workload)
    if args.probe_set_scene:
        run_set_scene_probe()
        return

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    from chameleonrt_b200 import RenderCUDA
    from chameleonrt_b200.distributed import FrameGatherer, PeerFrame

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; the backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"

    scene, view = make_workload()
    # a dedicated (non-default) torch stream: the renderer launches on it, so torch.cuda.Event
    # brackets exactly the kernels of the timed region
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    gpu = RenderCUDA(local_rank, max_depth=MAX_DEPTH, rank=rank, world_size=world, stream=stream.cuda_stream)
    gpu.initialize(WIDTH, HEIGHT)
    gpu.set_scene(scene)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # CRT_BENCH_FRAME=peer: no gather, every rank's resolve kernel writes into rank 0's frame over NVLink (PeerFrame;
    # opt-in until it has been timed: the gather is the path the committed multi-GPU numbers were measured with)
    peer_frame = os.environ.get("CRT_BENCH_FRAME", "gather") == "peer"
    gatherer = (PeerFrame(gpu) if peer_frame else FrameGatherer(gpu)) if world > 1 else None

    def frame(f, readback):
        """One step. N > 1: the frame-end gather of frame f is started here and overlaps the
        rendering of frame f+1 (it is completed by the next submit or by flush()); with
        readback the assembled frame is needed now, so the gather is finished immediately."""
        st = gpu.render(*view, f == 0, readback and world == 1)
        if world > 1:
            gatherer.submit()
            if readback:
                gatherer.finish()
                if rank == 0:
                    gpu.img[...] = gpu.read_img()
        return st

    def flush():
        if world > 1:
            gatherer.finish()

    # Frames in flight: with N GPUs each rank owns 1/N of the image, so N consecutive frames are
    # rendered as one wavefront (crtc_render_async num_frames) — every GPU keeps one full frame's worth
    # of samples in flight, whatever N. The result is bit-identical to frame-by-frame rendering; the
    # accumulated tiles are gathered once per batch. (The e2e region below stays frame by frame.)
    frames_in_flight = max(1, min(world, args.steps))
    # untimed frames before the timed region: W single frames + one batch (so that the path-state
    # buffers already have their batch size when the clock starts)
    n_warm = args.warmup + frames_in_flight

    # ---- device-timed region: inputs resident in HBM, no readback ----
    f = 0
    for _ in range(args.warmup):
        frame(f, False)
        f += 1
    flush()  # no collective in flight while the batch-sized path state is (re)allocated
    # The renderer has by now settled the descent order of its shadow rays for this scene (option any_far_first =
    # 2: frames 1 and 2 try one order each; the image does not depend on it). The counting pass must use the same.
    shadow_far_first = gpu.get_option("any_far_first_decision")

    # ---- instrumented pass (not timed): exact node/triangle visit counts of the same frames ----
    counts = np.zeros(6, np.float64)
    if not args.profile_mode:
        inst = RenderCUDA(local_rank, max_depth=MAX_DEPTH, rank=rank, world_size=world, count_traversal=True,
                          stream=stream.cuda_stream, any_far_first=1 if shadow_far_first == 1 else 0)
        inst.initialize(WIDTH, HEIGHT)
        inst.set_scene(scene)
        acc = np.zeros(6, np.float64)
        for fi in range(n_warm + args.steps):
            inst.render(*view, fi == 0, False)
            if fi >= n_warm:
                c = inst.counters()
                acc += [c["closest_rays"], c["closest_nodes_visited"], c["closest_tris_tested"], c["occlusion_rays"],
                        c["any_nodes_visited"], c["any_tris_tested"]]
        counts = acc
        del inst

    barrier()
    gpu.render_async(*view, False, frames_in_flight)  # warm-up batch
    f += frames_in_flight
    if world > 1:
        gatherer.submit()
        gatherer.finish()
    flush()
    barrier()
    gpu.sync()  # drop the warm-up batch's record
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        # frames are enqueued back to back (crtc_render_async): the host never waits inside the timed
        # region, so the GPU queue stays full (the host is kept at most two batches ahead)
        e0.record(stream)
        done = 0
        batch_done = []  # the host stays at most 2 batches ahead of the GPU (bounded launch queues)
        while done < args.steps:
            if len(batch_done) >= 2:
                batch_done[-2].synchronize()
            nb = min(frames_in_flight, args.steps - done)
            gpu.render_async(*view, f == 0, nb)
            f += nb
            done += nb
            if world > 1:
                # One gather per batch of N frames, completed right away (stream-ordered, no host wait):
                # ~15 MB over NVLink against >= 10 ms of rendering per batch, so overlapping it with the
                # next batch would buy < 1 % and is not worth keeping collectives in flight across launches.
                gatherer.submit()
                gatherer.finish()
            ev = torch.cuda.Event()
            ev.record(stream)
            batch_done.append(ev)
        flush()  # (no-op unless a gather is still pending)
        e1.record(stream)
        barrier()
    totals, stage_acc, csum, nframes = gpu.sync()
    assert nframes == args.steps
    rays = totals.num_rays
    n_batches = -(-args.steps // frames_in_flight)
    launches = csum["kernel_launches"] + (n_batches * world if (world > 1 and rank == 0 and not peer_frame) else 0)  # + k_assemble
    elapsed_ms = e0.elapsed_time(e1)
    clock_summary = clocks.summary()

    # ---- end-to-end region: the public call with host buffers (img readback every frame) ----
    barrier()
    t0 = time.perf_counter()
    e2e_rays = 0
    for _ in range(args.steps):
        st = frame(f, True)
        f += 1
        e2e_rays += st.num_rays
    barrier()
    e2e_s = time.perf_counter() - t0

    if world > 1:
        t = torch.tensor([elapsed_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, e2e_ms = float(t[0]), float(t[1])
        r = torch.tensor([float(rays), float(e2e_rays), float(launches)], dtype=torch.float64, device=dev)
        dist.all_reduce(r, op=dist.ReduceOp.SUM)
        rays, e2e_rays, launches = float(r[0]), float(r[1]), float(r[2])
        c = torch.tensor(counts, dtype=torch.float64, device=dev)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        counts = c.cpu().numpy()
        sa = torch.tensor([stage_acc[k] for k in sorted(stage_acc)], dtype=torch.float64, device=dev)
        dist.all_reduce(sa, op=dist.ReduceOp.MAX)
        stage_acc = {k: float(v) for k, v in zip(sorted(stage_acc), sa)}
    else:
        e2e_ms = e2e_s * 1e3

    if rank == 0:
        value = rays / (elapsed_ms * 1e3)  # MRays/s
        peak, peak_src = hbm_peak()
        # dominant kernel: k_traverse (persistent BVH8 traversal; every launch but the first carries
        # the shadow rays of bounce b and the continuation rays of bounce b+1)
        closest_bytes = counts[1] * S_NODE + counts[2] * S_TRI + counts[0] * (S_RAY + S_HIT)
        any_bytes = counts[4] * S_NODE + counts[5] * S_TRI + counts[3] * (S_RAY + 1)
        trav_bytes = closest_bytes + any_bytes
        t_trav_ms = stage_acc["traverse_primary"] + stage_acc["traverse"]  # max over ranks of per-rank sums
        n_launch = n_batches * (MAX_DEPTH + 1)  # k_traverse launches per rank in the timed region
        achieved = trav_bytes / world / (t_trav_ms * 1e-3) / 1e9  # per GPU GB/s
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic_k_traverse.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": "MRays/s", "value": value, "unit": "MRays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "width": WIDTH, "height": HEIGHT, "spp": SPP, "max_depth": MAX_DEPTH,
                       "parallelism": (f"image tiles 64x64 round-robin over {world} GPUs, scene replicated; {frames_in_flight} "
                                       "consecutive frames per wavefront (bit-identical to frame-by-frame), " +
                                       ("tiles written into rank 0's frame by every rank's resolve kernel through peer-mapped "
                                        "memory (no gather; a one-element all-reduce per batch as the barrier)" if peer_frame else
                                        "NCCL gather of the accumulated tiles to rank 0 once per batch (stream-ordered, not overlapped)") +
                                       "; e2e: frame by frame, assembly + readback every frame") if world > 1 else "single GPU",
                       "frames_in_flight": frames_in_flight,
                       "bvh_builder": ("host (binned SAH)", "device (PLOC)", "device (LBVH)")[gpu.get_option("bvh_builder")],
                       "traversal_variant": {"tri_pass_defer": gpu.get_option("tri_pass_defer"), "refill_idle": gpu.get_option("refill_idle"),
                                             "shade_sort": gpu.get_option("shade_sort")},
                       "shadow_ray_order": {1: "far-first", 0: "near-first"}.get(shadow_far_first, "undecided (near-first)") +
                                           " (chosen per scene from the traversal times of warm-up frames 1 and 2; same image either way)",
                       "l2": f"inputs larger than L2: ~{WIDTH * HEIGHT * SPP * 250 / 1e9:.1f} GB of per-frame path state "
                             f"streams through every bounce (L2 126 MB); scene = {gpu.scene_info()['node_bytes'] / 1e6:.0f} MB "
                             f"nodes + {2 * gpu.scene_info()['triangle_bytes'] / 1e6:.0f} MB triangle/shading records"},
            "roofline": {"kernel": "k_traverse", "bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": trav_bytes / n_launch / world,
                         "avg_launch_ms": t_trav_ms / n_launch,
                         "closest": {"rays": counts[0], "nodes_per_ray": counts[1] / max(1.0, counts[0]),
                                     "tris_per_ray": counts[2] / max(1.0, counts[0])},
                         "any_hit": {"rays": counts[3], "nodes_per_ray": counts[4] / max(1.0, counts[3]),
                                     "tris_per_ray": counts[5] / max(1.0, counts[3])}},
            "stage_ms_per_step": {k: v / args.steps for k, v in stage_acc.items()},
            "e2e": {"value": e2e_rays / (e2e_ms * 1e3), "unit": "MRays/s", "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": 52,  # ViewParams (camera basis + frame id) as kernel parameters
                    "d2h_bytes_per_step": WIDTH * HEIGHT * 4 + 36 * 4},
            "gpu_launches": int(launches),
            "clocks": clock_summary,
        }
        if not args.no_cpu_baseline and not args.profile_mode and world == 1:
            res, cpu_kind, cpu_desc = time_oracle(scene, view, budget_s=20.0, frames_cap=6)
            warm = res[1:] if len(res) > 1 else res
            cms = sum(r[0] for r in warm)
            crays = sum(r[1] for r in warm)
            line["cpu_baseline"] = {
                "value": crays / (cms * 1e3), "unit": "MRays/s", "cores": cpu_cores(), "kind": cpu_kind,
                "ms_per_frame": cms / len(warm),
                "sample": f"{len(warm)} full frame(s) of the same workload (1 warm-up frame discarded); {cpu_desc}"}
        if world == 1 and not args.profile_mode and os.environ.get("CRT_BENCH_SET_SCENE_PROBE", "1") != "0":
            # outside the timed region, in a child process: set_scene with the host and with the device BVH builder
            line["set_scene"] = set_scene_probe(args.workload)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
