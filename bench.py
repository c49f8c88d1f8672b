#!/usr/bin/env python
"""bench.py -- Mrays/s of the Rayito render hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c4|...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[3], the north-star target --
Stage 7 scene 1 (bumpy.obj, keyed transforms / motion blur, mirror BRDF) at
3840x2160, 256 spp (pixel samples hint 16), 1 light sample, ray depth 3.  One
"step" renders the whole frame once.  A "ray" is one scene.intersect /
scene.doesIntersect call of pathTrace: path segments + BSDF-MIS
probes + shadow rays.

value      whole-job Mrays/s, scene and camera resident in HBM, image left in HBM;
           CUDA events on the launching stream, max over ranks.
e2e        the same metric through the reference-facing C++ call Rayito::raytrace()
           (host scene -> prepare() -> flatten -> upload -> render -> image in host
           memory), wall clock, host<->device copies inside the timed region.
roofline   the traversal kernels (k_split_top_static, k_split_mesh, k_split_top;
           closest-hit and any-hit): algorithmic bytes per ray
           B = 48 + 32 N_pop + 36 N_tri + 16 N_shape + 40 K_xf (K_xf =
           evaluations of transforms that hold keys) over their summed CUDA-event
           time, against the measured HBM copy bandwidth in MEASURED_PEAKS.json.  The
           counts are the kernels' own exact work counters;
           tests/test_gpu_counters.py::test_render_counters_equal_oracle_on_recorded_rays
           requires them to equal the oracle's (oracle/port.c, pinned hit for hit to
           the compiled reference) on every ray the reference casts for a frame.
also       the default run adds config.also.c5: the 10 M-triangle mesh (configs[4]) at
           64 spp, device-timed, with its own roofline -- the HBM-bound configuration.
cpu_baseline / --impl reference
           the UNMODIFIED reference raytrace() (oracle/_ref, 16 worker threads by
           design) on the box's host cores, same scene and spp at a reduced
           resolution (throughput is a rate; the sample is stated).

The oracle is used here only as the timed CPU baseline, never on the measured path.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (recipe, width, height, pixel samples hint, light samples hint, depth, grid)
    "c4": dict(recipe=1, width=3840, height=2160, ps=16, ls=1, depth=3, grid=(0, 0),
               label="Rayito_Stage7 scene 1 (bumpy.obj, motion blur, mirror BRDF) 3840x2160 256spp ls1 depth3"),
    "c4-1080p": dict(recipe=1, width=1920, height=1080, ps=16, ls=1, depth=3, grid=(0, 0),
                     label="Rayito_Stage7 scene 1 1920x1080 256spp ls1 depth3"),
    "c4-small": dict(recipe=1, width=480, height=270, ps=16, ls=1, depth=3, grid=(0, 0),
                     label="Rayito_Stage7 scene 1 480x270 256spp ls1 depth3"),
    "scene2": dict(recipe=2, width=3840, height=2160, ps=16, ls=1, depth=3, grid=(0, 0),
                   label="Rayito_Stage7 scene 2 (falling spheres, tumbling boxes) 3840x2160 256spp"),
    "c5": dict(recipe=5, width=3840, height=2160, ps=32, ls=1, depth=3, grid=(2236, 2236),
               label="synthetic displaced sphere, 4 999 696 quads = 9 999 392 triangles, 3840x2160 1024spp"),
    "c5-64spp": dict(recipe=5, width=3840, height=2160, ps=8, ls=1, depth=3, grid=(2236, 2236),
                     label="synthetic displaced sphere, 4 999 696 quads = 9 999 392 triangles, 3840x2160 64spp"),
    "c5-small": dict(recipe=5, width=960, height=540, ps=8, ls=1, depth=3, grid=(2236, 2236),
                     label="synthetic displaced sphere, 9 999 392 triangles, 960x540 64spp (profiling size)"),
    # PERF MODE face BVH (binned SAH, rth_set_tree_mode(1)): measured parity, not bit-exact (tests/test_perf_tree.py);
    # its roofline counts the work done on ITS tree.  Never part of the default run.
    "c5-64spp-sah": dict(recipe=5, width=3840, height=2160, ps=8, ls=1, depth=3, grid=(2236, 2236), tree=1,
                         label="synthetic displaced sphere, 9 999 392 triangles, 3840x2160 64spp, PERF-MODE tree (binned SAH)"),
    # the default builds the face BVH of a mesh this large ON THE GPU inside the scene upload (rth_set_tree_mode(3),
    # rt_scene_create_ex); this workload forces the host build of the same tree (node for node, so `value` is the
    # same): its e2e leg shows what the host build and the node upload cost
    "c5-64spp-hostbuild": dict(recipe=5, width=3840, height=2160, ps=8, ls=1, depth=3, grid=(2236, 2236), tree=0,
                               label="synthetic displaced sphere, 9 999 392 triangles, 3840x2160 64spp, face BVH built on the host's cores"),
    "c3": dict(recipe=6, stage=6, width=1920, height=1080, ps=8, ls=1, depth=3, grid=(0, 0),
               label="Rayito_Stage6 scene (bumpy.obj, BVH, two area lights, Stage 6 rules) 1920x1080 64spp ls1 depth3"),
}
# Config C2: the Stage 3 program, pixel-sample sweep 1..256 spp at 512x512 (BASELINE.json configs[1])
C2_SWEEP = [(1, 1), (2, 2), (4, 4), (8, 8), (16, 16)]
WORKLOADS["c2"] = dict(stage23=3, width=512, height=512, sweep=C2_SWEEP,
                       label="Rayito_Stage3 built-in scene 512x512, pixel-sample sweep 1/4/16/64/256 spp, "
                             "2 area lights x 16 light samples")
CAMERA_SPEC = {       # fov, origin, target, up, focal distance, lens radius, shutter open/close (GUI defaults)
    1: [30, -4, 5, 15, 0, 0, 0, 0, 1, 0, 16, 0, 0, 1],
    2: [30, -4, 10, 30, 0, 5, 0, 0, 1, 0, 16, 0, 0, 1],
    5: [30, -4, 5, 15, 0, 0, 0, 0, 1, 0, 16, 0, 0, 1],
    6: [30, -2, 5, 15, 0, 0, 0, 0, 1, 0, 16, 0, 0, 0],
}
NEEDS_OBJ = (1, 6)
CPU_SAMPLE = dict(width=480, height=270)      # same scene, same spp, reduced resolution


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version
# banner on fd 1), so fd 1 is pointed at stderr for the whole run and the result line goes to the
# saved descriptor.
_RESULT_FD = None


def claim_stdout():
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region.  The sampler is started before the warm-up
    (nvidia-smi takes a few hundred ms to print its first line) and every line carries nvidia-smi's own
    timestamp; only the lines between begin() and end() count (a region shorter than the sampling period
    falls back to the lines nearest to it, and says so)."""
    QUERY = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        import datetime
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)            # let the line that covers the end of the region arrive
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                stamp = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((stamp, float(parts[2]), float(parts[3]), float(parts[4]),
                             [n for n, v in zip(names, parts[5:9]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        t0 = self.t0 if self.t0 is not None else -1e30
        t1 = self.t1 if self.t1 is not None else 1e30
        inside = [r for r in rows if t0 <= r[0] <= t1]
        note = None
        if not inside and rows:
            mid = 0.5 * (t0 + t1)
            inside = sorted(rows, key=lambda r: abs(r[0] - mid))[:2]
            note = "timed region shorter than the sampling period: the %d samples nearest to it (within %.2f s)" % (
                len(inside), max(abs(r[0] - mid) for r in inside))
        sm = sorted(r[1] for r in inside)
        reasons = set()
        for r in inside:
            reasons.update(r[4])
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(r[2] for r in inside) if inside else None,
               "power_w_max": max(r[3] for r in inside) if inside else None, "samples": len(inside), "reasons": sorted(reasons)}
        if note:
            out["note"] = note
        return out


def algorithmic_bytes(stats, strict=False):
    """B = 32 (ray in) + 16 (hit out) per ray + 32 per node popped
    + 36 per triangle tested + 16 per analytic shape tested + 40 per transform
    evaluation that reads keys.  A transform with no keys (the ShapeSet's own) reads
    nothing and is charged nothing; one with a single key reads that key (charged);
    strict=True charges only transforms with two or more keys (a real key PAIR)."""
    rays = stats["closest_rays"] + stats["any_rays"]
    return (48 * rays + 32 * stats["node_pops"] + 36 * stats["tri_tests"] + 16 * stats["shape_tests"]
            + 40 * stats["xform_pairs" if strict else "xform_keyed"])


def run_reference(args, wl, rank, world):
    """--impl reference: the unmodified reference raytrace() on the host cores."""
    if rank != 0:
        return
    from oracle import refapi
    from rayito_b200 import build
    import numpy as np
    cores = os.cpu_count() or 1
    obj = build.model_path("bumpy.obj") if wl["recipe"] in NEEDS_OBJ else None
    scene = refapi.RefScene(wl["recipe"], obj, wl["grid"], stage=wl.get("stage", 7))
    spec = np.array(CAMERA_SPEC[wl["recipe"]], np.float32)
    W, H = CPU_SAMPLE["width"], CPU_SAMPLE["height"]
    times, rays = [], 0
    for step in range(args.warmup + args.steps):
        _img, st = scene.render(spec, W, H, wl["ps"], ls=wl["ls"], depth=wl["depth"])
        if step >= args.warmup:
            times.append(st.render_seconds)
            rays = st.closest_calls + st.any_calls
        log("[reference] step %d: %.2f s, %.2f Mrays/s" % (step, st.render_seconds,
                                                           (st.closest_calls + st.any_calls) / st.render_seconds / 1e6))
    total = sum(times)
    value = rays * len(times) / total / 1e6
    threads = min(16, cores)
    sample = ("same scene/spp/ls/depth at %dx%d (%.1f M samples, %.1f M rays per step); reference raytrace() "
              "incl. prepare(); 16 worker threads by design (RaytraceMain.cpp:504-519) on %d host cores"
              % (W, H, W * H * wl["ps"] ** 2 / 1e6, rays / 1e6, cores))
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["label"], "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_stage23(args, wl, rank, world, local_rank):
    """Config C2.  A step = the whole sweep (five renders of the Stage 3 program).  The
    program is one serial Rng stream, so it does not shard: replicas only -- rank 0 runs it."""
    if rank != 0:
        return
    from oracle import refapi
    stage, W, H = wl["stage23"], wl["width"], wl["height"]
    if args.impl == "reference":
        # the reference is a single-threaded program; bounded sample: the 16 spp level it ships with
        times, rays = [], 0
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            _rgb, _rgb8, _flags, rays = refapi.stage_render(stage, W, H, 4, 4)
            dt = time.perf_counter() - t0
            log("[reference] step %d: %.2f s, %.2f Mrays/s" % (step, dt, rays / dt / 1e6))
            if step >= args.warmup:
                times.append(dt)
        value = rays * len(times) / sum(times) / 1e6
        sample = "Stage 3 program at its built-in 4x4 = 16 spp level only (%.1f M rays per step), single thread as written" % (rays / 1e6)
        emit({
            "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["label"], "sample": sample},
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": 1, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0})
        return

    import torch
    from rayito_b200 import capi
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the render core has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def sweep():
        out = []
        for (nu, nv) in wl["sweep"]:
            flush.zero_()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            _rgb, _rgb8, st = capi.stage23_render(stage, W, H, nu, nv, device=local_rank)
            out.append(dict(spp=nu * nv, wall_ms=1e3 * (time.perf_counter() - t0), rays=st.closest_rays,
                            device_ms=st.render_ms, prepass_ms=st.upload_ms, shade_ms=st.trace_ms,
                            launches=st.kernel_launches, rounds=st.trace_launches, samples=st.samples))
        return out

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        sweep()
    sampler.begin()
    runs = [sweep() for _ in range(args.steps)]
    sampler.end()
    clocks = sampler.stop()
    rays = sum(l["rays"] for l in runs[0])
    dev_ms = sum(l["device_ms"] for r in runs for l in r)
    wall_ms = sum(l["wall_ms"] for r in runs for l in r)
    shade_ms = sum(l["shade_ms"] for r in runs for l in r)
    launches = sum(l["launches"] for r in runs for l in r)
    samples = sum(l["samples"] for l in runs[0])
    peak, peak_src = load_peaks()
    # k_s23_shade reads 4 B (stream position) and writes 16 B (radiance) per pixel sample; everything
    # else (5 analytic shapes, 33 rays per hit sample) stays in registers and the constant bank
    bytes_per_step = 20.0 * samples
    achieved = bytes_per_step * args.steps / (shade_ms / 1e3) / 1e9
    per_level = [{k: (round(v, 3) if isinstance(v, float) else v) for k, v in l.items()} for l in runs[-1]]
    cpu = None
    if not args.no_cpu_baseline:
        t0 = time.perf_counter()
        _rgb, _rgb8, _flags, r16 = refapi.stage_render(stage, W, H, 4, 4)
        dt = time.perf_counter() - t0
        cpu = {"value": r16 / dt / 1e6, "unit": "Mrays/s", "cores": 1, "kind": "reference", "seconds": dt,
               "sample": "unmodified Stage 3 code (oracle/_ref/libref_s3.so) at the program's built-in 16 spp level, "
                         "%.1f M rays, single thread as written" % (r16 / 1e6)}
    line = {
        "metric": "Mrays/s", "value": rays * args.steps / (dev_ms / 1e3) / 1e6, "unit": "Mrays/s", "n_gpus": 1,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["label"], "parallelism": "replicas only (one serial Rng stream per image)",
                   "rays_per_step": rays, "samples_per_step": samples, "levels": per_level,
                   "l2": "256 MB flush buffer written before every render"},
        "clocks": clocks, "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "k_s23_shade", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "peak_source": peak_src, "traffic": None,
                     "note": "register-resident FP32 work (5 analytic shapes in the constant bank, 33 rays per hit "
                             "sample, sequential MWC draws): neither HBM nor tensor bound; the HBM fraction is "
                             "reported because the contract asks for one of the two"},
        "e2e": {"value": rays * args.steps / (wall_ms / 1e3) / 1e6, "unit": "Mrays/s",
                "h2d_bytes_per_step": 5 * 2048, "d2h_bytes_per_step": 5 * W * H * 15,
                "includes": "rth_stage23_render per level: scene build, device allocation, stream pre-pass, shading, "
                            "float + 8-bit image download"},
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    emit(line)


def load_traffic(workload):
    """Measured DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per traversal launch at THIS
    workload's batch size, from the committed ncu capture summarised by tools/ncu_traffic.py into
    profiles/traffic.json; None when that workload was not captured."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f).get(workload)


class Env:
    """Process-wide handles shared by the measurements of one bench.py run."""

    def __init__(self, args):
        import numpy as np
        import torch
        import torch.distributed as dist
        from rayito_b200 import build, capi
        self.np, self.torch, self.dist, self.build, self.capi = np, torch, dist, build, capi
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the render core has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        self.dev = torch.device("cuda", self.local_rank)
        # the render core's own communicator for the tile assembly (rt_render_multi): rank 0 draws the
        # NCCL unique id, torch.distributed only carries its 128 bytes to the other ranks
        self.comm = None
        if self.world > 1:
            uid = torch.zeros(capi.Comm.ID_BYTES, dtype=torch.uint8, device=self.dev)
            if self.rank == 0:
                uid.copy_(torch.frombuffer(bytearray(capi.Comm.unique_id()), dtype=torch.uint8))
            dist.broadcast(uid, src=0)
            self.comm = capi.Comm(uid.cpu().numpy().tobytes(), self.rank, self.world, self.local_rank)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)     # > 126 MB L2
        self.stream = torch.cuda.current_stream(self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def allsum(self, x):
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def allmax(self, x):
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def measure_workload(env, name, steps, warmup, want_e2e, want_cpu):
    """One workload, device-timed: returns the fields of a result line (rank 0 uses them)."""
    args, capi, torch, dist = env.args, env.capi, env.torch, env.dist
    rank, world, local_rank, dev = env.rank, env.world, env.local_rank, env.dev
    wl = WORKLOADS[name]

    # ---- scene: built with the C++ host API, flattened, uploaded once ------------
    obj = env.build.model_path("bumpy.obj") if wl["recipe"] in NEEDS_OBJ else None
    t0 = time.perf_counter()
    # the product's default (TREE_AUTO): meshes of 65 536 faces or more get their face BVH built on the device
    # during the upload, the same tree node for node (tests/test_gpu_build.py)
    tree = wl.get("tree", capi.TREE_AUTO)
    hscene = capi.HostScene(wl["recipe"], obj, wl["grid"], tree=tree)
    host_prepare_s = time.perf_counter() - t0
    dscene = capi.DeviceScene(hscene.desc, device=local_rank, build_bvh_on_device=tree in (capi.TREE_DEVICE, capi.TREE_AUTO))
    spec = hscene.default_camera_spec()
    cam = capi.camera_from_spec(spec)
    W, H, ps, ls, depth = wl["width"], wl["height"], wl["ps"], wl["ls"], wl["depth"]
    scene_bytes = hscene_bytes(hscene)

    image = torch.zeros((H, W, 3), dtype=torch.float32, device=dev)
    flush, stream = env.flush, env.stream
    extra_flags = int(os.environ.get("RT_BENCH_FLAGS", "0"))      # A/B switches (RT_RENDER_* bits), not for reported runs

    shard_rank, shard_world = rank, world
    if args.shard and world == 1:
        shard_rank, shard_world = (int(v) for v in args.shard.split("/"))

    def params(flags=0):
        return capi.RtRenderParams(W, H, ps, ls, depth, args.tile, shard_rank, shard_world, args.batch, flags | extra_flags)

    assemble_ms = []

    def step(flags=0):
        if world == 1:
            return dscene.render_device(cam, params(flags), image.data_ptr(), stream.cuda_stream)
        # the one collective of the path, inside the product call: every rank renders its tiles packed,
        # ncclSend / ncclRecv moves them to rank 0, one kernel there scatters them into the frame
        st, ms = env.comm.render_multi(dscene, cam, params(flags), image.data_ptr(), 0, stream.cuda_stream)
        assemble_ms.append(ms)
        return st

    # exact work counters (deterministic per frame) from one untimed instrumented step
    image.zero_()
    counted = step(capi.RT_RENDER_COUNT_WORK).as_dict()
    log("[rank %d] %s counted step: %s" % (rank, name, counted))

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(warmup):
        image.zero_()
        flush.zero_()
        step()

    env.barrier()
    sampler.begin()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    trace_ms = 0.0
    trace_launches = 0
    render_ms = 0.0
    ev0.record(stream)
    for _ in range(steps):
        image.zero_()
        flush.zero_()             # L2 flush between timed steps
        st = step(capi.RT_RENDER_TIME_TRACE)
        launches += st.kernel_launches
        trace_ms += st.trace_ms
        trace_launches += st.trace_launches
        render_ms += st.render_ms
    ev1.record(stream)
    env.barrier()
    sampler.end()
    clocks = sampler.stop()
    elapsed_ms = ev0.elapsed_time(ev1)

    # ---- aggregate over ranks ------------------------------------------------------
    allsum, allmax = env.allsum, env.allmax
    rays_rank = counted["closest_rays"] + counted["any_rays"]
    rays_total = allsum(rays_rank)
    samples_total = allsum(counted["samples"])
    max_ms = allmax(elapsed_ms)
    launches_total = allsum(launches)
    value = rays_total * steps / (max_ms / 1e3) / 1e6
    # traversal roofline: sum over ranks of algorithmic bytes / max over ranks of traversal time
    bytes_rank = algorithmic_bytes(counted)
    bytes_total = allsum(bytes_rank)
    bytes_strict_total = allsum(algorithmic_bytes(counted, strict=True))
    trace_ms_max = allmax(trace_ms)
    peak, peak_src = load_peaks()
    achieved = bytes_total * steps / (trace_ms_max / 1e3) / 1e9 / world    # per GPU
    launches_per_step = trace_launches / max(steps, 1)
    traffic = load_traffic(name) if world == 1 and not args.batch and not args.tile else None
    roofline = {
        "bound": "hbm",
        "kernel": "traversal passes: k_split_top_static (tabulated top-level walk) + k_split_mesh (face BVH) + "
                  "k_split_top (resume), closest-hit and any-hit instantiations",
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
        "frac_strict": bytes_strict_total * steps / (trace_ms_max / 1e3) / 1e9 / world / peak,
        "traffic": traffic["bytes_per_launch"] if traffic else None,
        "traffic_detail": traffic,
        "bytes_per_ray": bytes_total / rays_total,
        "bytes_per_ray_strict": bytes_strict_total / rays_total,
        "per_ray": {k: allsum(counted[k]) / rays_total for k in
                    ("node_pops", "tri_tests", "shape_tests", "xform_evals", "xform_keyed", "xform_pairs")},
        "counters": "the kernels' own exact work counters (one RT_RENDER_COUNT_WORK step); "
                    "tests/test_gpu_counters.py requires them to equal the oracle's (oracle/port.c) on the same rays",
        "bytes_formula": "48 per ray + 32 per node popped + 36 per triangle tested + 16 per analytic shape tested + "
                         "40 per transform evaluation that reads keys (xform_keyed: >= 1 key; frac_strict charges only "
                         "xform_pairs: >= 2 keys); keyless transforms (the ShapeSet's own) charge nothing",
        "launches_per_step": launches_per_step,
        "bytes_per_launch": bytes_rank / max(launches_per_step, 1),
        "avg_launch_ms": trace_ms / max(trace_launches, 1),
        "trace_share_of_step": trace_ms_max / max_ms,
        "trace_mrays_per_s_per_gpu": rays_total * steps / (trace_ms_max / 1e3) / 1e6 / world,
        "note": ("scene is %.1f MB: %s" % (scene_bytes / 1e6,
                 "L2-resident by nature of this config, so the HBM fraction is reported as the contract asks but "
                 "L2 latency / SIMT divergence binds first (SURVEY.md 8d)" if scene_bytes < 100e6 else
                 "HBM-resident, random node/triangle gathers: the HBM roofline is the binding one")),
    }

    # free the device scene and wavefront state before the end-to-end leg builds its own
    dscene.close()
    del image

    # ---- e2e: through Rayito::raytrace() with host buffers -------------------------
    e2e = None
    if want_e2e:
        e2e = measure_e2e(args, wl, capi, obj, spec, rank, world, local_rank, dev, dist, torch, env.np, rays_total,
                          scene_bytes, steps, env.comm)

    # ---- CPU baseline (rank 0, N=1 only) -------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and want_cpu:
        cpu = measure_cpu_baseline(wl, obj, spec)
    hscene.close()

    out = {
        "value": value, "ms_per_step": max_ms / steps, "steps": steps, "warmup": warmup,
        "config": {"workload": wl["label"],
                   "parallelism": "screen tiles x%d (lattice interleave (tx + m ty) mod world), scene replicated%s" % (
                       world, "" if world == 1 else "; tile assembly on rank 0 by rt_render_multi (packed tiles, grouped "
                       "ncclSend/ncclRecv, scatter kernel): %.3f ms per step on rank 0" % (
                           sum(assemble_ms[-steps:]) / max(steps, 1))),
                   "samples_per_step": samples_total, "rays_per_step": rays_total,
                   "rays_per_sample": rays_total / samples_total,
                   "msamples_per_s": samples_total * steps / (max_ms / 1e3) / 1e6,
                   "l2": "256 MB flush buffer written between timed steps; per-batch path state (tens of GB) >> L2",
                   "host_prepare_s": host_prepare_s, "render_ms_per_step_rank0": render_ms / steps},
        "clocks": clocks, "gpu_launches": int(launches_total), "roofline": roofline,
    }
    if e2e is not None:
        out["e2e"] = e2e
    if cpu is not None:
        out["cpu_baseline"] = cpu
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary C5 measurement of the default run")
    ap.add_argument("--batch", type=int, default=0, help="max samples per wavefront batch (0 = core default)")
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--shard", default="", help="R/W: render only rank R's screen tiles of a W-rank partition on this one GPU "
                                                "(what one rank of a W-GPU run does, without the other ranks; for A/B runs)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    claim_stdout()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if "stage23" in wl:
        run_stage23(args, wl, rank, world, local_rank)
        return
    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return

    env = Env(args)
    head = measure_workload(env, args.workload, args.steps, args.warmup, not args.no_e2e, not args.no_cpu_baseline)
    # The HBM-bound configuration (BASELINE.json configs[4], the 10 M-triangle mesh) rides along with the
    # default run at 64 spp (the full 1024 spp frame takes 18 s): its own device-timed value and roofline,
    # under config.also.c5.  C4 stays the headline.
    also = None
    if args.workload == "c4" and not args.no_also:
        also = measure_workload(env, "c5-64spp", max(2, min(args.steps, 3)), 3, False, False)

    if env.rank == 0:
        line = {
            "metric": "Mrays/s", "value": head["value"], "unit": "Mrays/s", "n_gpus": env.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": head["config"], "clocks": head["clocks"], "gpu_launches": head["gpu_launches"],
            "roofline": head["roofline"],
        }
        for key in ("e2e", "cpu_baseline"):
            if key in head:
                line[key] = head[key]
        if also is not None:
            line["config"]["also"] = {"c5": {
                "metric": "Mrays/s", "value": also["value"], "unit": "Mrays/s", "ms_per_step": also["ms_per_step"],
                "steps": also["steps"], "warmup": also["warmup"], "config": also["config"], "clocks": also["clocks"],
                "gpu_launches": also["gpu_launches"], "roofline": also["roofline"]}}
            line["gpu_launches"] += also["gpu_launches"]
        emit(line)
    if env.world > 1:
        env.dist.destroy_process_group()


def hscene_bytes(hscene):
    d = hscene.desc.contents
    return (d.num_mesh_nodes * 32 + d.num_top_nodes * 32 + d.num_vertices * 12 + d.num_normals * 12
            + d.num_indices * 8 + d.num_faces * 8 + d.num_cdf * 4 + d.num_keys * 44)


def measure_e2e(args, wl, capi, obj, spec, rank, world, local_rank, dev, dist, torch, np, rays_total, scene_bytes, steps,
                comm=None):
    """Wall-clock Mrays/s through the reference-facing call Rayito::raytrace(scene, cam, W, H, ps, ls, depth)
    (rth_app_raytrace), every step: findLights + prepare() (host BVH builds) + flatten + scene upload + render +
    image download.  The application's scene-building code (OBJ read, recipe) runs once before the timed
    region, as it does for the reference arm and the cpu_baseline (which time the reference's raytrace()).
    N = 1: host image out of raytrace().  N > 1: every rank calls rayito_b200::raytraceMulti() on its copy of
    the scene (prepare + upload + render of its tiles), the packed tiles go to rank 0 over NCCL inside that
    call, and rank 0's call returns the assembled host Image."""
    import ctypes as C
    W, H, ps, ls, depth = wl["width"], wl["height"], wl["ps"], wl["ls"], wl["depth"]
    stats = capi.RtRenderStats()
    lib = capi.host()
    path = obj.encode() if obj else None
    t0 = time.perf_counter()
    app = lib.rth_app_create(wl["recipe"], path, wl["grid"][0], wl["grid"][1])
    if not app:
        raise RuntimeError("rth_app_create: " + lib.rth_last_error_string().decode())
    build_s = time.perf_counter() - t0
    pixels = C.c_void_p()

    def one():
        if world == 1:
            # the Image raytrace() returns stays with the app handle: the host frame is read in place
            rc = lib.rth_app_raytrace_image(app, spec.ctypes.data, W, H, ps, ls, depth, local_rank, rank, world, 0,
                                            C.byref(pixels), C.byref(stats))
        else:
            # rayito_b200::raytraceMulti(): every rank prepares and renders its tiles, rank 0 gets the Image
            rc = lib.rth_app_raytrace_multi(app, spec.ctypes.data, W, H, ps, ls, depth, comm.handle, 0,
                                            C.byref(pixels), C.byref(stats))
        if rc != 0:
            raise RuntimeError("rth_app_raytrace: " + lib.rth_last_error_string().decode())
        if rank == 0:
            # the step's result is in host memory: touch it (mean of the first row) like a consumer would
            row = np.ctypeslib.as_array(C.cast(pixels, C.POINTER(C.c_float)), shape=(W * 3,))
            if not np.isfinite(float(row.mean())):
                raise RuntimeError("e2e frame is not finite")
        return stats.render_ms

    lib.rth_set_tree_mode(wl.get("tree", capi.TREE_AUTO))
    one()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    n = max(1, steps)
    t0 = time.perf_counter()
    for _ in range(n):
        one()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - t0
    lib.rth_set_tree_mode(capi.TREE_AUTO)
    lib.rth_app_destroy(app)
    t = torch.tensor([wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall = float(t.item())
    return {"value": rays_total * n / wall / 1e6, "unit": "Mrays/s",
            "h2d_bytes_per_step": int(scene_bytes) * world, "d2h_bytes_per_step": int(W * H * 12),
            "steps": n, "ms_per_step": 1e3 * wall / n, "scene_build_s": build_s,
            "includes": "Rayito::raytrace() per step: findLights + prepare() (host BVH build) + flatten + scene upload + "
                        "render + image download" + ("" if world == 1 else " (raytraceMulti per rank: NCCL tile assembly "
                        "on rank 0 inside the call, one download)") + "; the application's scene-building code (OBJ read) runs once, "
                        "untimed, as for the reference arm.  Unlike the device-timed `value` loop there is no L2 flush "
                        "and no per-stage CUDA-event timing between the kernels of a call, so e2e can come out a "
                        "percent above `value` on the same box"}


def measure_cpu_baseline(wl, obj, spec):
    from oracle import refapi
    if not refapi.available(wl.get("stage", 7)):
        return {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
    cores = os.cpu_count() or 1
    scene = refapi.RefScene(wl["recipe"], obj, wl["grid"], stage=wl.get("stage", 7))
    W, H = CPU_SAMPLE["width"], CPU_SAMPLE["height"]
    if wl["recipe"] == 5:
        W, H = 240, 135
    _img, st = scene.render(spec, W, H, wl["ps"], ls=wl["ls"], depth=wl["depth"])
    rays = st.closest_calls + st.any_calls
    return {"value": rays / st.render_seconds / 1e6, "unit": "Mrays/s", "cores": min(16, cores), "kind": "reference",
            "host_cores": cores, "seconds": st.render_seconds, "prepare_seconds": st.prepare_seconds,
            "sample": "unmodified reference raytrace() (oracle/_ref), same scene/spp/ls/depth at %dx%d = %.1f M samples, "
                      "%.1f M rays; includes prepare(); reference uses exactly 16 worker threads" % (
                          W, H, W * H * wl["ps"] ** 2 / 1e6, rays / 1e6)}


if __name__ == "__main__":
    main()
