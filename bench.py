#!/usr/bin/env python
"""bench.py — frames/s (and Gtris/s) of the geometry hot path: two-pass meshlet cull + visbuffer raster + HiZ build.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one frame of a camera sweep over the synthetic scene of BASELINE.json config 3 (the configuration the
north-star target is quoted on: 10x10x10 lattice of a 100,352-triangle patch = 100.35M triangles, 3840x2160, two-pass HiZ).
Multi-GPU = independent views sharded per GPU: the scene is replicated, every rank renders its own K
views, no data-path collective; value = (N*K frames) / max-over-ranks device time  ("scaling": "weak").

--impl reference times the CPU implementation of the same path on the host cores (the oracle port; the reference's own
shaders cannot run here: no Vulkan ICD / glslang — DESIGN.md §Oracle), rank 0 only.

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CONFIGS = {
    # id: (label, builder kwargs, resolution)
    1: ("cfg1: icosphere 64,980 tris, 640x480", dict(kind="icosphere", frequency=57), (640, 480)),
    2: ("cfg2: atrium 262,144 tris int16-quantised, 1920x1080", dict(kind="atrium", detail=128), (1920, 1080)),
    3: ("cfg3: 10x10x10 lattice of a 224x224-quad patch (100.35M tris), 3840x2160, two-pass HiZ", dict(kind="lattice", n=(10, 10, 10), quads=224), (3840, 2160)),
    4: ("cfg4: city 50x40 unique buildings x ~10k tris (~20M tris), 1920x1080, 64-view sweep", dict(kind="city", n=(50, 40), tris=10000), (1920, 1080)),
    # cfg 4 with KHR_mesh_quantization-style int16 positions (--positions f32|i16 then selects what the rasteriser reads)
    41: ("cfg4q: city 50x40 unique buildings x ~10k tris (~20M tris), int16-quantised, 1920x1080, 64-view sweep", dict(kind="cityq", n=(50, 40), tris=10000), (1920, 1080)),
    5: ("cfg5: 22x22x21 lattice (1.02B tris), 7680x4320", dict(kind="lattice", n=(22, 22, 21), quads=224), (7680, 4320)),
}
NVIEWS = 64
SHARD_BLOCK_LOG2 = int(os.environ.get("VKV_SHARD_BLOCK_LOG2", "11"))   # range sharding: blocks of 2^k MeshletDraws dealt round-robin to the ranks


def build_scene(spec):
    from vk_gltf_viewer_b200.scene import Scene
    k = spec["kind"]
    if k == "icosphere":
        return Scene.icosphere(spec["frequency"])
    if k == "atrium":
        return Scene.atrium(spec["detail"])
    if k == "lattice":
        return Scene.lattice(*spec["n"], spec["quads"], 0x5EED0003)
    if k == "city":
        return Scene.city(*spec["n"], spec["tris"], 0x5EED0004)
    if k == "cityq":
        return Scene.city_quantized(*spec["n"], spec["tris"], 0x5EED0004)
    raise ValueError(k)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md): the sampler runs from
    before the warm-up, every row is time-stamped on arrival and only rows inside [t0, t1] of the timed region count."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.idx)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.03 and len(r) >= 9]
        scope = "timed region"
        if len(rows) < 3:  # very short runs: fall back to everything since the warm-up started
            rows = [r for (t, r) in self.rows if len(r) >= 9]
            scope = "warm-up + timed region"
        num = lambda v: float(v) if v.replace(".", "", 1).isdigit() else None
        sm = [x for x in (num(r[1]) for r in rows) if x is not None]
        mx = [x for x in (num(r[2]) for r in rows) if x is not None]
        pw = [x for x in (num(r[3]) for r in rows) if x is not None]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            for n, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(pw) if pw else None, "scope": scope}


def config_dict(label, W, H, cnt, passes, builder_note, world, shard):
    """the `config` object of the JSON line: IDENTICAL on the GPU arm and on the reference arm (it names the workload and how each
    arm runs it; per-run observations live under `run`)"""
    return {"workload": label, "resolution": [W, H], "meshlet_draws": cnt.draws, "triangles": cnt.triangles_instanced, "passes": passes,
            "meshlets": builder_note, "views": f"{NVIEWS}-view camera sweep, one view per step",
            "parallelism": ((f"GPU arm: views sharded over {world} GPU(s), scene replicated, no collective" if shard == "views" else
                             f"GPU arm: one view, MeshletDraw list sharded over {world} GPU(s) in interleaved 2048-draw blocks, screen-strip owners "
                             "pull dirty visbuffer tiles over NVLink peer memory and all-gather the pyramid") +
                            "; reference arm: CPU port of the same path on all host cores, rank 0 only"),
            "l2": "GPU arm: 256 MB scratch written between timed frames (L2 flush), each frame timed by its own CUDA event pair; reference arm: n/a (CPU)"}


def roofline(dom, stages, ent, hbm, peak_src, traffic, clocks):
    """The dominant kernel against the limit that actually bounds it.  Cull and raster are bound by instruction ISSUE:
    achieved = warp-instructions per launch (counted by ncu for this workload, profiles/traffic.json) / live CUDA-event duration,
    peak = SMs x 4 schedulers x SM clock.  The HBM figure (algorithmic input bytes / time against the measured copy bandwidth) is kept
    beside it; for the HBM-bound kernels (pyramid build, clear) it is the primary one."""
    st = stages[dom]
    hbm_part = {"achieved": st["GB/s"], "peak": hbm, "unit": "GB/s", "frac": st["frac_hbm"], "peak_source": peak_src,
                "note": "input-side algorithmic bytes (SURVEY §8d) / CUDA-event time"}
    issue_bound = dom.startswith("raster") or dom.startswith("cull")
    if issue_bound and ent and ent.get("warp_instructions"):
        mhz = (clocks or {}).get("sm_max_mhz") or 1965.0
        peak = 148 * 4 * mhz * 1e6 / 1e9          # G warp-instructions / s
        ach = ent["warp_instructions"] / (st["ms"] * 1e-3) / 1e9
        return {"bound": "issue", "kernel": dom, "achieved": round(ach, 1), "peak": round(peak, 1), "unit": "Gwarp-inst/s", "frac": round(ach / peak, 4),
                "traffic": traffic, "hbm": hbm_part, "warp_instructions_per_launch": ent["warp_instructions"], "instruction_count_from": ent.get("capture"),
                "note": "software rasteriser / cull: bound by SM instruction issue (and L2 atomics), working set L2 resident (dram traffic << algorithmic "
                        "bytes); peak = 148 SMs x 4 issue slots x SM clock; `hbm` keeps the byte-based figure; other kernels' fractions are in `stages`"}
    return {"bound": "hbm", "kernel": dom, "traffic": traffic, **hbm_part}


def range_sharded_leg(rank, world, local_rank, dist, steps, builder_note):
    """BASELINE config 5 on `world` GPUs: ONE view of the 1.02-billion-triangle lattice at 7680x4320, the MeshletDraw list dealt to
    the ranks in interleaved 2048-draw blocks, owners of interleaved 16-row screen strips pulling dirty tiles over NVLink and all-gathering the pyramid
    (VKV_FRAME_MERGE_STRIPS, csrc/strips.cu).  Returns the `range_sharded` object of the JSON line (every rank computes it; rank 0
    prints): device-timed ms per frame (max over ranks), the merge stages, the same frames on ONE GPU for the speed-up, the
    NVLink bytes per frame against the all-reduce bound, and `merge_parity`: on EVERY rank, after the same three views rendered
    from a cleared pyramid, the 64-bit digest of the rank's own strip and of the whole pyramid must equal the digests of the same
    frames rendered unsharded by that rank alone."""
    import torch
    from vk_gltf_viewer_b200 import api, multigpu
    from vk_gltf_viewer_b200.scene import Camera
    label, spec, (W, H) = CONFIGS[5]
    scene = build_scene(spec)
    cnt = scene.counts()
    views = [scene.default_view(i, NVIEWS) for i in range(NVIEWS)]
    r = api.Renderer(W, H, device=local_rank)
    cam = Camera(W, H)
    cam.look_at(*views[0])
    pc = r.upload_scene(scene, cam)
    cam_addrs = []
    for v in views[1:steps + 8]:
        cam.look_at(*v)
        cam_addrs.append(r.upload(np.frombuffer(cam.raw(), np.uint8)))
    first_cam = pc.cameraBuffer
    zeros = np.zeros(r.pyramid_floats, np.float32)

    def sync_all():
        r.sync()
        dist.barrier()
        torch.cuda.synchronize()

    def sweep(flags, n, timed=False, stages=None):
        """n frames of the sweep from a cleared pyramid; returns the summed device ms (FRAME_TIMED) and the last stats"""
        r._ck(r.L.vkv_write_pyramid(r.h, zeros.ctypes.data, zeros.size))
        pc.cameraBuffer = first_cam
        r.frame(pc, flags)
        total, st, pulled, sent = 0.0, None, 0, 0
        for k in range(n):
            pc.cameraBuffer = cam_addrs[k % len(cam_addrs)]
            if timed:
                r.flush_l2(256 << 20)
            st = r.frame(pc, flags | (api.FRAME_TIMED if timed else 0) | (api.FRAME_STAGES if stages is not None else 0))
            total += st.total_ms
            pulled += st.strip_tiles_pulled; sent += st.strip_texels_sent
            if stages is not None:
                for name in stages:
                    stages[name] += getattr(st, name)
        return total, st, pulled, sent

    sharded = api.FRAME_TWO_PASS | api.FRAME_MERGE_STRIPS
    single = api.FRAME_TWO_PASS
    r.set_shard_interleaved(rank, world, SHARD_BLOCK_LOG2)
    multigpu.attach_peers(r, dist)
    # ---- parity first (three views from a cleared pyramid, sharded vs the same rank alone)
    sync_all()
    sweep(sharded, 2)
    h_strip, h_pyr = r.hash(0, rank, world), r.hash(1)
    sync_all()
    r.set_shard_interleaved(0, 1, SHARD_BLOCK_LOG2)   # the whole list on this GPU, no exchange
    sweep(single, 2)
    ok = (h_strip == r.hash(0, rank, world)) and (h_pyr == r.hash(1))
    # ---- one GPU: the same frames, device-timed (every rank measures; they are independent here)
    n1 = max(3, min(steps, 10))
    sweep(single, 3)
    n1_ms, _, _, _ = sweep(single, n1, timed=True)
    sync_all()
    # ---- sharded: device-timed
    r.set_shard_interleaved(rank, world, SHARD_BLOCK_LOG2)
    sweep(sharded, 3)
    sync_all()
    ms, st, pulled, sent = sweep(sharded, steps, timed=True)
    sync_all()
    names = ("cull_a_ms", "raster_a_ms", "merge_a_ms", "hiz_a_ms", "cull_b_ms", "raster_b_ms", "merge_b_ms", "hiz_b_ms")
    stg = {k: 0.0 for k in names}
    ks = min(steps, 20)
    sweep(sharded, ks, timed=True, stages=stg)
    sync_all()
    t = torch.tensor([ms / steps, n1_ms / n1] + [stg[k] / ks for k in names], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tr = torch.tensor([pulled / steps, sent / steps], device="cuda", dtype=torch.float64)
    trmax = tr.clone()
    dist.all_reduce(tr)                          # whole job
    dist.all_reduce(trmax, op=dist.ReduceOp.MAX)  # busiest rank
    okt = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    r.ipc_detach()
    r.close()
    vals = t.tolist()
    bound = 2.0 * (world - 1) / world * 8 * W * H
    per_rank_bytes = float(trmax[0]) * 8192 + float(trmax[1]) * 4
    out = {"workload": label, "resolution": [W, H], "meshlet_draws": cnt.draws, "triangles": cnt.triangles_instanced, "meshlets": builder_note,
           "n_gpus": world, "steps": steps, "ms_per_frame": vals[0], "frames_per_s": 1e3 / vals[0], "gtris_per_s": cnt.triangles_instanced / vals[0] / 1e6,
           "n1_ms_per_frame": vals[1], "speedup_vs_n1": vals[1] / vals[0], "merge_parity": bool(int(okt.item())),
           "merge_parity_how": "per rank: vkv_hash of the visbuffer rows it owns and of the whole pyramid after 3 views from a cleared pyramid == the same frames rendered unsharded on that rank",
           "stages_ms_max_over_ranks": {k[:-3]: round(v, 5) for k, v in zip(names, vals[2:])},
           "merge_a_ms": vals[2 + names.index("merge_a_ms")], "merge_b_ms": vals[2 + names.index("merge_b_ms")],
           "nvlink": {"tiles_pulled_per_frame_all_ranks": float(tr[0]), "pyramid_texels_sent_per_frame_all_ranks": float(tr[1]),
                      "bytes_per_frame_busiest_rank": per_rank_bytes, "allreduce_bound_bytes_per_rank": bound,
                      "fraction_of_allreduce_bound": per_rank_bytes / bound,
                      "note": "received: 8 KB per (64x16-pixel tile, peer that drew into it) inside the rank's strip; sent: 4 B per changed pyramid texel per peer; "
                              "bound = 2*(n-1)/n * 8*W*H, what a ring all-reduce of the 64-bit visbuffer moves per rank per merge (two merges per frame)"},
           "scaling": "strong", "l2": "256 MB scratch written between timed frames"}
    return out


def meshlet_averages(scene):
    """average vertices / triangles per MeshletDraw of the scene (weighted by how often each primitive is drawn)"""
    d = scene.draws()
    per_prim = np.bincount(d["primitiveIndex"].astype(np.int64), minlength=scene.counts().primitives)
    v = t = n = 0.0
    for i, draws in enumerate(per_prim):
        if not draws:
            continue
        ml = scene.primitive(i)["meshlets"]
        inst = draws / max(1, ml.shape[0])
        v += inst * float(ml["vertexCount"].astype(np.int64).sum()); t += inst * float(ml["triangleCount"].astype(np.int64).sum()); n += draws
    return (v / n, t / n) if n else (0.0, 0.0)


def cpu_frames(scene, W, H, views, nframes, threads=0, warmup=1, stage_s=None):
    """the oracle (CPU port of the reference path) timed on the host cores: two-pass frames of the same sweep"""
    from tests import oracle_lib as O
    from vk_gltf_viewer_b200.scene import Camera
    cam = Camera(W, H)
    cam.look_at(*views[0])
    pc = scene.host_push_constants(cam)
    tg = O.Targets(W, H)
    O.lib().orc_set_diagnostics(0)    # time the path, not the oracle's ambiguity bookkeeping
    for _ in range(max(1, warmup)):   # the first one also fills the pyramid the first timed frame culls against
        O.frame(pc, tg, two_pass=True, threads=threads)
    times = []
    for k in range(nframes):
        cam.look_at(*views[(k + 1) % len(views)])
        t0 = time.perf_counter()
        O.frame(pc, tg, two_pass=True, threads=threads, stage_s=stage_s)
        times.append(time.perf_counter() - t0)
    return times


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default: 200 frames on the GPU arm, 20 on the CPU reference arm)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=sorted(CONFIGS))
    ap.add_argument("--shard", default="auto", choices=["auto", "views", "range"],
                    help="multi-GPU: independent views per GPU (default, cfg 1-4) or one view sharded by MeshletDraw range (cfg 5)")
    ap.add_argument("--merge", default="strips", choices=["strips", "allreduce"],
                    help="range sharding: screen-strip owners pull dirty tiles + all-gather the pyramid (default) or the round-1 u64 min all-reduce of the whole visbuffer")
    ap.add_argument("--positions", default="f32", choices=["f32", "i16"],
                    help="what the rasteriser reads: the expanded f32 Vertex records the reference uploads (default) or, for KHR_mesh_quantization "
                         "scenes (cfg 2, cfg 41), the accessor's own int16 data dequantised in registers (extension, bit-identical image)")
    ap.add_argument("--cone-cull", action="store_true", help="enable the optional normal-cone backface cull (extension, identical image)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-range-leg", action="store_true", help="multi-GPU default run: skip the cfg-5 range-sharded leg (`range_sharded` key)")
    ap.add_argument("--one-pass", action="store_true", help="reference one-pass mode instead of the two-pass extension")
    ap.add_argument("--meshlets", default="meshopt", choices=["meshopt", "morton"],
                    help="meshlet builder of the synthetic scene: the reference's partition (meshopt_buildMeshlets + optimizeMeshlet, default) "
                         "or the round-1 Morton packer (40 %% more, smaller meshlets)")
    args = ap.parse_args()
    from vk_gltf_viewer_b200.scene import select_builder
    select_builder(args.meshlets)
    builder_note = ("meshopt_buildMeshlets(64,124,0)+meshopt_optimizeMeshlet partition (assets.cpp:331-346), host/clusterizer.cpp" if args.meshlets == "meshopt"
                    else "Morton-order greedy packer (host/meshlet_builder.cpp)")

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    label, spec, (W, H) = CONFIGS[args.config]
    ncores = os.cpu_count() or 1

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return
        scene = build_scene(spec)
        cnt = scene.counts()
        views = [scene.default_view(i, NVIEWS) for i in range(NVIEWS)]
        steps = max(1, args.steps if args.steps is not None else 20)
        warm = max(1, min(args.warmup, 3))   # every warm-up is a full CPU frame (~0.5 s at cfg 3): at most 3, reported as run
        cpu_stage = {}
        times = cpu_frames(scene, W, H, views, steps, warmup=warm, stage_s=cpu_stage)
        total = sum(times)
        fps = steps / total
        line = {
            "impl": "reference", "metric": "frames/s (two-pass cull + HiZ + visbuffer)", "value": fps, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": 1e3 * total / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32+u64", "data": "synthetic",
            "gtris_per_s": cnt.triangles_instanced * fps / 1e9,
            "config": config_dict(label, W, H, cnt, 2, builder_note, max(1, args.gpus), "range" if (args.shard == "range" or (args.shard == "auto" and args.config == 5)) else "views"),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": ncores, "kind": "port",
                             "sample": f"{steps} full two-pass frames of the same camera sweep (CPU oracle, all host threads)"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "stages": {k: {"ms": round(1e3 * v / steps, 3)} for k, v in cpu_stage.items()},   # the same table as the GPU arm's `stages`, on the host cores
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm (CUDA)
    from vk_gltf_viewer_b200 import api, multigpu
    from vk_gltf_viewer_b200.scene import Camera

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    shard = args.shard if args.shard != "auto" else ("range" if args.config == 5 else "views")
    scene = build_scene(spec)
    cnt = scene.counts()
    views = [scene.default_view(i, NVIEWS) for i in range(NVIEWS)]
    steps, warm = max(1, args.steps if args.steps is not None else 200), max(3, args.warmup)
    if shard == "views":   # independent views per GPU, no data-path collective.  Every rank walks the same camera
        # sweep with the same step, phase-shifted by rank * NVIEWS / world (rank r starts at view r*64/n): frame-to-frame
        # coherence — what the previous-frame HiZ test feeds on — is then the same at every N.  (A round-robin i mod n
        # assignment makes each rank jump n views per frame and silently inflates pass B with N.)
        my_views = [views[(rank * NVIEWS // world + i) % NVIEWS] for i in range(warm + steps + 1)]
    else:                  # every rank renders the SAME views, each its own range of the draw list
        my_views = [views[i % NVIEWS] for i in range(warm + steps + 1)]

    r = api.Renderer(W, H, device=local_rank)
    pyr_floats = r.pyramid_floats
    r_layout = [(o, w, h) for (o, w, h) in r.layout]   # (offset, width, height) per mip
    cam = Camera(W, H)
    cam.look_at(*my_views[0])
    pc = r.upload_scene(scene, cam)
    flags = api.FRAME_ONE_PASS if args.one_pass else api.FRAME_TWO_PASS
    if args.positions == "i16":
        r.upload_quantized(scene)
    if args.cone_cull:
        r.upload_cones(scene)
        flags |= api.FRAME_CONE_CULL
    if shard == "range" and world > 1:
        r.set_shard_interleaved(rank, world, SHARD_BLOCK_LOG2)  # 2048-draw blocks round-robin: balances the surviving work (a contiguous half does not)
        multigpu.attach_peers(r, dist)
        flags |= api.FRAME_MERGE if args.merge == "allreduce" else api.FRAME_MERGE_STRIPS
    # all cameras of the sweep resident in HBM: the device-timed loop switches the camera ADDRESS per frame
    cam_addrs = []
    for v in my_views[1:]:
        cam.look_at(*v)
        cam_addrs.append(r.upload(np.frombuffer(cam.raw(), np.uint8)))

    def barrier():
        r.sync()
        if dist is not None:
            dist.barrier()
            import torch
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # warm-up (>= 3): also seeds the pyramid
    r.frame(pc, flags)
    for k in range(warm):
        pc.cameraBuffer = cam_addrs[k % len(cam_addrs)]
        r.frame(pc, flags)

    # ---- device-timed: K frames, each bracketed by CUDA events on the launching stream; L2 flushed between frames
    barrier()
    t_begin = time.time()
    per = []
    stage_names = ("clear_ms", "cull_a_ms", "raster_a_ms", "merge_a_ms", "hiz_a_ms", "cull_b_ms", "raster_b_ms", "merge_b_ms", "hiz_b_ms")
    stage = {k: 0.0 for k in stage_names}
    vis_a = vis_b = occ_a = 0
    hiz_tiles_b = 0
    launches = 0
    for k in range(steps):
        pc.cameraBuffer = cam_addrs[(warm + k) % len(cam_addrs)]
        r.flush_l2(256 << 20)
        st = r.frame(pc, flags | api.FRAME_TIMED)   # one event pair around the frame: nothing sits between its launches
        per.append(st.total_ms)
        vis_a += st.visible_a; vis_b += st.visible_b; occ_a += st.occluded_a
        hiz_tiles_b += st.hiz_tiles_b
        launches += st.kernel_launches + 1  # + the L2-flush fill kernel
    barrier()
    t_end = time.time()
    dev_ms = sum(per)
    # ---- per-stage breakdown: a second, shorter pass over the same sweep with an event after every stage (VKV_FRAME_STAGES);
    # those events keep the pass-B cull from starting under the pyramid's tail, so the stages add up to slightly more than a frame
    KS = min(steps, 50)
    for k in range(KS):
        pc.cameraBuffer = cam_addrs[(warm + k) % len(cam_addrs)]
        r.flush_l2(256 << 20)
        st = r.frame(pc, flags | api.FRAME_TIMED | api.FRAME_STAGES)
        for sname in stage:
            stage[sname] += getattr(st, sname)
    barrier()

    # ---- end to end through the C ABI with HOST buffers (pinned): per step H2D camera + transforms (the reference re-uploads
    # both every frame: camera.cpp:180-193, world.cpp:321-344) and D2H of the frame's counters (vkv_stats); `e2e_readback` adds the
    # D2H of the whole R32_UINT id image (vkv_read_ids), i.e. "result back on the host" for a CPU consumer
    import torch
    FO = 3  # frames in flight = the reference's Application::frameOverlap (application.hpp:146): per-slot camera / transform buffers, slot k
    # is waited for (vkv_frame_wait <-> the frame fence, application.cpp:642-660) just before it is reused
    transforms = np.ascontiguousarray(scene.transforms())
    pin_tr = torch.from_numpy(transforms.reshape(-1).copy()).pin_memory()
    pin_cam = [torch.zeros(352, dtype=torch.uint8).pin_memory() for _ in range(FO)]
    pin_ids = torch.zeros(W * H, dtype=torch.int32).pin_memory()
    cam_np = [t.numpy() for t in pin_cam]
    slot_cam = [r.upload(np.frombuffer(cam.raw(), np.uint8)) for _ in range(FO)]
    slot_tr = [pc.transformBuffer] + [r.upload(transforms.reshape(-1).view(np.uint8)) for _ in range(FO - 1)]
    h2d = 352 + transforms.nbytes
    d2h = 256
    e2e_visible = [0]

    def e2e_loop(n, readback):
        """every step: camera + node transforms copied from pinned host memory (vkv_update_staged), the frame, its counters copied back
        and consumed on the host (vkv_frame_wait) — FO frames in flight, as Application::run() keeps them"""
        barrier()
        t0 = time.perf_counter()
        tickets = [0] * FO
        for k in range(n):
            sl = k % FO
            if tickets[sl]:
                e2e_visible[0] += r.frame_wait(tickets[sl]).visible_a   # the slot's fence: its buffers and pinned staging are free again
            cam.look_at(*my_views[1 + (warm + k) % (len(my_views) - 1)])
            cam_np[sl][:] = np.frombuffer(cam.raw(), np.uint8)
            r.update_staged(slot_cam[sl], pin_cam[sl].data_ptr(), 352)
            r.update_staged(slot_tr[sl], pin_tr.data_ptr(), transforms.nbytes)
            pc.cameraBuffer, pc.transformBuffer = slot_cam[sl], slot_tr[sl]
            if readback:  # a CPU consumer of the id image cannot run ahead of it: blocking frame + blocking image copy
                r.frame(pc, flags)
                r._ck(r.L.vkv_read_ids(r.h, pin_ids.data_ptr()))
            else:
                tickets[sl] = r.frame_submit(pc, flags)
        for t in tickets:
            if t:
                e2e_visible[0] += r.frame_wait(t).visible_a
        barrier()
        return time.perf_counter() - t0

    def e2e_blocking_loop(n):
        """the same uploads and read-back with ONE frame in flight (vkv_update + blocking vkv_frame): round 1's e2e definition"""
        barrier()
        t0 = time.perf_counter()
        pc.cameraBuffer, pc.transformBuffer = slot_cam[0], slot_tr[0]
        for k in range(n):
            cam.look_at(*my_views[1 + (warm + k) % (len(my_views) - 1)])
            cam_np[0][:] = np.frombuffer(cam.raw(), np.uint8)
            r._ck(r.L.vkv_update(r.h, slot_cam[0], pin_cam[0].data_ptr(), 352))
            r._ck(r.L.vkv_update(r.h, slot_tr[0], pin_tr.data_ptr(), transforms.nbytes))
            r.frame(pc, flags)
        barrier()
        return time.perf_counter() - t0

    e2e_loop(min(steps, 8), False)  # warm the slots
    e2e_s = e2e_loop(steps, False)
    e2e_blk_s = e2e_blocking_loop(steps)
    KR = min(steps, 50)
    e2e_rb_s = e2e_loop(KR, True)

    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None

    # ---- N > 1: after the view-sharded measurement, the meshlet-range-sharded configuration (BASELINE config 5) with its exchange
    # step, parity-checked against one GPU — so that the driver's scaling record carries both multi-GPU paths
    range_leg = None
    if dist is not None and shard == "views" and not args.no_range_leg:
        r.close()
        try:
            range_leg = range_sharded_leg(rank, world, local_rank, dist, min(steps, 20), builder_note)
        except Exception as e:  # noqa: BLE001 — the headline line must still print; the failure is reported in it
            range_leg = {"error": repr(e), "merge_parity": False}
        r = None

    if dist is not None:
        import torch
        t = torch.tensor([dev_ms, e2e_s, e2e_rb_s, e2e_blk_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s, e2e_rb_s, e2e_blk_s = t.tolist()
        ln = torch.tensor([launches, vis_a, vis_b, occ_a], device="cuda", dtype=torch.int64)
        dist.all_reduce(ln)
        launches = int(ln[0].item())
        if shard == "range":
            vis_a, vis_b, occ_a = int(ln[1].item()), int(ln[2].item()), int(ln[3].item())

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        K = steps
        frames_total = world * K if shard == "views" else K
        fps = frames_total / (dev_ms / 1e3)
        # algorithmic bytes per launch (SURVEY §8d definitions), per GPU
        per_gpu = 1.0 / world if shard == "range" else 1.0
        N = cnt.draws * per_gpu
        U = cnt.meshlets_unique * 36 + cnt.transforms * 64 + cnt.primitives * 64 + 352
        pyr_bytes = 4 * pyr_floats
        avg = lambda x: x / K * per_gpu
        bytes_cull_a = 12 * N + U + 4 * (avg(vis_a) + avg(occ_a))
        bytes_cull_b = 4 * avg(occ_a) + 12 * avg(occ_a) + U + 4 * avg(vis_b)
        bytes_hiz = 8 * W * H + pyr_bytes  # depth is read fused from the 64-bit visbuffer: 8 B/pixel, not 4
        # the second build of a two-pass frame reduces only the 64x16-pixel tiles a small pass B drew into (vkv_stats.hiz_tiles_b, counted on
        # the device): 8 KB read + 340 exact-mip texels written per tile, plus the small mips (always rebuilt: they hang off every tile)
        tiles_all = ((W + 63) // 64) * ((H + 15) // 16)
        tiles_b = hiz_tiles_b / K if hiz_tiles_b else tiles_all
        frac_b = min(1.0, tiles_b / tiles_all)
        bytes_hiz_b = frac_b * (8 * W * H + 4 * sum(w * h for (_, w, h) in r_layout[:4])) + 4 * sum(w * h for (_, w, h) in r_layout[4:])
        bytes_clear = 8 * W * H
        avg_v, avg_t = meshlet_averages(scene)
        per_meshlet = 48 + 64 + 28 * avg_v + 3 * avg_t   # headers + transform + (4 B index + 24 B vertex stride) per vertex + 3 B per triangle
        clear_fused = os.environ.get("VKV_SEPARATE_CLEAR", "0") != "1" and cnt.draws > 0 and (W * H) % 2 == 0  # vkv_frame: the clear rides inside the pass-A cull launch (cull.cu), its bytes are that launch's
        if clear_fused:
            bytes_cull_a += bytes_clear
            bytes_clear = 0
        bytes_merge = 8 * W * H * 2 if (flags & api.FRAME_MERGE) else 0  # all-reduce merge, per GPU: strip read from n ranks + written to n ranks = 2 * 8WH
        if flags & api.FRAME_MERGE_STRIPS:  # strip mode: the owner reads its own strip once (+ the dirty peer tiles, reported separately)
            bytes_merge = 8 * W * H / world
        stages = {}
        for name, b, ms in (("clear", bytes_clear, stage["clear_ms"]), ("cull_a", bytes_cull_a, stage["cull_a_ms"]),
                            ("raster_a", avg(vis_a) * per_meshlet, stage["raster_a_ms"]), ("merge_a", bytes_merge, stage["merge_a_ms"]),
                            ("hiz_a", bytes_hiz, stage["hiz_a_ms"]), ("cull_b", bytes_cull_b, stage["cull_b_ms"]),
                            ("raster_b", avg(vis_b) * per_meshlet, stage["raster_b_ms"]), ("merge_b", bytes_merge, stage["merge_b_ms"]),
                            ("hiz_b", bytes_hiz_b, stage["hiz_b_ms"])):
            m = ms / KS
            if (m <= 0 and b == 0) or (name == "clear" and clear_fused):
                continue
            gbs = (b / 1e9) / (m / 1e3) if m > 0 else 0.0
            stages[name] = {"ms": round(m, 5), "bytes": int(b), "GB/s": round(gbs, 1), "frac_hbm": round(gbs / hbm, 4)}
        dom = max(stages, key=lambda sname: stages[sname]["ms"])
        traffic = None
        issue_of = None
        try:  # dram bytes and warp-instructions per launch of the dominant kernel from the committed `ncu --set full` capture of this workload
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            ent = tr.get(f"cfg{args.config}" + ("" if args.meshlets == "meshopt" else "_morton"), {}).get(dom)
            if ent and world == 1:
                traffic = ent["dram_read_bytes"] + ent["dram_write_bytes"]
                issue_of = ent
        except Exception:
            pass
        line = {
            "metric": "frames/s (two-pass cull + HiZ + visbuffer)" if not args.one_pass else "frames/s (one-pass cull + visbuffer + HiZ)",
            "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": warm,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak" if shard == "views" else "strong", "vs_baseline": None,
            "dtype": "f32+u64", "data": "synthetic",
            "gtris_per_s": cnt.triangles_instanced * fps / 1e9,
            "config": config_dict(label, W, H, cnt, 1 if args.one_pass else 2, builder_note, world, shard),
            "run": {"positions": args.positions, "cone_cull": bool(args.cone_cull), "stages_from": f"a second pass of {KS} frames with an event after every stage (the frame time above has none inside the frame)",
                    "visible_a_avg": vis_a / K, "occluded_a_avg": occ_a / K, "visible_b_avg": vis_b / K,
                    "hiz_b_tiles_avg": round(tiles_b, 1), "hiz_tiles": tiles_all,
                    "avg_vertices_per_meshlet": round(avg_v, 2), "avg_triangles_per_meshlet": round(avg_t, 2),
                    "clear": "fused into the pass-A cull launch (its 8*W*H bytes are counted there)" if clear_fused else "separate launch"},
            "e2e": {"value": frames_total / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "frames_in_flight": FO, "one_frame_in_flight": frames_total / e2e_blk_s,
                    "note": "wall clock, no L2 flush; every frame: camera + all node transforms copied from pinned host memory (vkv_update_staged), frame "
                            "(vkv_frame_submit), its counters copied back and read on the host (vkv_frame_wait); frames in flight = the reference's frameOverlap "
                            "(application.hpp:146). one_frame_in_flight = the same with vkv_update + blocking vkv_frame"},
            "e2e_readback": {"value": (world * KR if shard == "views" else KR) / e2e_rb_s, "unit": "frames/s", "steps": KR, "h2d_bytes_per_step": h2d,
                             "d2h_bytes_per_step": d2h + 4 * W * H,
                             "note": "as e2e, plus the whole R32_UINT id image copied to pinned host memory every frame (vkv_read_ids): PCIe-bound"},
            "gpu_launches": launches,
            "roofline": roofline(dom, stages, issue_of, hbm, peak_src, traffic, clocks),
            "stages": stages,
            "clocks": clocks,
        }
        if range_leg is not None:
            line["range_sharded"] = range_leg
        if not args.no_cpu_baseline:
            nfr = 2 if args.config in (3, 5) else 5
            cpu_stage = {}
            ct = cpu_frames(scene, W, H, views, nfr, stage_s=cpu_stage)
            cfps = nfr / sum(ct)
            line["cpu_baseline"] = {"value": cfps, "unit": "frames/s", "cores": ncores, "kind": "port",
                                    "stages_ms": {k: round(1e3 * v / nfr, 3) for k, v in cpu_stage.items()},
                                    "sample": f"{nfr} full two-pass frames of the same sweep after 1 warm-up frame (CPU oracle, all host threads)"}
        print(json.dumps(line))
    if r is not None:
        r.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
