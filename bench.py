#!/usr/bin/env python
"""bench.py -- particle-steps/s of the Tendrils step (integrate + flow splat + respawn) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3|cfg2|cfg1|cfg5]

N > 1 is launched by torchrun (one rank per GPU); rank 0 prints ONE JSON line.

Workload (BASELINE.json): at N = 1 `configs[2]` -- 4096x4096 particles, 1024x1024 flow grid with
flowDecay + trail splat, image-pixel (best-sample) respawn every 60th step -- which is the
configuration the metric's target (">= 70 % of HBM roofline ... at 16M particles on 1 B200") is
quoted on; at N > 1 `configs[3]`, the same per GPU (weak scaling, global texture 4096 x 4096N).

A "step" = Tendrils.step() + the flow half of Tendrils.draw() (+ the respawn pass when due).
`value` counts inputs resident in HBM; `e2e` re-measures with the particle state crossing
PCIe in both directions every step.  `--impl reference` times the CPU oracle port (the
reference itself -- WebGL shaders -- cannot run here).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: particle texture root, flow grid, respawn
    "cfg1": dict(R=512, G=256, respawn="ball", every=0, desc="512x512 particles, 256^2 flow grid, ball spawn"),
    "cfg2": dict(R=2048, G=512, respawn="ball", every=0, desc="2048x2048 particles, 512^2 flow grid, defaults, noise on"),
    "cfg3": dict(R=4096, G=1024, respawn="best-sample", every=60,
                 desc="4096x4096 particles per GPU, 1024^2 flow grid, flowDecay + trail splat, "
                      "direct image-pixel spawn (spawnImage) at step 0, best-sample image respawn (spawnSamples) every 60th step"),
    # strong scaling: the texture is 4096 x 32768 (= 8 shards of 4096^2) whatever the rank count
    "cfg5": dict(R=4096, rows=32768, G=2048, respawn="best-sample", every=30, optical=True,
                 desc="2^27 particles in total (4096 x 32768 texture, column shards), 2048^2 flow grid, every step the optical "
                      "flow of a synthetic two-frame video pair (moving gaussian blobs) is drawn into the flow grid, "
                      "best-sample respawn from the current frame every 30th step"),
}
METRIC = "particle_steps_per_sec"
OPTICAL = {"speed": 0.08, "offset": 0.1, "scaleUV": [-1, -1]}      # as the parity test drives src/optical-flow
UNIT = "particle-steps/s"


def synthetic_image(n):
    from util import synthetic_image as mk
    return mk(n, n)


def synthetic_frames(n, count=8):
    from util import synthetic_video as mk
    return mk(n, n, count)


def algorithmic_bytes(n_particles, grid, optical=False):
    """state read+write 32 B/particle; flow grid gathered once 16 B/texel;
    flow update read+write 32 B/texel.  The optical-flow pass adds two RGBA8 frames read (8 B/texel)
    and one more read+write of the grid (32 B/texel)."""
    extra = 40 * grid if optical else 0
    return {"integrate": 32 * n_particles + 16 * grid, "splat": 32 * grid, "step": 32 * n_particles + 48 * grid + extra}


def config_of(name, wl, world, n_local, n_total):
    """The `config` object, identical in both arms (the driver compares them)."""
    return {"workload": f"{name}: {wl['desc']}", "particles_per_gpu": n_local, "particles_total": n_total,
            "flow_grid": [wl["G"], wl["G"]], "state": "reference defaults (src/index.js:29-57), noise on",
            "splat": "exact ordered alpha-over (reference semantics)",
            "l2": "inputs larger than L2 (state 2 x %d MiB per GPU)" % (n_local * 16 >> 20),
            "parallelism": (f"particle columns sharded over {world} GPU(s), flow blend shared over peer memory (bins owned round-robin)"
                            if world > 1 else "1 GPU")}


def host_info():
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    model = ""
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                model = ln.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return cores, model


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # under load = the upper half of the samples (idle samples at the edges pull the median down)
        sm_sorted = sorted(sm)
        load = sm_sorted[len(sm_sorted) // 2:]
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def build_sim(wl, rank, world, local_rank, group):
    import tendrils_b200 as T
    from tendrils_b200.spawn import PixelSpawner, spawnBall
    from tendrils_b200.spawn.pixels import bestSampleFrag, mat3_identity, mat3_scale, pixelsFrag
    R, G = wl["R"], wl["G"]
    t = T.Tendrils(T.Device(G, G, device=local_rank, rank=rank, world_size=world, group=group))
    # N ranks: one R x (R*N) texture sharded by columns (R/N columns x R*N rows = R^2 particles per rank).
    # Widening instead (R*N x R) would leave the exact 1:1 vertex->column range of the reference's LUT.
    t.setup([R, wl.get("rows") or R * world])
    t.resize()
    first = spawnBall(t.gl, {"uniforms": {"radius": 0.3, "speed": 0.005}})  # src/demo.main.js:1402-1405
    sp = None
    if wl["respawn"] == "best-sample":
        img = synthetic_image(G)
        flip = mat3_scale(mat3_identity(), [-1, 1])                       # src/demo.main.js:462-463
        # spawnImage: direct pixel spawn, speed 0.3 (src/demo.main.js:511-512); spawnSamples: best-sample, speed 1 (:514-515)
        first = PixelSpawner(t.gl, {"shader": pixelsFrag, "buffer": img, "speed": 0.3, "jitterRad": 2, "spawnSize": [1, 1]})
        sp = PixelSpawner(t.gl, {"shader": bestSampleFrag, "buffer": img, "speed": 1, "bias": 1, "jitterRad": 2,
                                 "spawnSize": [1, 1]})
        first.spawnMatrix = sp.spawnMatrix = flip
    if wl.get("optical"):
        # cfg5: the video frames live on the device (a decoder would put them there); every step the pair
        # (frame k, frame k-1) drives the optical-flow pass, every respawn samples the current frame
        import torch
        from tendrils_b200.optical_flow import OpticalFlow
        dev = torch.device("cuda", local_rank)
        t.video = [torch.as_tensor(f, device=dev) for f in synthetic_frames(G)]
        t.optical = OpticalFlow(t.gl, None, dict(OPTICAL))
        t.optical.resize([G, G])
        first.setPixels(t.video[0].float() / 255.0)
    return t, first, sp


def sharded_parity_check(rank, world, local_rank, group):
    """N > 1, before anything is timed: a small TALL texture (64 columns x 64*N rows, the shape of the weak-scaling workload)
    stepped through the default transport must equal the CPU oracle bit for bit -- this rank's shard of the state and the
    whole flow grid.  Returns True only if every rank agrees."""
    import torch
    import torch.distributed as dist
    import tendrils_b200 as T
    from oracle import oracle as O
    from tendrils_b200.spawn import spawnBall
    from util import bits_equal
    PW, PH, G, steps = 64, 64 * world, 128, 6
    t = T.Tendrils(T.Device(G, G, device=local_rank, rank=rank, world_size=world, group=group))
    t.setup([PW, PH]); t.resize()
    spawnBall(t.gl, {"uniforms": {"radius": 0.3, "speed": 0.005}}).spawn(t)
    P = O.make_params()
    cur, prev = O.spawn_ball(PW, PH, 0.3, 0.005), O.spawn_init(PW, PH)
    targets, flow = np.zeros((PW, PH, 4), np.float32), np.zeros((G, G, 4), np.float32)
    for _ in range(steps):
        t.timer.tick()
        t.step().draw()
        new = O.integrate(P, cur, targets, flow, np.float32(t.timer.time), np.float32(t.timer.dt))
        prev, cur = cur, new
        O.splat(P, cur, prev, flow, np.float32(t.timer.time))
    c0, c1 = t.particles.col0, t.particles.col1
    ok = bool(bits_equal(t.particles.buffers[0].download(), cur[c0:c1]).all() and bits_equal(t.flow.download(), flow).all())
    t.dispose()
    flag = torch.tensor([1 if ok else 0], device="cuda", dtype=torch.int32)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(flag.item())


def grids_identical(P, world):
    """after the timed region every rank must hold the same flow grid: compare a checksum of the raw bits"""
    import torch
    import torch.distributed as dist
    bits = P._flow_tensor().view(torch.int32).to(torch.int64)
    w = torch.arange(1, bits.numel() + 1, device=bits.device, dtype=torch.int64) % 65521
    cs = torch.stack([bits.sum(), (bits * w).sum()])
    lo, hi = cs.clone(), cs.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return bool((lo == hi).all().item())


def run_ours(args, wl, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        group = dist.group.WORLD
    torch.cuda.set_device(local_rank)
    parity = None
    if world > 1:
        parity = sharded_parity_check(rank, world, local_rank, group)
        if not parity:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "error": "multi-GPU parity check failed", "multi_gpu_parity": False}), flush=True)
            dist.destroy_process_group()
            sys.exit(3)
    t, first, sp = build_sim(wl, rank, world, local_rank, group)
    P = t.particles
    n_local = (P.col1 - P.col0) * P.shape[1]
    n_total = P.shape[0] * P.shape[1]
    grid = wl["G"] * wl["G"]
    stream = torch.cuda.ExternalStream(P.stream_handle(), device=torch.device("cuda", local_rank))
    counter = {"k": 0}

    def one_step():
        k = counter["k"]
        if k == 0:
            first.spawn(t)                    # spawnShader ticks the timer itself (src/index.js:433)
        elif wl["every"] and k % wl["every"] == 0:
            if wl.get("optical"):
                sp.setPixels(t.video[k % len(t.video)].float() / 255.0)
            sp.spawn(t)
        t.timer.tick()
        t.step().draw()
        if wl.get("optical"):                  # drawn into the flow FBO after the particles (src/demo.main.js:1131-1159)
            of = t.optical
            of.setPixels(t.video[k % len(t.video)])
            of.update({"speedLimit": t.state["speedLimit"], "time": t.timer.time, "viewSize": t.viewSize}).render(t)
            of.step()
        counter["k"] = k + 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_step()
    barrier()
    P.timing(reset=True)
    launches0 = P.stats()["kernel_launches"]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        one_step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = P.stats()["kernel_launches"] - launches0
    frags = P.stats()["last_fragments"]
    tm = P.timing()
    # the dominant kernel alone: a few more steps with the noise/splat overlap off (one fused k_integrate launch per step, not
    # time-sliced under the splat), timed by the library's own CUDA events on its stream
    P.set_overlap(False)
    P.timing(reset=True)
    for _ in range(min(args.steps, 10)):
        one_step()
    barrier()
    tm_iso = P.timing()
    P.set_overlap(True)
    same_grids = None
    if world > 1:
        tt = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
        ft = torch.tensor([frags], device="cuda", dtype=torch.int64)
        dist.all_reduce(ft, op=dist.ReduceOp.SUM)
        frags = int(ft.item())
        same_grids = grids_identical(P, world)
        if not same_grids:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "error": "the ranks' flow grids differ after the timed region"}), flush=True)
            dist.destroy_process_group()
            sys.exit(4)

    # ---- e2e: the same steps with the particle state crossing PCIe both ways every step -------
    # The public call for callers that keep the state on the host: Tendrils.stepStreamed(host_in, host_out) uploads, steps and
    # reads back in column chunks (tb_step_streamed); the state round-trips through ONE pinned buffer, every step.
    e2e_steps = max(1, min(args.steps, 12))
    host = torch.empty((P.col1 - P.col0, P.shape[1], 4), dtype=torch.float32, pin_memory=True)
    hview = host.numpy()
    P.buffers[0].download(out=hview)

    def e2e_step():
        k = counter["k"]
        if wl["every"] and k % wl["every"] == 0 and k > 0:
            P.sync()                               # a respawn works on the device state: bring the host copy in first
            P.buffers[0].upload(hview)
            if wl.get("optical"):
                sp.setPixels(t.video[k % len(t.video)].float() / 255.0)
            sp.spawn(t)
            P.buffers[0].download(out=hview)
        t.timer.tick()
        t.stepStreamed(hview, hview).draw()
        if wl.get("optical"):
            of = t.optical
            of.setPixels(t.video[k % len(t.video)])
            of.update({"speedLimit": t.state["speedLimit"], "time": t.timer.time, "viewSize": t.viewSize}).render(t)
            of.step()
        counter["k"] = k + 1

    for _ in range(2):
        e2e_step()
    P.sync()
    barrier()
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    P.sync()                                       # the last step's state is back in host memory
    barrier()
    e2e_s = time.perf_counter() - w0
    if world > 1:
        tt = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return None

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    ab = algorithmic_bytes(n_local, grid, bool(wl.get("optical")))
    traffic = {}
    try:        # DRAM bytes per launch from the committed ncu --set full captures of this workload (profiles/traffic.json)
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload, {})
    except Exception:
        pass
    fin_us = 1e3 * tm["integrate_ms"] / max(tm["n_integrate"], 1)      # main-stream launch (fused, or the finish half)
    noise_us = 1e3 * tm["noise_ms"] / max(tm["n_integrate"], 1)        # side-stream noise launch, time-sliced under the splat
    spl_us = 1e3 * tm["splat_ms"] / max(tm["n_splat"], 1)
    int_us = 1e3 * tm_iso["integrate_ms"] / max(tm_iso["n_integrate"], 1)   # the fused launch on its own
    spl_iso_us = 1e3 * tm_iso["splat_ms"] / max(tm_iso["n_splat"], 1)
    dom_bytes, dom_us = ab["integrate"], int_us
    achieved = dom_bytes / (dom_us * 1e-6) / 1e9
    splat_gbs = ab["splat"] / (spl_iso_us * 1e-6) / 1e9 if spl_iso_us > 0 else 0.0
    step_gbs = ab["step"] / (ms / args.steps * 1e-3) / 1e9
    cfg = config_of(args.workload, wl, world, n_local, n_total)
    cfg["fragments_last_step"] = frags
    out = {
        "metric": METRIC, "value": n_total * args.steps / (ms * 1e-3), "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "strong" if wl.get("rows") else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": cfg,
        "clocks": clocks,
        "e2e": {"value": n_total * e2e_steps / e2e_s, "unit": UNIT, "steps": e2e_steps,
                "h2d_bytes_per_step": n_local * 16, "d2h_bytes_per_step": n_local * 16,
                "api": "Tendrils.stepStreamed(host, host).draw(): tb_step_streamed, 16 column chunks, one pinned buffer round trip"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "k_integrate (logic.frag)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic.get("k_integrate"), "peak_source": peak_kind,
                     "algorithmic_bytes_per_launch": dom_bytes, "avg_launch_us": dom_us,
                     "integrate_us": int_us, "splat_us": spl_iso_us,
                     "timed_region": {"overlap": "noise of step k+1 on a low-priority side stream under the splat of step k",
                                      "integrate_finish_main_stream_us": fin_us, "integrate_noise_side_stream_us": noise_us,
                                      "splat_us": spl_us},
                     "note": "k_integrate is FP32-issue bound (2 simplex noises per particle), not HBM bound; see DESIGN.md section 3",
                     "issue_active": traffic.get("issue_active"),
                     "splat_pipeline": {"kernels": "k_splat_hist + k_splat_rows + k_splat_plan + k_splat_scatter + k_splat_fold",
                                        "algorithmic_bytes": ab["splat"], "avg_us": spl_iso_us, "achieved": splat_gbs, "frac": splat_gbs / peak,
                                        "traffic": traffic.get("splat_pipeline"),
                                        "note": "algorithmic bytes = flow grid read + write only; the fragments the ordered blend has to "
                                                "materialise (16 B each, written once, read once) are implementation traffic"},
                     "whole_step": {"algorithmic_bytes": ab["step"], "achieved": step_gbs, "frac": step_gbs / peak}},
    }
    if world > 1:
        out["multi_gpu_parity"] = bool(parity)
        out["grids_identical_after_timed_region"] = bool(same_grids)
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_arm(args, wl, budget_s=args.cpu_budget, as_reference=False, world=1)
    if world > 1:
        dist.destroy_process_group()
    return out


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores (cpu_baseline and --impl reference)
# ------------------------------------------------------------------------------------------------
def cpu_arm(args, wl, budget_s, as_reference, world=1):
    from oracle import oracle as O
    O.build()
    cores, cpu_model = host_info()
    O.set_threads(cores)                     # every host core, whatever OMP_NUM_THREADS the launcher exported (torchrun: 1)
    R, G = wl["R"], wl["G"]
    PH = wl.get("rows") or R * max(world, 1)
    cols = (0, max(min(R // 4, (1 << 22) // PH), 1))   # the bounded sample: the first columns, at most a quarter / 4M particles
    n_sample = (cols[1] - cols[0]) * PH
    Pm = O.make_params()
    S = O.make_spawn_pixels(jitter=(np.float32(np.float32(1.0 / G) * 2), np.float32(np.float32(1.0 / G) * 2)),
                            spawnMatrix=(-1, 0, 0, 0, 1, 0, 0, 0, 1))
    img = synthetic_image(G)
    state = {"cur": O.spawn_init(R, PH), "prev": O.spawn_init(R, PH), "k": 0, "time": 0.0}
    targets = np.zeros((R, PH, 4), np.float32)
    flow = np.zeros((G, G, 4), np.float32)
    dt = 1000 / 60
    video = synthetic_frames(G) if wl.get("optical") else None
    last_frame = np.zeros((G, G, 4), np.uint8)
    if video:
        img = video[0].astype(np.float32) / np.float32(255.0)

    def one_step():
        k = state["k"]
        if k == 0 or (wl["every"] and k % wl["every"] == 0):
            state["time"] += dt
            if k == 0 and wl["respawn"] == "ball":
                new = O.spawn_ball(R, PH, 0.3, 0.005, cols=cols)
            elif k == 0:
                S.speed = 0.3
                new = O.spawn_pixels_direct(S, R, PH, img, state["time"], cols=cols)
                S.speed = 1.0
            else:
                src = video[k % len(video)].astype(np.float32) / np.float32(255.0) if video else img
                new = O.spawn_pixels_sample(S, "best", state["cur"], src, state["time"], cols=cols)
            state["prev"], state["cur"] = state["cur"], new
        state["time"] += dt
        new = O.integrate(Pm, state["cur"], targets, flow, state["time"], dt, cols=cols)
        state["prev"], state["cur"] = state["cur"], new
        O.splat(Pm, state["cur"], state["prev"], flow, state["time"], cols=cols, mt=True)
        if video:
            nonlocal last_frame
            frame = video[k % len(video)]
            O.optical_flow(flow, frame, last_frame, viewSize=(1.0, 1.0), scaleUV=tuple(OPTICAL["scaleUV"]), offset=OPTICAL["offset"],
                           lambda_=0.001, speed=OPTICAL["speed"], speedLimit=Pm.speedLimit, time=np.float32(state["time"]))
            last_frame = frame
        state["k"] = k + 1

    if as_reference:
        warm, steps = args.warmup, args.steps
    else:
        one_step()                                   # calibrate the bounded sample
        t0 = time.perf_counter(); one_step(); per = time.perf_counter() - t0
        warm, steps = 0, int(min(max(budget_s / max(per, 1e-3), 3), 200))
    for _ in range(warm):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    el = time.perf_counter() - t0
    return {"value": n_sample * steps / el, "unit": UNIT, "cores": O.num_threads(), "kind": "port", "nproc": cores, "cpu": cpu_model,
            "sample": f"columns [{cols[0]},{cols[1]}) of the {R}x{PH} particle texture ({n_sample} particles) on the full "
                      f"{G}^2 flow grid, {steps} steps of integrate + ordered splat (+ respawn / optical flow when due), OpenMP over "
                      f"{O.num_threads()} threads; oracle/tendrils_oracle.c",
            "steps": steps, "seconds": el, "ms_per_step": 1e3 * el / steps}


def run_reference(args, wl, rank, world):
    if rank != 0:
        return None
    cb = cpu_arm(args, wl, budget_s=0, as_reference=True, world=world)
    R = wl["R"]
    n_total = R * (wl.get("rows") or R * world)
    cfg = config_of(args.workload, wl, world, n_total // world, n_total)
    cfg["note"] = "CPU oracle port of the reference shaders (the WebGL reference cannot run here); host cores only"
    return {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
        "scaling": "strong" if wl.get("rows") else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=120)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS) + ["all"],
                    help="'all': cfg1, cfg2 and cfg3 in turn, one JSON line each (1 GPU)")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            sys.exit("bench.py: --gpus N > 1 must be launched with torchrun (one rank per GPU)")
    if args.workload == "all":
        for name in ("cfg1", "cfg2", "cfg3"):
            subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", name, "--steps", str(args.steps), "--warmup", str(args.warmup),
                            "--impl", args.impl] + (["--no-cpu-baseline"] if args.no_cpu_baseline else []), check=False)
        return
    wl = WORKLOADS[args.workload]
    # stdout carries exactly ONE JSON line: everything else a library prints there (e.g. NCCL's version
    # banner) is sent to stderr while the benchmark runs
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    out = run_reference(args, wl, rank, world) if args.impl == "reference" else run_ours(args, wl, rank, world, local_rank)
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    if out is not None:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
