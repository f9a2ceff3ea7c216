#!/usr/bin/env python
"""Benchmark of the draw/composite hot path (contract: task statement §④, DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl product|reference] [--workload c2|c1|c3|c4|c5|bg|c2x|c3p] [--present u8|yuv420p]

A *step* is one frame of the workload: the whole recorded command stream of one canvas is binned and
composited (ncr_bin_coarse -> ncr_bin_fine -> ncr_composite with the fused u8 image).

  value      frames/s with the frame's commands and textures already resident in HBM: the last flushed batch is
             re-executed K times from its device buffers (NcrRerunLastFlush), timed with CUDA events on the
             context's stream, L2 scrubbed (256 MB memset) before every step.  N ranks render independent
             frames (frame sharding, no collective): value = N*K / max-over-ranks(time).
  e2e        the same frames through the reference-facing C ABI from HOST buffers: the recorded stream is replayed
             call by call (csrc/ncr_replay.cpp -> CreateRenderContext/Translate/.../DrawTexture/GetBufferAsUInt8
             of the product library), which includes the host state machine, the H2D copy of the command batch
             from pinned staging and the D2H readback of the RGBA8 frame; wall clock, T host threads with one
             context each (the reference API has no globals, so this is legal for it too).
  roofline   ncr_composite: algorithmic bytes per launch / its mean CUDA-event duration, vs MEASURED_PEAKS.json.
  cpu_baseline / --impl reference   the unmodified reference build (oracle/_ref) — or the C restatement when the
             reference could not be built — on the host cores, one context per thread, on a bounded sample of the
             same stream.  This is the only place bench.py executes anything under oracle/.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (width, height, alpha, description)
    "c1": (1920, 1080, True, "BASELINE config 1: 1920x1080 RGBA, image.png as 1,000 affine alpha-blended quads (seed 0)"),
    "c2": (1920, 1080, True, "BASELINE config 2 (reference-ABI subset, SURVEY C2): 1920x1080 RGBA, 20,000 mixed draws "
                             "(60% DrawTexture, 20% DrawSplittedTexture, 10% DrawRect, 5% DrawVerticalGrd, 3% DrawCircle, "
                             "2% DrawLine; rotated/scaled, per-draw alpha), nearest sampling, seed 2"),
    "c3": (3840, 2160, True, "BASELINE config 3 (affine variant, SURVEY C3): 3840x2160 RGBA, 50,000 atlas sprites via "
                             "DrawSplittedTexture from a 2048^2 8x8-cell atlas, seed 3"),
    "c2x": (1920, 1080, True, "BASELINE config 2 with the extensions it names (PRODUCT ONLY, parity unpinned): the C2 stream with bilinear "
                             "sampling on half of the textured draws, N-gon fills instead of rects, and clip rects toggled every 500 draws"),
    "c3p": (3840, 2160, True, "BASELINE config 3, perspective variant (PRODUCT ONLY, parity unpinned): 3840x2160 RGBA, 50,000 "
                              "perspective-warped sprites from a 2048^2 atlas via NcrDrawTexturePerspective"),
    "bg": (1920, 1080, True, "low-overdraw end of the path: 1920x1080 RGBA, SetColor + full-screen DrawTexture (identity path) + FillColor "
                            "dim, u8 readback (3 commands per tile; the HBM-leaning regime)"),
    "c5": (3840, 2160, False, "BASELINE config 5 frame (SURVEY C5): 3840x2160 RGB milrenderer-shaped chart frame (C4's generator at 4K), "
                              "~1,500 notes, 12 lines, 100 hit effects, u8 readback; frames are sharded across ranks"),
    "c4": (1920, 1080, False, "BASELINE config 4 frame (SURVEY C4): 1920x1080 RGB milrenderer-shaped chart frame, ~1,500 notes, "
                              "12 lines, 100 hit effects, u8 readback"),
}


def build_workload(name: str, n_draws: int | None = None):
    """Returns (width, height, alpha, texture arrays, trace bytes as aligned uint8 array, draws, full draw count)."""
    from libnativecpurenderer_b200 import streams, trace

    w, h, alpha, _ = WORKLOADS[name]
    rec = trace.TraceRecorder(w, h, alpha)
    if name == "c1":
        tex_np = [np.load(os.path.join(ROOT, "tests", "golden", "image_rgba.npz"))["rgba"]]
        full = 1000
        streams.stream_k1(rec, trace.TexSlot(0, 128, 128), n=n_draws or full)
    elif name == "c2":
        tex_np = streams.make_c2_textures()
        full = 20000
        streams.stream_c2(rec, [trace.TexSlot(k, t.shape[1], t.shape[0]) for k, t in enumerate(tex_np)], n=n_draws or full)
    elif name == "c3":
        tex_np = [streams.make_atlas()]
        full = 50000
        streams.stream_c3(rec, trace.TexSlot(0, 2048, 2048), n=n_draws or full)
    elif name == "c2x":
        tex_np = streams.make_c2_textures()
        full = 20000
        streams.stream_c2x(rec, [trace.TexSlot(k, t.shape[1], t.shape[0]) for k, t in enumerate(tex_np)], n=n_draws or full)
    elif name == "c3p":
        tex_np = [streams.make_atlas()]
        full = 50000
        streams.stream_c3p(rec, trace.TexSlot(0, 2048, 2048), n=n_draws or full)
    elif name == "bg":
        tex_np = [np.ascontiguousarray(np.resize(streams.make_noise_texture(256, 7), (h, w, 4)))]
        full = 2
        slot = trace.TexSlot(0, w, h)
        rec.set_color(0, 0, 0, 1)
        rec.draw_texture(slot, 0, 0, w, h)
        rec.fill_color(0, 0, 0, .6)
    elif name in ("c4", "c5"):
        chart = streams.make_chart_textures()
        bg = np.ascontiguousarray(np.resize(streams.make_noise_texture(256, 7), (h, w, 4)))
        tex_np = [bg] + chart
        slots = [trace.TexSlot(k, t.shape[1], t.shape[0]) for k, t in enumerate(tex_np)]
        full = 1500
        streams.stream_c4_frame(rec, slots[0], slots[1:], frame=123, n_notes=n_draws or full)
    else:
        raise SystemExit(f"unknown workload {name}")
    rec.present()
    return w, h, alpha, tex_np, rec.as_array(), rec.n_draws, full


# ------------------------------------------------------------------------------------------------ helpers
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.path = None

    def __enter__(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.path or not os.path.exists(self.path):
            return out
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def measured_peak_gbs() -> tuple[float, str]:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except (KeyError, ValueError):
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def oracle_library() -> tuple[str, str]:
    ref = os.path.join(ROOT, "oracle", "_ref", "libNativeCPURenderer.so")
    if os.path.exists(ref):
        return ref, "reference"
    return os.path.join(ROOT, "oracle", "libncr_oracle.so"), "port"


def cpu_arm(workload: str, threads: int, sample_draws: int, steps: int, warmup: int):
    """Times the CPU implementation (all `threads` host threads, one context + one frame sample each per step)."""
    from libnativecpurenderer_b200 import trace
    from libnativecpurenderer_b200.binding import Renderer

    lib, kind = oracle_library()
    w, h, alpha, tex_np, arr, draws, full = build_workload(workload, sample_draws)
    R = Renderer(lib)
    tex = [R.Texture.from_numpy(t) for t in tex_np]
    rp = trace.Replayer(os.path.join(ROOT, "libnativecpurenderer_b200", "lib", "libncr_replay.so"), lib)
    frac = min(1.0, sample_draws / full)
    for _ in range(warmup):
        rp.run_threads(threads, w, h, alpha, arr, tex, repeats=1)
    secs = [rp.run_threads(threads, w, h, alpha, arr, tex, repeats=1) for _ in range(steps)]
    total = sum(secs)
    fps = threads * frac * steps / total
    sample = (f"{threads} threads x {steps} step(s), each thread renders the first {sample_draws} of {full} draws of one frame "
              f"({frac:.3f} frame); frames/s = threads*fraction*steps/wall")
    return fps, kind, sample, total / steps


# ------------------------------------------------------------------------------------------------ arms
def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = min(host_threads(), 64)
    _, _, _, _, _, _, full = build_workload(args.workload, 8)
    # one full C2 frame is ~16 s of one core; size the per-step sample so K+W steps end within ~2.5 minutes
    if args.workload in ("c2x", "c3p"):
        print(json.dumps({"impl": "reference", "unavailable": "the reference has no clip/bilinear/polygon/perspective entry points"}))
        return
    per_frame_s = {"c1": 1.1, "c2": 16.0, "c3": 60.0, "c4": 3.0, "c5": 12.0, "bg": 0.1}[args.workload]
    budget = 150.0 / max(1, args.steps + args.warmup)
    sample = int(max(min(full, 200), min(full, full * budget / per_frame_s)))
    fps, kind, text, step_s = cpu_arm(args.workload, threads, sample, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": metric_name(args.workload), "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload][3], "parallelism": f"{threads} host threads, one context each"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind, "sample": text},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def sharding_frames(n_frames: int, rank: int, world: int):
    from libnativecpurenderer_b200 import sharding

    return sharding.frames_for_rank(n_frames, rank, world)


def metric_name(workload: str) -> str:
    res = "4K" if WORKLOADS[workload][0] == 3840 else "1080p"
    return f"{res} frames/s ({workload})"


def run_product(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ["NCR_DEVICE"] = str(local_rank)

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod

        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod

    def barrier():
        if dist:
            dist.barrier()

    def max_over_ranks(v: float) -> float:
        if not dist:
            return v
        import torch

        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v: float) -> float:
        if not dist:
            return v
        import torch

        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    from libnativecpurenderer_b200 import trace
    from libnativecpurenderer_b200.binding import Renderer

    R = Renderer()   # raises if the CUDA library is missing; CreateRenderContext returns NULL without a GPU
    w, h, alpha, tex_np, arr, draws, full = build_workload(args.workload)
    tex = [R.Texture.from_numpy(t) for t in tex_np]
    ctx = R.RenderContext(w, h, alpha)
    ipp = 4 if alpha else 3
    frame_bytes = w * h * ipp
    pinned = R.lib.NcrAllocHost(frame_bytes)
    assert pinned, R.last_error()

    # one pass through the ABI: records, uploads, renders, reads back; leaves the batch resident in HBM
    ctx.set_stats_mode(1)
    trace.submit_trace(ctx, arr, tex)
    ctx.get_buffer_as_uint8_into(pinned)
    st = ctx.stats()
    ctx.set_stats_mode(0)
    trace.submit_trace(ctx, arr, tex)
    ctx.get_buffer_as_uint8_into(pinned)
    n_cmds, fine_entries, blended = st.n_cmds, st.fine_entries, st.blended_pixels
    aux_bytes = 0
    h2d = n_cmds * (240 + 16 + 4) + aux_bytes

    # ---- value: K steps from resident buffers -------------------------------------------------------------
    K, W = args.steps, max(args.warmup, 3)
    ms = (ctypes.c_float * (4 * max(K, W)))()
    assert R.lib.NcrRerunLastFlush(ctx._ptr, W, 1, ms) == 0, R.last_error()
    launches0 = R.lib.NcrKernelLaunchCount()
    barrier()
    with ClockSampler(local_rank) as clocks:
        t0 = time.perf_counter()
        assert R.lib.NcrRerunLastFlush(ctx._ptr, K, 1, ms) == 0, R.last_error()
        wall_value = time.perf_counter() - t0
    barrier()
    launches = R.lib.NcrKernelLaunchCount() - launches0
    per = np.frombuffer(ms, dtype=np.float32)[: 4 * K].reshape(K, 4).astype(np.float64)
    dev_s = float(per[:, 0].sum()) / 1e3
    dev_s_max = max_over_ranks(dev_s)
    value = world * K / dev_s_max
    comp_ms = float(per[:, 3].mean())
    clk = clocks.summary()

    # ---- e2e: host buffers in, host frame out, through the reference C ABI --------------------------------
    rp = trace.Replayer(os.path.join(ROOT, "libnativecpurenderer_b200", "lib", "libncr_replay.so"), R.path)   # the replayer is only a C caller
    if args.present == "yuv420p":
        rp.set_present("yuv420p")   # video present path: planes come back instead of the RGB(A)8 image
    d2h_bytes = w * h + 2 * ((w + 1) // 2) * ((h + 1) // 2) if args.present == "yuv420p" else frame_bytes
    # 8 contexts per GPU keep it fed; with fewer cores than contexts per rank the library's waits yield the core (NCR_SYNC auto)
    T = args.e2e_threads or 8
    e2e_frames_per_thread = max(2, min(K, args.e2e_frames))
    barrier()
    e2e_s = rp.run_threads(T, w, h, alpha, arr, tex, repeats=e2e_frames_per_thread, warm_repeats=3)
    barrier()
    e2e_value = world * T * e2e_frames_per_thread / max_over_ranks(e2e_s)
    # single-context latency view of the same path
    lat = []
    for _ in range(5):
        t0 = time.perf_counter()
        rp.run(ctx, arr, tex, frame_address=pinned)
        lat.append(time.perf_counter() - t0)

    # ---- optional: N consecutive DISTINCT frames through the batch renderer (BASELINE configs 4/5 as stated) ----
    video = None
    if args.video > 0 and args.workload in ("c4", "c5"):
        from libnativecpurenderer_b200 import batch, streams

        n_video = len(sharding_frames(args.video, rank, world))
        slots = [trace.TexSlot(k, t.shape[1], t.shape[0]) for k, t in enumerate(tex_np)]
        vtraces = []
        for f in sharding_frames(args.video, rank, world):   # this rank's frames (frame f -> rank f mod N, no collective)
            rec = trace.TraceRecorder(w, h, alpha)
            streams.stream_c4_frame(rec, slots[0], slots[1:], frame=f, n_notes=full)
            rec.present()
            vtraces.append(rec.as_array())
        seen = []
        with batch.FramePool(R, w, h, alpha, workers=T) as pool:
            pool.render(vtraces[: 2 * T], tex, present=args.present)   # warm-up: first-use allocations of every context
            barrier()
            t0 = time.perf_counter()
            pool.render(vtraces, tex, on_frame=lambda i, px: seen.append(i), present=args.present)
            v_s = time.perf_counter() - t0
            barrier()
        assert seen == list(range(n_video))
        video = {"frames": args.video, "value": args.video / max_over_ranks(v_s), "unit": "frames/s", "workers_per_gpu": T,
                 "present": args.present, "d2h_bytes_per_frame": int(d2h_bytes),
                 "h2d_bytes_per_frame": int(h2d),
                 "path": "N distinct consecutive frames: recorded traces -> NcrRenderFrames (worker contexts, in-order delivery)"}

    # ---- roofline of the dominant kernel (ncr_composite) --------------------------------------------------
    peak, peak_src = measured_peak_gbs()
    tex_bytes = sum(int(t.nbytes) for t in tex_np)
    fb_out = w * h * ipp * 8
    fb_in = 0   # every workload here starts with SetColor: the canvas is not read
    algo = fb_in + fb_out + frame_bytes + tex_bytes + n_cmds * 240 + fine_entries * 4
    achieved = algo / (comp_ms * 1e-3) / 1e9
    # f64 pipe view: ~26 f64 flops per blended pixel-op
    blended_per_s = blended * value / world
    # FP64-pipe view: the reference's expression trees cost >= 26 f64 operations per blended pixel-op
    #  and may not be fused; the peak is the measured rate of non-fused DMUL/DADD on this GPU.
    f64_peak = R.lib.NcrMeasureF64Rate()
    f64_achieved = blended * 26.0 / (comp_ms * 1e-3)

    line = {
        "metric": metric_name(args.workload), "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": dev_s_max / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload][3], "draw_calls": draws, "recorded_commands": int(n_cmds),
                   "tile_list_entries": int(fine_entries), "blended_pixel_ops_per_frame": int(blended),
                   "parallelism": f"frame-sharded replicas x{world}, no collective",
                   "l2": "256 MB memset scrubs L2 before every timed step"},
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h_bytes),
                "host_threads": T, "frames_per_thread": e2e_frames_per_thread,
                "single_context_ms_per_frame": statistics.median(lat) * 1e3,
                "path": "trace -> C ABI calls (state machine + recorder) -> H2D -> bin+composite -> D2H into pinned host memory of "
                        + ("the YUV 4:2:0 planes (NcrGetBufferAsYUV420P, parity unpinned)" if args.present == "yuv420p" else "the RGB(A)8 frame (GetBufferAsUInt8)")},
        "gpu_launches": int(launches),
        "clocks": {"sm_mhz": clk["sm_mhz"], "sm_max_mhz": clk["sm_max_mhz"], "reasons": clk["reasons"]},
        "roofline": {"bound": "hbm", "kernel": "ncr_composite", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": int(algo), "kernel_ms": comp_ms,
                     "note": "high-overdraw streams are bound by the FP64 pipe, not HBM (DESIGN.md); see blended_gpixel_per_s"},
        "kernel_ms": {"step": float(per[:, 0].mean()), "ncr_bin_coarse": float(per[:, 1].mean()),
                      "ncr_bin_fine": float(per[:, 2].mean()), "ncr_composite": comp_ms},
        "roofline_fp64": {"bound": "fp64 pipe (non-fused mul/add)", "kernel": "ncr_composite", "achieved": f64_achieved / 1e12,
                          "peak": f64_peak / 1e12, "unit": "T f64 instr/s", "frac": (f64_achieved / f64_peak) if f64_peak else None,
                          "algorithmic_f64_ops_per_blended_pixel_op": 26,
                          "peak_source": "measured in this run: NcrMeasureF64Rate (8 independent DMUL->DADD chains per thread)"},
        "blended_gpixel_per_s": sum_over_ranks(blended_per_s) / 1e9 if dist else blended_per_s / 1e9,
        "gpixel_per_s": value * w * h / 1e9,
        "device": R.lib.NcrDeviceName().decode(),
        "host_wall_s_value_region": wall_value,
    }
    if video:
        line["video"] = video

    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload not in ("c2x", "c3p"):
        threads = min(host_threads(), 64)
        sample = {"c1": 1000, "c2": 20000, "c3": 12000, "c4": 1500, "c5": 1500, "bg": 2}[args.workload]
        fps, kind, text, _ = cpu_arm(args.workload, threads, sample, 1, 0)
        line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind, "sample": text}
    elif rank == 0:
        line["cpu_baseline"] = None

    traffic_path = os.path.join(ROOT, "profiles", "composite_traffic.json")
    if os.path.exists(traffic_path):
        try:
            line["roofline"]["traffic"] = json.load(open(traffic_path)).get(args.workload)
        except ValueError:
            pass

    if rank == 0:
        print(json.dumps(line), flush=True)
    R.lib.NcrFreeHost(pinned)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--e2e-threads", type=int, default=0)
    ap.add_argument("--e2e-frames", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--video", type=int, default=0,
                    help="c4/c5 only: also render this many DISTINCT consecutive frames through NcrRenderFrames (configs 4/5)")
    ap.add_argument("--present", default="u8", choices=["u8", "yuv420p"],
                    help="what the e2e leg reads back per frame: the RGB(A)8 image (reference ABI) or the YUV 4:2:0 planes")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
