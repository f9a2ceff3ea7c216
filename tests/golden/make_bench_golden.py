"""Full-size parity fixtures for bench.py's parity gate and the full-size GPU tests, generated from the UNMODIFIED reference
build (oracle/_ref) — and, for the two product-only extension workloads and the YUV present path, from the C restatement
(oracle/libncr_oracle.so; "parity unpinned": that column is labelled `restatement`).

    python tests/golden/make_bench_golden.py            # everything: ~30 minutes on 8 cores (c3 / c3p dominate)
    python tests/golden/make_bench_golden.py --only-yuv # only the restatement's YUV planes of the video frames (~3 minutes)

Output (committed): bench_golden.json
  workloads.<name>.u8            sha1 of the GetBufferAsUInt8 frame of bench.py's workload <name> at its full size
  video.<c4|c5>.<frame>.u8       the same for chart frame <frame> of the video legs (distinct consecutive frames)
  video.<c4|c5>.<frame>.yuv420p  sha1 of the YUV 4:2:0 planes (restatement)
/root/reference does not exist on the GPU box: bench.py and the GPU tests only read this file."""
import hashlib
import json
import multiprocessing as mp
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

REF = os.path.join(ROOT, "oracle", "_ref", "libNativeCPURenderer.so")
PORT = os.path.join(ROOT, "oracle", "libncr_oracle.so")
REPLAY = os.path.join(ROOT, "libnativecpurenderer_b200", "lib", "libncr_replay.so")

WORKLOADS_REF = ["c1", "c2", "c3", "c4", "c5", "bg"]
WORKLOADS_PORT = ["c2x", "c3p"]
VIDEO_FRAMES = {"c4": [0, 123, 359, 719, 3600, 7199], "c5": [0, 123, 239, 479, 959, 1919]}


def render(lib, w, h, alpha, tex_np, arr, want_yuv=False):
    from libnativecpurenderer_b200 import trace
    from libnativecpurenderer_b200.binding import Renderer

    R = Renderer(lib)
    tex = [R.Texture.from_numpy(t) for t in tex_np]
    ctx = R.RenderContext(w, h, alpha)
    trace.Replayer(REPLAY, lib).run(ctx, arr, tex)
    out = {"u8": hashlib.sha1(bytes(ctx.get_buffer_as_uint8())).hexdigest()}
    if want_yuv:
        out["yuv420p"] = hashlib.sha1(ctx.get_buffer_as_yuv420p().tobytes()).hexdigest()
    return out


def job(item):
    import bench

    kind, name, frame, lib = item
    if kind == "workload":
        w, h, alpha, tex_np, arr, draws, full = bench.build_workload(name)
        return item, render(lib, w, h, alpha, tex_np, arr)
    w, h, alpha, tex_np, arr = bench.build_video_frame(name, frame)
    return item, render(lib, w, h, alpha, tex_np, arr, want_yuv=(lib == PORT))


def main():
    only_yuv = "--only-yuv" in sys.argv
    items = [] if only_yuv else [("workload", n, None, REF) for n in WORKLOADS_REF] + [("workload", n, None, PORT) for n in WORKLOADS_PORT]
    for n, frames in VIDEO_FRAMES.items():
        for f in frames:
            if not only_yuv:
                items.append(("video", n, f, REF))
            items.append(("video", n, f, PORT))
    items.sort(key=lambda it: {"c3": 0, "c3p": 0, "c2": 1, "c2x": 1, "c5": 2}.get(it[1], 3))   # longest first
    out = {"workloads": {}, "video": {"c4": {}, "c5": {}},
           "source": {"u8": "unmodified reference build (oracle/_ref)",
                      "yuv420p": "C restatement of libswscale's conversion (pinned to libswscale 9.1.100 fixtures, tests/test_oracle.py)",
                      "c2x": "C restatement (product-only extensions, parity unpinned)",
                      "c3p": "C restatement (product-only extensions, parity unpinned)"}}
    if only_yuv:
        with open(os.path.join(HERE, "bench_golden.json")) as f:
            out = json.load(f)
        out["source"]["yuv420p"] = "C restatement of libswscale's conversion (pinned to libswscale 9.1.100 fixtures, tests/test_oracle.py)"
    with mp.get_context("spawn").Pool(min(len(items), os.cpu_count() or 1)) as pool:
        for item, res in pool.imap_unordered(job, items):
            kind, name, frame, lib = item
            print(kind, name, frame, "ref" if lib == REF else "port", res, flush=True)
            if kind == "workload":
                out["workloads"][name] = res
            else:
                slot = out["video"][name].setdefault(str(frame), {})
                if lib == REF:
                    slot["u8"] = res["u8"]
                else:
                    slot["yuv420p"] = res["yuv420p"]
                    slot["u8_restatement"] = res["u8"]
    for n in out["video"]:
        for f, slot in out["video"][n].items():
            assert slot["u8"] == slot["u8_restatement"], (n, f)   # the restatement agrees with the reference on every frame
    with open(os.path.join(HERE, "bench_golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
