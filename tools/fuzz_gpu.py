#!/usr/bin/env python
"""Development helper: long fuzz campaign of the CUDA path against the CPU checkers (run under gpurun).

    python tools/fuzz_gpu.py FIRST_SEED COUNT [--ref]   -> gpurun_out/fuzz.log

Per seed: the randomised reference-ABI stream of tests/cases.py (five canvas shapes, three flushes each, u8 + f64 digests,
YUV planes) against the C restatement, every 10th seed also against the unmodified reference build (--ref), plus the
extension stream (clip / bilinear / polygon / perspective, mixed) against the restatement.

    python tools/fuzz_gpu.py FIRST_SEED COUNT --pins

Per seed: each extension on its own against its reference-derived build (oracle/Makefile): bilinear vs the reference with its
commented-out sampler switched on, clip rect vs the unmodified reference with the outside pixels put back, N-gon fill and perspective
quads vs the shim build (DrawLine's / DrawTexture's loops through the reference's own functions)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import cases  # noqa: E402
from libnativecpurenderer_b200 import streams  # noqa: E402
from libnativecpurenderer_b200.binding import Renderer  # noqa: E402


def pins(first, count):
    gpu = Renderer()
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    ref = Renderer(os.path.join(ref_dir, "libNativeCPURenderer.so"))
    refb = Renderer(os.path.join(ref_dir, "libNativeCPURenderer_bilinear.so"))
    refp = Renderer(os.path.join(ref_dir, "libNativeCPURenderer_polygon.so"))
    img = np.load(os.path.join(ROOT, "tests", "golden", "image_rgba.npz"))["rgba"]
    bad, t0 = [], time.time()
    for seed in range(first, first + count):
        # RGBA shapes only on the reference side of the bilinear stream: SetColor on a 3-channel canvas overruns its buffer there
        if cases.RANDOM_SHAPES[seed % len(cases.RANDOM_SHAPES)][2]:
            run = cases.make_bilinear_case(f"random_{seed}")
            if run(gpu, img, switch=True) != run(refb, img, switch=False):
                bad.append(("bilinear", seed))
        run = cases.make_clip_case(seed)
        if run(gpu, img, native=True) != run(ref, img, native=False):
            bad.append(("clip", seed))
        run = cases.make_polygon_case(seed)
        if run(gpu, img) != run(refp, img):
            bad.append(("polygon", seed))
        run = cases.make_perspective_case(seed)
        if run(gpu, img) != run(refp, img):
            bad.append(("perspective", seed))
        if (seed - first) % 50 == 49:
            print(f"{seed - first + 1} seeds, {len(bad)} mismatches, {time.time() - t0:.0f} s", flush=True)
    print(f"DONE pins, seeds {first}..{first + count - 1}: {len(bad)} mismatches {bad[:20]}")
    return 1 if bad else 0


def main():
    first, count = int(sys.argv[1]), int(sys.argv[2])
    if "--pins" in sys.argv:
        return pins(first, count)
    use_ref = "--ref" in sys.argv
    gpu = Renderer()
    port = Renderer(os.path.join(ROOT, "oracle", "libncr_oracle.so"))
    ref = Renderer(os.path.join(ROOT, "oracle", "_ref", "libNativeCPURenderer.so")) if use_ref else None
    img = np.load(os.path.join(ROOT, "tests", "golden", "image_rgba.npz"))["rgba"]
    bad, t0 = [], time.time()
    for seed in range(first, first + count):
        run = cases.make_random_case(seed, use_apply_pixel=(seed % 2 == 0))
        g = run(gpu, img)
        if g != run(port, img):
            bad.append(("abi/port", seed))
        if ref is not None and seed % 10 == 1:   # odd seeds only: the reference build does not export ApplyPixel
            if g != run(ref, img):
                bad.append(("abi/ref", seed))
        w, h, alpha = [(160, 90, True), (97, 61, False), (256, 144, True)][seed % 3]
        got = []
        for R in (gpu, port):
            ctx = R.RenderContext(w, h, alpha)
            tex = cases.tiny_textures(R, img)
            tex.append(R.Texture(7, 6, True, np.random.RandomState(5).rand(6, 7, 4).tobytes(), is_uint8=False))
            tex.append(R.Texture.from_numpy(np.random.RandomState(8).randint(0, 256, (12, 10, 3)).astype(np.uint8)))
            streams.stream_extensions(ctx, tex, seed, n=80)
            rs = np.random.RandomState(seed)   # present path incl. the scaling branch: same planes at a random other size
            dw, dh = int(rs.randint(2, 2 * w)), int(rs.randint(2, 2 * h))
            got.append((ctx.get_buffer_as_yuv420p().tobytes(), ctx.get_buffer_as_yuv420p_scaled(dw, dh).tobytes(), cases.digest(ctx)))
        if got[0] != got[1]:
            bad.append(("ext/port", seed))
        if (seed - first) % 100 == 99:
            print(f"{seed - first + 1} seeds, {len(bad)} mismatches, {time.time() - t0:.0f} s", flush=True)
    print(f"DONE seeds {first}..{first + count - 1}: {len(bad)} mismatches {bad[:20]}")
    return 1 if bad else 0


if __name__ == "__main__":
    raise SystemExit(main())
