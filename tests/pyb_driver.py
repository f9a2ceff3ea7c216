"""Runs inside a subprocess whose working directory holds `libNativeCPURenderer.so` (product or reference build) and whose
sys.path holds the reference's UNCHANGED ctypes binding (oracle/_ref/pyb/libNativeCPURendererPybind.py, staged by
oracle/Makefile): SURVEY K1 (BASELINE config 1) and the binding's own smoke loop K2 (pyb:675-716, minus audio/video) are
rendered through that file — every call a per-call ctypes crossing, exactly as a user of the reference makes them — and the
sha1 digests are printed as JSON.  argv: path of tests/golden/image_rgba.npz, then the case names."""
import hashlib
import json
import math
import random
import sys
import time

import numpy as np
from PIL import Image

import libNativeCPURendererPybind as CPURenderer   # binds ./libNativeCPURenderer.so at import (pyb:9)


def sha(b) -> str:
    return hashlib.sha1(bytes(b)).hexdigest()


def k1(rgba):
    ctx = CPURenderer.RenderContext(1920, 1080, True)
    ctx.set_color(0, 0, 0, 1)
    tex = CPURenderer.Texture.from_pilimg(Image.fromarray(rgba, "RGBA"))
    rng = random.Random(0)
    t0 = time.perf_counter()
    for _ in range(1000):
        ctx.save_state()
        ctx.translate(rng.uniform(0, 1920), rng.uniform(0, 1080))
        ctx.rotate(rng.uniform(0, 2 * math.pi))
        s = rng.uniform(0.5, 2.0)
        ctx.scale(s, s)
        ctx.apply_color_transform(1, 1, 1, rng.uniform(0.2, 0.9))
        ctx.draw_texture(tex, -64, -64, 128, 128)
        ctx.restore_state()
    u8 = ctx.get_buffer_as_uint8()
    dt = time.perf_counter() - t0
    return {"u8": sha(u8), "seconds": dt, "pil_size": list(ctx.as_pilimg().size)}


def k2(rgba):
    ctxS = 4
    ctx = CPURenderer.RenderContext(1024 // ctxS, 1024 // ctxS, True)
    ctx.scale(1 / ctxS, 1 / ctxS)
    tex = CPURenderer.Texture.from_pilimg(Image.fromarray(rgba, "RGBA")).resample(16, 16)
    out = {}
    for i in range(121):
        t = i / 60
        ctx.set_color(1, 1, 1, 1)
        ctx.save_state()
        ctx.apply_color_transform(t % 1, (t + 1.4) % 1, (t + 2.8) % 1, 1)
        w = 768 * (1 + math.sin(t * 2 * math.pi) / 4)
        h = 768 * (1 + math.cos(t * 3 * math.pi) / 4)
        ctx.draw_texture(tex, w * 1.5 / 2, h * 1.3 / 2, w, h)
        ctx.draw_line(w * 0.1, h * 0.1, w, h, (w + h) / 300, 0, 1, 0, 1)
        ctx.draw_circle(w * 0.3, h * 0.3, 100, 1, 1, 0, 0.4);
        ctx.draw_rect(w * 0.6, h * 0.6, w * 0.1, h * 0.1, 0, 1, 0, 0.4)
        ctx.restore_state()
        if i in (0, 1, 30, 60, 120):
            out[f"u8_{i}"] = sha(ctx.get_buffer_as_uint8())
    return out


def main():
    rgba = np.load(sys.argv[1])["rgba"]
    res = {"version": CPURenderer.get_version()}
    for name in sys.argv[2:]:
        res[name] = {"k1": k1, "k2": k2}[name](rgba)
    print("PYB_RESULT " + json.dumps(res))


if __name__ == "__main__":
    main()
