"""Generates the committed parity fixtures from the UNMODIFIED reference build (oracle/_ref).

Run in the build container (needs /root/reference and `make -C oracle ref`):

    python tests/golden/make_golden.py

Outputs (committed):
  image_rgba.npz   raw RGBA of the reference's test_files/image.png (the only media file on the path)
  golden.json      sha1 of the reference's u8 readback / f64 canvas for the streams in cases.py
  golden_bilinear.json   the same for cases.bilinear_cases(), from oracle/_ref/libNativeCPURenderer_bilinear.so — the reference
                   translation unit with its own commented-out four-tap sampler (cpp:575-620) switched on by oracle/Makefile
  golden_polygon.json  cases.polygon_cases() from oracle/_ref/libNativeCPURenderer_polygon.so: the unmodified reference translation
                   unit + oracle/ref_polygon_shim.cpp (DrawLine's loop through the reference's own pointInPolygon / ApplyPixel)
  golden_perspective.json   cases.perspective_cases() from the same shim build: the projective map is this repo's spec, the bounds /
                   sampling / blend after it are the reference's own DrawTexture tail (InterpolateColorFromBuffer, ApplyPixel)
  golden_apply_pixel.json   the random streams that call ApplyPixel directly, from the same shim build (it exports the reference's
                   inline ApplyPixel, cpp:515-549)
  golden_clip.json  cases.clip_cases(): the UNMODIFIED reference drawing unclipped, with the pixels outside the clip rect put back
                   after every draw (cases.ClipEmulated) — what the clip-rect extension must reproduce

/root/reference does not exist on the GPU box, so the GPU tests compare against these files.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from libnativecpurenderer_b200.binding import Renderer  # noqa: E402
import cases  # noqa: E402


def main():
    from PIL import Image

    img = Image.open("/root/reference/test_files/image.png")
    assert img.mode == "RGBA" and img.size == (128, 128)
    rgba = np.frombuffer(img.tobytes(), dtype=np.uint8).reshape(128, 128, 4)
    np.savez_compressed(os.path.join(HERE, "image_rgba.npz"), rgba=rgba)

    ref = Renderer(os.path.join(ROOT, "oracle", "_ref", "libNativeCPURenderer.so"))
    out = {}
    for name, fn in cases.all_cases(reference_abi_only=True):
        out[name] = fn(ref, rgba)
        print(name, out[name])
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)

    refb = Renderer(os.path.join(ROOT, "oracle", "_ref", "libNativeCPURenderer_bilinear.so"))
    outb = {}
    for name, fn in cases.bilinear_cases():
        outb[name] = fn(refb, rgba, switch=False)   # that build has no switch: it always samples with four taps
        print("bilinear", name, outb[name])
    with open(os.path.join(HERE, "golden_bilinear.json"), "w") as f:
        json.dump(outb, f, indent=1, sort_keys=True)

    refp = Renderer(os.path.join(ROOT, "oracle", "_ref", "libNativeCPURenderer_polygon.so"))
    outp = {}
    for name, fn in cases.polygon_cases():
        outp[name] = fn(refp, rgba)
        print("polygon", name, outp[name])
    with open(os.path.join(HERE, "golden_polygon.json"), "w") as f:
        json.dump(outp, f, indent=1, sort_keys=True)

    outq = {}
    for name, fn in cases.perspective_cases():
        outq[name] = fn(refp, rgba)
        print("perspective", name, outq[name])
    with open(os.path.join(HERE, "golden_perspective.json"), "w") as f:
        json.dump(outq, f, indent=1, sort_keys=True)

    # ApplyPixel is declared in the reference header (h:109) but defined `inline` (cpp:515), so the plain reference build does not
    # export it; the shim build does (oracle/ref_polygon_shim.cpp forwards to the reference's own function)
    outa = {}
    for name, fn in cases.all_cases(reference_abi_only=False):
        if name.startswith("random_ap_"):
            outa[name] = fn(refp, rgba)
            print("apply_pixel", name, outa[name])
    with open(os.path.join(HERE, "golden_apply_pixel.json"), "w") as f:
        json.dump(outa, f, indent=1, sort_keys=True)

    outc = {}
    for name, fn in cases.clip_cases():
        outc[name] = fn(ref, rgba, native=False)   # the reference has no clip rect: emulated around its own draws
        print("clip", name, outc[name])
    with open(os.path.join(HERE, "golden_clip.json"), "w") as f:
        json.dump(outc, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
