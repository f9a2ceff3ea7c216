"""Pins the present path to a REAL libswscale: runs the call PutRendererContextFrame makes (reference
src/libNativeCPURenderer.cpp:241-256: sws_getContext(w, h, RGBA|RGB24, w, h, YUV420P, SWS_BILINEAR, 0, 0, 0) + sws_scale) through
ctypes on a handful of images and stores inputs and outputs as fixtures.

    python tests/golden/make_swscale_fixtures.py        # needs a libswscale; uses the one bundled with opencv-python-headless

Output (committed): swscale_fixtures.npz  (img_k, yuv_k arrays; `scaled` = (image, dst_w, dst_h) rows with their syuv_n planes;
`version` = libswscale version the fixtures came from).
The reference pins no FFmpeg version; these fixtures are libswscale 9.1.100 (FFmpeg 8) on x86-64 (AVX2 host), no SWS flags
beyond SWS_BILINEAR — i.e. the non-bitexact SIMD vertical scaler, which is what the reference runs."""
import ctypes
import glob
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SWS_BILINEAR = 2


def load_swscale():
    """(libswscale, libavutil) from the opencv wheel's private library directory, or None."""
    try:
        import cv2
    except ImportError:
        return None
    libdir = os.path.join(os.path.dirname(os.path.dirname(cv2.__file__)), "opencv_python_headless.libs")
    found = {}
    for name in ("libavutil", "libswresample", "libswscale"):
        hits = sorted(glob.glob(os.path.join(libdir, name + "-*.so*")))
        if not hits:
            return None
        found[name] = hits[0]
    # the wheel's libraries find each other by name inside libdir: preload that directory's dependencies through cv2 (already
    # imported above, which maps them), then open the three we need
    try:
        avutil = ctypes.CDLL(found["libavutil"], mode=ctypes.RTLD_GLOBAL)
        ctypes.CDLL(found["libswresample"], mode=ctypes.RTLD_GLOBAL)
        sws = ctypes.CDLL(found["libswscale"], mode=ctypes.RTLD_GLOBAL)
    except OSError:
        return None
    sws.swscale_version.restype = ctypes.c_uint
    avutil.av_get_pix_fmt.restype = ctypes.c_int
    avutil.av_get_pix_fmt.argtypes = [ctypes.c_char_p]
    sws.sws_getContext.restype = ctypes.c_void_p
    sws.sws_getContext.argtypes = [ctypes.c_int] * 7 + [ctypes.c_void_p] * 3
    sws.sws_scale.restype = ctypes.c_int
    sws.sws_scale.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    sws.sws_freeContext.argtypes = [ctypes.c_void_p]
    return sws, avutil


def swscale_yuv420p(libs, img: np.ndarray, dst_w: int | None = None, dst_h: int | None = None) -> np.ndarray:
    """img: (h, w, 3|4) uint8 -> planar Y, U, V concatenated, exactly as the reference's PutRendererContextFrame converts
    (dst_w x dst_h = the VideoCap's size; default: the canvas size)."""
    sws, avutil = libs
    h, w, c = img.shape
    dw, dh = dst_w or w, dst_h or h
    src_fmt = avutil.av_get_pix_fmt(b"rgba" if c == 4 else b"rgb24")
    dst_fmt = avutil.av_get_pix_fmt(b"yuv420p")
    ctx = sws.sws_getContext(w, h, src_fmt, dw, dh, dst_fmt, SWS_BILINEAR, None, None, None)
    assert ctx
    src = np.zeros(w * h * c + 64, dtype=np.uint8)     # libswscale reads a pixel past an odd-width row: keep that inside the buffer
    src[: w * h * c] = np.ascontiguousarray(img).ravel()
    cw, ch = (dw + 1) // 2, (dh + 1) // 2
    out = np.zeros(dw * dh + 2 * cw * ch, dtype=np.uint8)
    y, u, v = out[: dw * dh], out[dw * dh: dw * dh + cw * ch], out[dw * dh + cw * ch:]
    srcp = (ctypes.c_void_p * 4)(src.ctypes.data, None, None, None)
    srcs = (ctypes.c_int * 4)(w * c, 0, 0, 0)
    dstp = (ctypes.c_void_p * 4)(y.ctypes.data, u.ctypes.data, v.ctypes.data, None)
    dsts = (ctypes.c_int * 4)(dw, cw, cw, 0)
    assert sws.sws_scale(ctx, srcp, srcs, 0, h, dstp, dsts) == dh
    sws.sws_freeContext(ctx)
    return out


def fixture_images():
    rs = np.random.RandomState(2026)
    imgs = [rs.randint(0, 256, (64, 96, 4)).astype(np.uint8),                  # noise, RGBA
            rs.randint(0, 256, (34, 50, 3)).astype(np.uint8),                  # noise, RGB24, width not a multiple of 16
            (rs.randint(0, 2, (16, 16, 3)) * 255).astype(np.uint8),            # saturated primaries
            np.zeros((8, 8, 4), dtype=np.uint8)]                               # smallest pinned size
    yy, xx = np.mgrid[0:48, 0:64]
    imgs.append(np.stack([xx * 4 % 256, yy * 5 % 256, (xx + yy) * 2 % 256], axis=-1).astype(np.uint8))   # ramps
    imgs[3][...] = [[[255, 255, 255, 7]]]
    return imgs


def main():
    libs = load_swscale()
    if libs is None:
        sys.exit("no libswscale found (opencv-python-headless bundles one)")
    v = libs[0].swscale_version()
    out = {"version": np.array([v >> 16, (v >> 8) & 255, v & 255])}
    for k, img in enumerate(fixture_images()):
        out[f"img_{k}"] = img
        out[f"yuv_{k}"] = swscale_yuv420p(libs, img)
    # the scaling branch (cap size != canvas size): shrink, enlarge, odd destination sizes, one axis only
    scaled = [(0, 64, 48), (0, 48, 32), (1, 75, 51), (1, 26, 18), (4, 128, 96), (4, 64, 20), (2, 16, 8), (0, 96, 32), (0, 40, 64)]
    out["scaled"] = np.array(scaled)
    for n, (k, dw, dh) in enumerate(scaled):
        out[f"syuv_{n}"] = swscale_yuv420p(libs, out[f"img_{k}"], dw, dh)
    np.savez_compressed(os.path.join(HERE, "swscale_fixtures.npz"), **out)
    print("libswscale", out["version"], "fixtures:", len(fixture_images()))


if __name__ == "__main__":
    main()
