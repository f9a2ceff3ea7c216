"""Parity cases shared by the golden generator, the CPU tests and the GPU tests.

Each case is ``fn(renderer, image_rgba) -> {"u8": sha1, "f64": sha1, ...}`` and runs the same calls on
whichever library ``renderer`` wraps (reference build, C restatement, product)."""
from __future__ import annotations

import hashlib

import numpy as np

from libnativecpurenderer_b200 import streams


def sha(b) -> str:
    return hashlib.sha1(bytes(b)).hexdigest()


def digest(ctx) -> dict:
    return {"u8": sha(ctx.get_buffer_as_uint8()), "f64": sha(ctx.get_buffer_np().tobytes())}


def swscale_fixtures():
    """Inputs and outputs of a real libswscale for the present path's conversion (tests/golden/make_swscale_fixtures.py)."""
    import os

    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "swscale_fixtures.npz"))


def swscale_model(img):
    """numpy restatement of libswscale's RGB(A) -> YUV420P conversion for even sizes (the formulas of oracle/ncr_oracle.c,
    vectorised; validated against the real library by tests/test_oracle.py).  Returns the concatenated Y, U, V planes."""
    h, w, _ = img.shape
    assert h % 2 == 0 and w % 2 == 0 and h >= 4
    r, g, b = (img[..., k].astype(np.int64) for k in range(3))
    y = np.clip((((8414 * r + 16519 * g + 3208 * b + (32 << 14) + (1 << 8)) >> 9 << 1) + 64) >> 7, 0, 255)
    r2, g2, b2 = (c[:, 0::2] + c[:, 1::2] for c in (r, g, b))
    u15 = np.minimum(((-4865 * r2 - 9528 * g2 + 14392 * b2 + (0x4001 << 9)) >> 10) << 1, 32767)
    v15 = np.minimum(((14392 * r2 - 12061 * g2 - 2332 * b2 + (0x4001 << 9)) >> 10) << 1, 32767)

    def vertical(p):
        ch = h // 2
        out = np.zeros((ch, w // 2), dtype=np.int64)
        for c in range(ch):
            taps = {}
            for k, cf in zip(range(2 * c - 1, 2 * c + 3), (512, 1536, 1536, 512)):
                kk = min(max(k, 0), h - 1)
                taps[kk] = taps.get(kk, 0) + cf
            if c == ch - 1:
                out[c] = ((64 << 12) + sum(p[kk] * cf for kk, cf in taps.items())) >> 19
            else:
                out[c] = (5 + sum((p[kk] * cf) >> 16 for kk, cf in taps.items())) >> 3
        return np.clip(out, 0, 255)

    return np.concatenate([a.astype(np.uint8).ravel() for a in (y, vertical(u15), vertical(v15))])


def canvas_holding_u8_image(R, img):
    """A context whose GetBufferAsUInt8 readback is exactly ``img`` (h, w, 3|4 uint8), built through the reference ABI: an f64
    texture with texels (k + 0.5) / 255 — (iu8)(v * 255) truncates to k for every k — drawn 1:1 on the identity path.  The
    texture is one texel larger than the canvas because the last texel row / column is never sampled."""
    h, w, c = img.shape
    tex = np.zeros((h + 1, w + 1, 4), dtype=np.float64)
    tex[:h, :w, :c] = (img.astype(np.float64) + 0.5) / 255.0
    if c == 3:
        tex[..., 3] = 1.0     # a == 1: the source colour is stored untouched (cpp:533)
    else:
        # RGBA canvas: alpha is both the stored alpha and the blend factor; blend over a black canvas would scale rgb, so the
        # image's alpha plane is injected with set_pixel afterwards
        tex[..., 3] = 1.0
    ctx = R.RenderContext(w, h, c == 4)
    ctx.set_color(0, 0, 0, 0)
    t = R.Texture(w + 1, h + 1, True, tex.tobytes(), is_uint8=False)
    ctx.draw_texture(t, 0, 0, w + 1, h + 1)
    if c == 4:
        for (j, i), a in np.ndenumerate(img[..., 3]):
            if a != 255:
                px = (img[j, i].astype(np.float64) + 0.5) / 255.0
                ctx.set_pixel(int(i), int(j), *px)
    return ctx


def tiny_textures(R, image_rgba):
    rs = np.random.RandomState(99)
    return [
        R.Texture.from_numpy(image_rgba),
        R.Texture.from_numpy(rs.randint(0, 256, (9, 5, 4)).astype(np.uint8)),
        R.Texture.from_numpy(rs.randint(0, 256, (2, 2, 4)).astype(np.uint8)),
        R.Texture.from_numpy(rs.randint(0, 256, (33, 64, 4)).astype(np.uint8)),
    ]


# ---- SURVEY.md §8c known answers -----------------------------------------------------------------
def case_k1(R, image_rgba):
    ctx = R.RenderContext(1920, 1080, True)
    streams.stream_k1(ctx, R.Texture.from_numpy(image_rgba))
    return digest(ctx)


def case_k2(R, image_rgba):
    ctx = R.RenderContext(256, 256, True)
    ctx.scale(.25, .25)
    t16 = R.Texture.from_numpy(image_rgba).resample(16, 16)
    out = {}
    for i in range(121):
        streams.stream_k2_frame(ctx, t16, i)
        if i in (0, 1, 30, 60, 120):
            out[f"u8_{i}"] = sha(ctx.get_buffer_as_uint8())
    return out


def case_k3(R, image_rgba):
    """resample(16,16) of image.png, drawn 1:1 so the texels are observable through the ABI."""
    t16 = R.Texture.from_numpy(image_rgba).resample(16, 16)
    ctx = R.RenderContext(20, 20, True)
    ctx.set_color(0, 0, 0, 0)
    ctx.draw_texture(t16, 0, 0, 16, 16)
    d = digest(ctx)
    d["size"] = [t16.width, t16.height, int(t16.enableAlpha)]
    return d


def case_k4(R, image_rgba):
    """CreateMilthmHitEffectTexture on image.png.resample(64,48) (non-square: exercises the transposed indexing)."""
    mask = R.Texture.from_numpy(image_rgba).resample(64, 48)
    fx = R.Helpers.create_milthm_hit_effect_textures(mask, 3, seed=0.25)
    out = {}
    for k, t in enumerate(fx):
        ctx = R.RenderContext(70, 50, True)
        ctx.set_color(0, 0, 0, 0)
        ctx.draw_texture(t, 0, 0, 64, 48)
        out[f"fx{k}"] = digest(ctx)["f64"]
    return out


def case_k5(R, image_rgba):
    """Micro-cases of the quirks: exact pixel counts and values."""
    out = {}
    grad = np.zeros((4, 4, 4), dtype=np.uint8)
    grad[..., 0] = np.arange(4)[None, :] * 60
    grad[..., 3] = 255
    g = R.Texture.from_numpy(grad)
    for name, args in (("q2_frac_origin", (2.5, 2.5, 4, 4)), ("q4_last_texel", (0, 0, 4, 4))):
        ctx = R.RenderContext(12, 12, True)
        ctx.set_color(0, 0, 0, 0)
        ctx.draw_texture(g, *args)
        out[name] = digest(ctx)["f64"]
    for name, rect in (("q3_rect_int", (2, 2, 4, 4)), ("q3_rect_frac", (2.5, 2.5, 4, 4))):
        ctx = R.RenderContext(12, 12, True)
        ctx.set_color(0, 0, 0, 0)
        ctx.translate(0, 1e-3)   # force the transformed path for rects regardless
        ctx.draw_rect(*rect, 1, 0, 0, 1)
        buf = ctx.get_buffer_np().reshape(12, 12, 4)
        out[name] = int((buf[..., 0] == 1).sum())
    ctx = R.RenderContext(8, 8, True)   # quirk 1: scale(.25) and translate(-3,-3) are "no transform" for DrawTexture
    ctx.set_color(0, 0, 0, 0)
    ctx.scale(.25, .25)
    ctx.translate(-3, -3)
    ctx.draw_texture(g, 1, 1, 4, 4)
    out["q1_ignored_matrix"] = digest(ctx)["f64"]
    ctx = R.RenderContext(4, 4, True)   # quirk 5 + 8
    ctx.set_color(.5, .5, .5, 1)
    ctx.translate(0, 1e-3)
    ctx.draw_rect(0, 0, 4, 4, 1, 0, 0, .25)
    out["q5_alpha_overwrite"] = list(ctx.get_buffer_as_uint8()[16:20])   # pixel (0,1): (159, 95, 95, 63)
    return out


def case_k6(R, image_rgba):
    ctx = R.RenderContext(320, 180, False)
    streams.stream_k6(ctx, R.Texture.from_numpy(streams.k6_texture()))
    return digest(ctx)


# ---- randomised streams --------------------------------------------------------------------------
RANDOM_SHAPES = [(97, 61, True), (64, 64, True), (130, 34, False), (16, 16, True), (257, 129, True), (48, 50, False)]


def make_random_case(seed: int, use_apply_pixel: bool = False):
    w, h, alpha = RANDOM_SHAPES[seed % len(RANDOM_SHAPES)]

    def run(R, image_rgba):
        ctx = R.RenderContext(w, h, alpha)
        ctx.set_color(.2, .2, .2, .2)   # defined starting contents (the reference's are uninitialised)
        tex = tiny_textures(R, image_rgba)
        out = {}
        for part in range(3):   # three flushes per context: later parts start from a non-trivial canvas
            streams.stream_random(ctx, tex, seed * 10 + part, n=70, use_apply_pixel=use_apply_pixel)
            out[f"part{part}"] = digest(ctx)
        return out

    return run


def case_c2_small(R, image_rgba):
    """C2's generator at 480x270 with 600 draws (finishes in seconds on the CPU)."""
    ctx = R.RenderContext(480, 270, True)
    tex = [R.Texture.from_numpy(t) for t in streams.make_c2_textures()]
    streams.stream_c2(ctx, tex, n=600)
    return digest(ctx)


def case_c3_small(R, image_rgba):
    ctx = R.RenderContext(512, 288, True)
    atlas = R.Texture.from_numpy(streams.make_atlas(cells=4, cell=64))
    streams.stream_c3(ctx, atlas, n=800, cells=4)
    return digest(ctx)


def case_c4_small(R, image_rgba):
    """Two consecutive frames of the chart-shaped stream on a 480x270 RGB canvas."""
    ctx = R.RenderContext(480, 270, False)
    tex = [R.Texture.from_numpy(t) for t in streams.make_chart_textures()]
    bg = R.Texture.from_numpy(streams.make_noise_texture(128, 7)).resample(480, 270)
    out = {}
    for f in (0, 37):
        streams.stream_c4_frame(ctx, bg, tex, f, n_notes=120, n_fx=10)
        out[f"frame{f}"] = digest(ctx)
    return out


def case_canvas_textures(R, image_rgba):
    """CreateTextureFromRenderContext (deep copy) and f64 CreateTexture, drawn back."""
    src = R.RenderContext(40, 30, True)
    src.set_color(.1, .2, .3, .4)
    src.translate(5, 5)
    src.rotate(.3)
    src.draw_rect(0, 0, 20, 10, 1, .5, .25, .75)
    copy = src.as_texure()
    src.set_color(1, 1, 1, 1)   # must not affect the copy
    dst = R.RenderContext(64, 48, True)
    dst.set_color(0, 0, 0, 1)
    dst.translate(10, 4)
    dst.rotate(-.2)
    dst.draw_texture(copy, 0, 0, 50, 40)
    vals = np.random.RandomState(5).rand(6, 7, 4)
    ftex = R.Texture(7, 6, True, vals.tobytes(), is_uint8=False)
    dst.draw_texture(ftex, 20, 10, 30, 30)
    return digest(dst)


# ---- bilinear sampling (extension X2), pinned to the reference's own commented-out sampler --------------------------------
# oracle/Makefile builds _ref/libNativeCPURenderer_bilinear.so: the reference translation unit with the four-tap code of
# cpp:575-620 un-commented.  That build samples bilinearly in EVERY textured draw and has no switch, so the same stream is run
# with NcrSetSampling(ctx, 1) on the product / restatement (`switch=True`) and without on that build.
def _bilinear_ctx(R, w, h, alpha, switch):
    ctx = R.RenderContext(w, h, alpha)
    if switch:
        ctx.set_sampling(1)
    return ctx


def make_bilinear_case(which):
    def run(R, image_rgba, switch=True):
        if which == "c2_small":
            ctx = _bilinear_ctx(R, 480, 270, True, switch)
            streams.stream_c2(ctx, [R.Texture.from_numpy(t) for t in streams.make_c2_textures()], n=600)
        elif which == "c3_small":
            ctx = _bilinear_ctx(R, 512, 288, True, switch)
            streams.stream_c3(ctx, R.Texture.from_numpy(streams.make_atlas(cells=4, cell=64)), n=800, cells=4)
        elif which == "k1_small":
            ctx = _bilinear_ctx(R, 640, 360, True, switch)
            streams.stream_k1(ctx, R.Texture.from_numpy(image_rgba), n=300)
        else:   # randomised reference-ABI stream (identity and mapped paths, degenerate sizes, RGB and RGBA canvases), three flushes
            seed = int(which.split("_")[1])
            w, h, alpha = RANDOM_SHAPES[seed % len(RANDOM_SHAPES)]
            ctx = _bilinear_ctx(R, w, h, alpha, switch)
            ctx.set_color(0, 0, 0, 0)   # a new canvas is uninitialised heap in the reference (cpp:15) and zero in the product
            tex = tiny_textures(R, image_rgba)   # RGBA8, every side >= 2 texels (the reference reads out of bounds below that)
            for part in range(3):
                streams.stream_random(ctx, tex, seed * 10 + part, n=70)
        return digest(ctx)

    return run


def bilinear_cases():
    names = ["c2_small", "c3_small", "k1_small"] + [f"random_{s}" for s in range(10)]
    return [(n, make_bilinear_case(n)) for n in names]


# ---- clip rect (extension X1), pinned to the UNMODIFIED reference ------------------------------------------------------
# "A draw under a clip rect touches only the pixels inside it" is expressible with the reference alone: let the reference draw
# unclipped, then put back every pixel outside the rect (SetPixel stores the four f64 of an RGBA pixel exactly, cpp:494-513).
class ClipEmulated:
    """Wraps a context of a library that has no clip rect (the reference build) and gives it one, pixel-exactly."""

    CLIPPED = ("fill_color", "draw_texture", "draw_splitted_texture", "draw_rect", "draw_vertical_grd", "draw_circle", "draw_line")

    def __init__(self, ctx):
        self._ctx, self._clip = ctx, None

    def set_clip_rect(self, x, y, w, h):
        W, H = self._ctx.width, self._ctx.height
        l, t = max(0, x), max(0, y)
        self._clip = (l, t, max(l, min(W, x + w)), max(t, min(H, y + h)))   # an empty rect stays empty (no negative slice ends)

    def clear_clip_rect(self):
        self._clip = None

    def __getattr__(self, name):
        fn = getattr(self._ctx, name)
        if name not in self.CLIPPED:
            return fn

        def clipped(*a, **k):
            if self._clip is None:
                return fn(*a, **k)
            ctx, (l, t, r, b) = self._ctx, self._clip
            W, H = ctx.width, ctx.height
            before = ctx.get_buffer_np().reshape(H, W, 4).copy()
            out = fn(*a, **k)
            after = ctx.get_buffer_np().reshape(H, W, 4)
            changed = (before.view(np.uint64) != after.view(np.uint64)).any(axis=2)
            changed[t:b, l:r] = False   # inside the rect the draw stands
            for j, i in zip(*np.nonzero(changed)):
                ctx.set_pixel(int(i), int(j), *before[j, i])
            return out

        return clipped


def make_clip_case(seed):
    def run(R, image_rgba, native=True):
        w, h = [(160, 90), (97, 61), (128, 72)][seed % 3]
        ctx = R.RenderContext(w, h, True)   # RGBA: SetPixel restores a pixel exactly (the 3-channel store spills, cpp:510)
        tex = tiny_textures(R, image_rgba)
        streams.stream_clip(ctx if native else ClipEmulated(ctx), tex, seed)
        return digest(ctx)

    return run


def clip_cases():
    return [(f"clip_{s}", make_clip_case(s)) for s in range(8)]


# ---- N-gon fill (extension X3), pinned to the reference's own DrawLine machinery ------------------------------------------
# oracle/_ref/libNativeCPURenderer_polygon.so = the unmodified reference translation unit + one entry point that runs DrawLine's
# pixel loop through the reference's own pointInPolygon / ApplyPixel on the caller's points (oracle/ref_polygon_shim.cpp).
def make_polygon_case(seed):
    def run(R, image_rgba):
        # RGBA canvases only: on a 3-channel canvas the reference's SetColor writes 8 bytes past its buffer (cpp:510), which is
        # harmless often enough for the committed random cases but not in a process that runs this many of them
        w, h = [(160, 90), (97, 61), (128, 72)][seed % 3]
        ctx = R.RenderContext(w, h, True)
        streams.stream_polygons(ctx, tiny_textures(R, image_rgba), seed)
        return digest(ctx)

    return run


def polygon_cases():
    return [(f"polygon_{s}", make_polygon_case(s)) for s in range(8)]


# ---- perspective quads (extension X4): the projective map is this repo's spec; everything after it is reference code --------
def make_perspective_case(seed):
    def run(R, image_rgba):
        w, h = [(160, 90), (97, 61), (128, 72)][seed % 3]
        ctx = R.RenderContext(w, h, True)   # RGBA only, see make_polygon_case
        streams.stream_perspective(ctx, tiny_textures(R, image_rgba), seed)
        return digest(ctx)

    return run


def perspective_cases():
    return [(f"perspective_{s}", make_perspective_case(s)) for s in range(6)]


def all_cases(reference_abi_only: bool = False):
    cases = [("k1", case_k1), ("k2", case_k2), ("k3", case_k3), ("k4", case_k4), ("k5", case_k5), ("k6", case_k6),
             ("c2_small", case_c2_small), ("c3_small", case_c3_small), ("c4_small", case_c4_small),
             ("canvas_textures", case_canvas_textures)]
    cases += [(f"random_{s}", make_random_case(s)) for s in range(18)]
    if not reference_abi_only:
        cases += [(f"random_ap_{s}", make_random_case(100 + s, use_apply_pixel=True)) for s in range(6)]
    return cases
