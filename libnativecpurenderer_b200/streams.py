"""Seeded synthetic command streams.

Every generator drives a *sink* with the ``RenderContext`` drawing interface — a live context of any
loaded library (``binding.Renderer``) or a ``trace.TraceRecorder`` — so the same stream reaches the
product, the reference build and the C restatement.  ``textures`` are ``Texture`` objects for a live
sink and ``TexSlot`` objects for a recorder; generators only use ``.width`` / ``.height`` of them.

Texture *content* for the synthetic configs comes from ``make_*`` helpers (numpy, seeded); the only
image file involved is the reference's 128x128 ``test_files/image.png``, committed as raw RGBA in
``tests/golden/image_rgba.npy`` by ``tests/golden/make_golden.py``.
"""
from __future__ import annotations

import math
import random

import numpy as np

TWO_PI = 2 * math.pi


# --------------------------------------------------------------------------------------------- textures
def make_noise_texture(size: int, seed: int) -> np.ndarray:
    """(size, size, 4) uint8: random RGB with a radial alpha falloff (C2's quads)."""
    rs = np.random.RandomState(seed)
    img = rs.randint(0, 256, (size, size, 4)).astype(np.uint8)
    yy, xx = np.mgrid[0:size, 0:size]
    rad = np.hypot(xx - size / 2, yy - size / 2) / (size / 2)
    img[..., 3] = np.clip(255 * (1.25 - rad), 0, 255).astype(np.uint8)
    return img


def make_atlas(cells: int = 8, cell: int = 256, seed: int = 3) -> np.ndarray:
    """(cells*cell, cells*cell, 4) uint8 atlas: each cell a tinted noise sprite with soft alpha (C3)."""
    rs = np.random.RandomState(seed)
    size = cells * cell
    img = np.empty((size, size, 4), dtype=np.uint8)
    yy, xx = np.mgrid[0:cell, 0:cell]
    rad = np.hypot(xx - cell / 2, yy - cell / 2) / (cell / 2)
    alpha = np.clip(255 * (1.2 - rad), 0, 255).astype(np.uint8)
    for cy in range(cells):
        for cx in range(cells):
            tint = rs.randint(64, 256, 3)
            noise = rs.randint(0, 64, (cell, cell, 3))
            blk = img[cy * cell:(cy + 1) * cell, cx * cell:(cx + 1) * cell]
            blk[..., :3] = np.clip(tint[None, None, :] - noise, 0, 255).astype(np.uint8)
            blk[..., 3] = alpha
    return img


def make_chart_textures(seed: int = 4) -> list[np.ndarray]:
    """Textures of the milrenderer-shaped stream: [0] note, [1] hold body, [2] line head, [3] hit effect 512^2."""
    rs = np.random.RandomState(seed)

    def sprite(w, h, tint):
        img = np.empty((h, w, 4), dtype=np.uint8)
        img[..., :3] = np.clip(np.array(tint)[None, None, :] + rs.randint(-20, 20, (h, w, 3)), 0, 255)
        yy, xx = np.mgrid[0:h, 0:w]
        edge = np.minimum(np.minimum(xx, w - 1 - xx), np.minimum(yy, h - 1 - yy))
        img[..., 3] = np.clip(edge * 40, 0, 255)
        return img

    note = sprite(256, 64, (120, 200, 250))
    hold = sprite(256, 256, (250, 220, 120))
    head = sprite(64, 64, (255, 255, 255))
    fx = np.zeros((512, 512, 4), dtype=np.uint8)
    fx[..., :3] = (150, 144, 253)
    yy, xx = np.mgrid[0:512, 0:512]
    ring = np.abs(np.hypot(xx - 256, yy - 256) - 180)
    fx[..., 3] = np.where(ring < 40, (rs.rand(512, 512) > 0.5) * 255, 0).astype(np.uint8)
    return [note, hold, head, fx]


# --------------------------------------------------------------------------------------------- known answers
def stream_k1(ctx, tex, n: int = 1000, seed: int = 0) -> None:
    """SURVEY.md §8c K1 == BASELINE config 1: ``image.png`` as n rotated/scaled/alpha quads on 1920x1080 RGBA."""
    w, h = ctx.width, ctx.height
    ctx.set_color(0, 0, 0, 1)
    rng = random.Random(seed)
    for _ in range(n):
        ctx.save_state()
        ctx.translate(rng.uniform(0, w), rng.uniform(0, h))
        ctx.rotate(rng.uniform(0, TWO_PI))
        s = rng.uniform(0.5, 2.0)
        ctx.scale(s, s)
        ctx.apply_color_transform(1, 1, 1, rng.uniform(0.2, 0.9))
        ctx.draw_texture(tex, -64, -64, 128, 128)
        ctx.restore_state()


def stream_k2_frame(ctx, tex16, i: int) -> None:
    """Frame i of the reference binding's smoke loop (pyb:704-716); the caller has applied scale(.25,.25) once."""
    t = i / 60
    ctx.set_color(1, 1, 1, 1)
    ctx.save_state()
    ctx.apply_color_transform(t % 1, (t + 1.4) % 1, (t + 2.8) % 1, 1)
    w = 768 * (1 + math.sin(t * 2 * math.pi) / 4)
    h = 768 * (1 + math.cos(t * 3 * math.pi) / 4)
    ctx.draw_texture(tex16, w * 1.5 / 2, h * 1.3 / 2, w, h)
    ctx.draw_line(w * 0.1, h * 0.1, w, h, (w + h) / 300, 0, 1, 0, 1)
    ctx.draw_circle(w * 0.3, h * 0.3, 100, 1, 1, 0, 0.4)
    ctx.draw_rect(w * 0.6, h * 0.6, w * 0.1, h * 0.1, 0, 1, 0, 0.4)
    ctx.restore_state()


def k6_texture() -> np.ndarray:
    return np.random.RandomState(6).randint(0, 256, (40, 64, 4)).astype(np.uint8)


def stream_k6(ctx, tex) -> None:
    """320x180 RGB canvas (as milrenderer uses), fill + gradient + 40 split-texture draws."""
    W, H = ctx.width, ctx.height
    ctx.set_color(.25, .25, .25, .25)
    ctx.fill_color(1, .5, 0, .3)
    ctx.draw_vertical_grd(0, H * .6, W, H * .4, 0, 0, 0, 0.0, 0, 0, 1, 0.9)
    rng = random.Random(6)
    for _ in range(40):
        ctx.save_state()
        tx, ty, an, s, al = rng.uniform(0, W), rng.uniform(0, H), rng.uniform(0, 6.28), rng.uniform(.3, 1.5), rng.uniform(.2, 1)
        ctx.translate(tx, ty)
        ctx.rotate(an)
        ctx.scale(s, s * 1.3)
        ctx.apply_color_transform(1, .9, .8, al)
        uv = (rng.uniform(0, .4), rng.uniform(.6, 1), rng.uniform(0, .4), rng.uniform(.6, 1))
        ctx.draw_splitted_texture(tex, -30, -20, 60, 40, *uv)
        ctx.restore_state()


# --------------------------------------------------------------------------------------------- benchmark configs
C2_TEXTURE_SIZES = (64, 128, 256, 512)


def make_c2_textures() -> list[np.ndarray]:
    return [make_noise_texture(s, 20 + k) for k, s in enumerate(C2_TEXTURE_SIZES)]


def stream_c2(ctx, textures, n: int = 20000, seed: int = 2) -> None:
    """BASELINE config 2 restricted to the reference ABI: n mixed draws on an RGBA canvas —
    60 % DrawTexture, 20 % DrawSplittedTexture, 10 % DrawRect, 5 % DrawVerticalGrd, 3 % DrawCircle, 2 % DrawLine."""
    W, H = ctx.width, ctx.height
    rng = random.Random(seed)
    ctx.set_color(.1, .1, .1, 1)
    for _ in range(n):
        kind = rng.random()
        ctx.save_state()
        ctx.translate(rng.uniform(0, W), rng.uniform(0, H))
        ctx.rotate(rng.uniform(0, TWO_PI))
        s = rng.uniform(0.1, 0.6)
        ctx.scale(s, s)
        alpha = 1.0 if rng.random() < 0.2 else rng.uniform(0.1, 1.0)
        ctx.apply_color_transform(1, 1, 1, alpha)
        if kind < 0.60:
            tex = textures[rng.randrange(len(textures))]
            ctx.draw_texture(tex, -tex.width / 2, -tex.height / 2, tex.width, tex.height)
        elif kind < 0.80:
            tex = textures[rng.randrange(len(textures))]
            u0, v0 = rng.uniform(0, .5), rng.uniform(0, .5)
            ctx.draw_splitted_texture(tex, -tex.width / 2, -tex.height / 2, tex.width, tex.height,
                                      u0, u0 + rng.uniform(.2, .5), v0, v0 + rng.uniform(.2, .5))
        elif kind < 0.90:
            ctx.draw_rect(-100, -60, 200, 120, rng.random(), rng.random(), rng.random(), rng.uniform(.2, 1))
        elif kind < 0.95:
            ctx.draw_vertical_grd(-120, -120, 240, 240, rng.random(), rng.random(), rng.random(), rng.uniform(0, .5),
                                  rng.random(), rng.random(), rng.random(), rng.uniform(.5, 1))
        elif kind < 0.98:
            ctx.draw_circle(0, 0, rng.uniform(40, 160), rng.random(), rng.random(), rng.random(), rng.uniform(.2, 1))
        else:
            ctx.draw_line(-300, rng.uniform(-50, 50), 300, rng.uniform(-50, 50), rng.uniform(4, 30),
                          rng.random(), rng.random(), rng.random(), rng.uniform(.3, 1))
        ctx.restore_state()


def stream_c3(ctx, atlas, n: int = 50000, seed: int = 3, cells: int = 8) -> None:
    """BASELINE config 3, affine variant: n DrawSplittedTexture sprites from an 8x8-cell atlas
    on a 3840x2160 RGBA canvas."""
    W, H = ctx.width, ctx.height
    rng = random.Random(seed)
    ctx.set_color(0, 0, 0, 1)
    cell = atlas.width / cells
    for _ in range(n):
        cx, cy = rng.randrange(cells), rng.randrange(cells)
        ctx.save_state()
        ctx.translate(rng.uniform(0, W), rng.uniform(0, H))
        ctx.rotate(rng.uniform(0, TWO_PI))
        s = rng.uniform(0.05, 0.5)
        ctx.scale(s, s)
        ctx.apply_color_transform(1, 1, 1, rng.uniform(0.2, 1.0))
        ctx.draw_splitted_texture(atlas, -cell / 2, -cell / 2, cell, cell,
                                  cx / cells, (cx + 1) / cells, cy / cells, (cy + 1) / cells)
        ctx.restore_state()


def stream_c4_frame(ctx, bg, textures, frame: int, n_notes: int = 1500, n_fx: int = 100, seed: int = 4) -> None:
    """One frame of the milrenderer-shaped chart (call mix of mil:865-1038): clear, full-screen
    background (identity path), dim, 4 gradients, 12 judgement lines (head sprite + DrawLine body), ~n_notes notes
    (3/4 taps via DrawTexture, 1/4 holds via 3x DrawSplittedTexture), ~n_fx hit effects of a 512^2 texture.
    Positions advance deterministically with ``frame``.  Sizes are relative to a 1080-line canvas."""
    W, H = ctx.width, ctx.height
    k = H / 1080.0
    note, hold, head, fx = textures
    rng = random.Random(seed)  # same per-object parameters every frame; motion comes from `frame`
    t = frame / 60.0
    ctx.set_color(0, 0, 0, 0)
    ctx.draw_texture(bg, W / 2 - bg.width / 2, H / 2 - bg.height / 2, bg.width, bg.height)
    ctx.fill_color(0, 0, 0, .6)
    for q in range(4):
        ctx.draw_vertical_grd(q * W / 4, 0, W / 4, H * .25, 0, 0, 0, .8, 0, 0, 0, 0.0)
    lines = []
    for li in range(12):
        ang = rng.uniform(-.4, .4) + .15 * math.sin(t * rng.uniform(.2, 1.0) + li)
        cx = W * (li + .5) / 12 + 40 * k * math.sin(t * .7 + li)
        cy = H * rng.uniform(.55, .85)
        lines.append((cx, cy, ang))
        ctx.save_state()
        ctx.translate(cx, cy)
        ctx.rotate(ang)
        ctx.apply_color_transform(1, 1, 1, .9)
        ctx.draw_line(-W * .12, 0, W * .12, 0, 6 * k, 1, 1, 1, .85)
        ctx.draw_texture(head, -32 * k, -32 * k, 64 * k, 64 * k)
        ctx.restore_state()
    for ni in range(n_notes):
        cx, cy, ang = lines[ni % 12]
        speed = rng.uniform(300, 900) * k
        phase = rng.uniform(0, 4)
        lane = rng.uniform(-W * .1, W * .1)
        is_hold = rng.random() < .25
        dist = ((phase - t) % 4.0) * speed * .35
        ctx.save_state()
        ctx.apply_color_transform(1, 1, 1, rng.uniform(.6, 1.0))
        ctx.translate(cx, cy)
        ctx.rotate(ang)
        ctx.translate(lane, -dist)
        nw, nh = 120 * k, 30 * k
        if not is_hold:
            ctx.draw_texture(note, -nw / 2, -nh / 2, nw, nh)
        else:
            body = rng.uniform(60, 260) * k
            ctx.draw_splitted_texture(hold, -nw / 2, -body - nh, nw, nh, 0, 1, 0, .2)
            ctx.draw_splitted_texture(hold, -nw / 2, -body, nw, body, 0, 1, .2, .8)
            ctx.draw_splitted_texture(hold, -nw / 2, 0, nw, nh, 0, 1, .8, 1)
        ctx.restore_state()
    for fi in range(n_fx):
        cx, cy, ang = lines[fi % 12]
        life = ((t * 2 + rng.uniform(0, 1)) % 1.0)
        size = (180 + 120 * life) * k
        ox = rng.uniform(-W * .1, W * .1)
        ctx.save_state()
        ctx.set_transform(math.cos(ang), math.sin(ang), -math.sin(ang), math.cos(ang),
                          cx + ox * math.cos(ang), cy + ox * math.sin(ang))
        ctx.apply_color_transform(1, 1, 1, 1 - life)
        ctx.draw_texture(fx, -size / 2, -size / 2, size, size)
        ctx.restore_state()


# --------------------------------------------------------------------------------------------- randomised parity streams
def stream_random(ctx, textures, seed: int, n: int = 60, use_apply_pixel: bool = False) -> None:
    """Seeded mix of every reference-ABI draw/state call, biased toward the edge cases of SURVEY.md §8a-Q:
    "no transform" matrices that are not the identity (quirk 1), fractional and negative origins (2, 3),
    off-canvas and degenerate sizes (10), 1:1 sampling of the last texel row/column (4), alpha exactly 1 (6),
    nested save/restore, non-invertible matrices (inv_det = 1e9, cpp:484)."""
    W, H = ctx.width, ctx.height
    rng = random.Random(seed)
    u = rng.uniform

    def colour():
        return (rng.choice([0.0, 1.0, u(0, 1)]), u(0, 1), u(0, 1), rng.choice([1.0, u(0, 1), u(0, 1)]))

    if rng.random() < 0.7:
        v = u(0, 1)
        ctx.set_color(*((v, v, v, v) if rng.random() < 0.5 else (u(0, 1), u(0, 1), u(0, 1), u(0, 1))))
    depth = 0
    for _ in range(n):
        op = rng.random()
        if op < 0.10:
            ctx.save_state()
            depth += 1
        elif op < 0.18:
            ctx.restore_state()   # may hit an empty stack on purpose (cpp:293)
            depth = max(0, depth - 1)
        elif op < 0.26:
            ctx.translate(u(-W * .3, W * 1.1), u(-H * .3, H * 1.1))
        elif op < 0.32:
            ctx.rotate(u(-7, 7))
        elif op < 0.38:
            s = rng.choice([u(.2, 3), u(.2, 3), -u(.5, 1.5), 0.0 if rng.random() < .1 else 1.0])
            ctx.scale(s, rng.choice([s, u(.2, 3)]))
        elif op < 0.42:
            ctx.set_transform(*rng.choice([
                (1, 0, 0, 1, 0, 0),
                (1, 0, 0, 1, -u(0, 9), -u(0, 9)),          # negative translation: still "no transform" (quirk 1)
                (u(.1, .9), 0, 0, u(.1, .9), 0, 0),        # down-scale: "no transform" as well
                (u(.5, 2), u(-1, 1), u(-1, 1), u(.5, 2), u(0, W), u(0, H)),
                (1, 2, 2, 4, u(0, W), u(0, H)),            # singular
            ]))
        elif op < 0.47:
            ctx.apply_color_transform(u(.3, 1.2), u(.3, 1.2), u(.3, 1.2), rng.choice([1.0, u(.2, 1.1)]))
        elif op < 0.50:
            ctx.set_color_transform(1, 1, 1, 1)
        elif op < 0.62:
            tex = rng.choice(textures)
            x, y = rng.choice([(0, 0), (u(-40, 40), u(-40, 40)), (2.5, 2.5)])
            w, h = rng.choice([(tex.width, tex.height), (u(1, 90), u(1, 90)), (-u(1, 30), u(1, 30)), (0, 5)])
            ctx.draw_texture(tex, x, y, w, h)
        elif op < 0.72:
            tex = rng.choice(textures)
            us, vs = u(0, .6), u(0, .6)
            ctx.draw_splitted_texture(tex, u(-40, 40), u(-40, 40), u(1, 80), u(1, 80), us, us + u(.1, .6), vs, vs + u(.1, .6))
        elif op < 0.80:
            ctx.draw_rect(u(-30, 60), u(-30, 60), rng.choice([u(1, 70), 4, -3]), u(1, 70), *colour())
        elif op < 0.85:
            ctx.draw_vertical_grd(u(-30, 60), u(-30, 60), u(1, 70), u(1, 70), *colour(), *colour())
        elif op < 0.90:
            ctx.draw_circle(u(-20, 60), u(-20, 60), rng.choice([u(1, 40), 0, -2]), *colour())
        elif op < 0.94:
            ctx.draw_line(u(-20, 80), u(-20, 80), u(-20, 80), u(-20, 80), rng.choice([u(.5, 12), 0]), *colour())
        elif op < 0.96:
            ctx.fill_color(*colour())
        elif op < 0.98:
            ctx.set_pixel(rng.randrange(-2, W + 2), rng.randrange(-2, H + 2), *colour())
        elif use_apply_pixel:
            ctx.apply_pixel(rng.randrange(-2, W + 2), rng.randrange(-2, H + 2), *colour())
    for _ in range(depth):
        ctx.restore_state()


# --------------------------------------------------------------------------------------------- extensions (pinned one by one in tests/cases.py)
def stream_extensions(ctx, textures, seed: int, n: int = 60) -> None:
    """The entry points BASELINE's configs name that the reference does not have (include/ncr_b200.h §2): clip rects,
    bilinear sampling, N-gon fill, perspective quads — mixed with reference-ABI draws.  Product vs C restatement only."""
    W, H = ctx.width, ctx.height
    rng = random.Random(seed)
    u = rng.uniform
    ctx.set_color(.15, .15, .2, 1)
    for k in range(n):
        ctx.save_state()
        ctx.translate(u(0, W), u(0, H))
        ctx.rotate(u(0, TWO_PI))
        s = u(.3, 1.4)
        ctx.scale(s, s)
        ctx.apply_color_transform(1, u(.6, 1), 1, rng.choice([1.0, u(.2, 1)]))
        op = rng.random()
        if op < .15:
            ctx.set_clip_rect(int(u(0, W * .6)), int(u(0, H * .6)), int(u(8, W * .7)), int(u(8, H * .7)))
        elif op < .25:
            ctx.clear_clip_rect()
        elif op < .35:
            ctx.set_sampling(rng.choice([0, 1]))
        elif op < .55:
            pts = [(u(-60, 60), u(-60, 60)) for _ in range(rng.choice([3, 5, 7]))]
            ctx.fill_polygon(pts, u(0, 1), u(0, 1), u(0, 1), rng.choice([1.0, u(.2, 1)]))
        elif op < .70:
            tex = rng.choice(textures)
            # inverse homography: a mild projective warp around the translate point
            tx_, ty_ = u(0, W), u(0, H)
            c_, s_ = math.cos(u(0, TWO_PI)), math.sin(u(0, TWO_PI))
            k_ = u(.5, 2)
            hinv = (c_ * k_, s_ * k_, -(c_ * k_ * tx_ + s_ * k_ * ty_), -s_ * k_, c_ * k_, (s_ * k_ * tx_ - c_ * k_ * ty_),
                    u(-1e-3, 1e-3), u(-1e-3, 1e-3), 1.0)
            ctx.draw_texture_perspective(tex, hinv, -40, -30, 80, 60)
        elif op < .85:
            tex = rng.choice(textures)
            ctx.draw_texture(tex, -tex.width / 2, -tex.height / 2, tex.width, tex.height)
        elif op < .92:
            ctx.draw_line(-50, u(-20, 20), 50, u(-20, 20), u(1, 9), u(0, 1), u(0, 1), u(0, 1), u(.3, 1))
        elif op < .96:
            ctx.fill_color(u(0, 1), u(0, 1), u(0, 1), u(.05, .3))
        else:
            tex = rng.choice(textures)
            ctx.draw_splitted_texture(tex, -30, -30, 60, 60, 0.1, 0.9, 0.2, 0.8)
        ctx.restore_state()
    ctx.clear_clip_rect()
    ctx.set_sampling(0)


def stream_clip(ctx, textures, seed: int, n: int = 70) -> None:
    """Reference-ABI draws of every kind under clip rects that come and go (tests/cases.py pins the clip extension with it:
    the unmodified reference, drawing unclipped and having the outside pixels put back, must give the same canvas)."""
    W, H = ctx.width, ctx.height
    rng = random.Random(1000 + seed)
    u = rng.uniform
    ctx.set_color(.2, .25, .15, 1)
    for k in range(n):
        op = rng.random()
        if op < .12:
            ctx.set_clip_rect(int(u(-10, W * .7)), int(u(-10, H * .7)), int(u(1, W * .8)), int(u(1, H * .8)))
            continue
        if op < .18:
            ctx.clear_clip_rect()
            continue
        ctx.save_state()
        ctx.translate(u(0, W), u(0, H))
        if rng.random() < .8:
            ctx.rotate(u(0, TWO_PI))
            s = u(.3, 1.5)
            ctx.scale(s, s)
        ctx.apply_color_transform(u(.5, 1.1), 1, u(.5, 1), rng.choice([1.0, u(.2, 1)]))
        if op < .40:
            tex = rng.choice(textures)
            ctx.draw_texture(tex, -tex.width / 2, -tex.height / 2, tex.width, tex.height)
        elif op < .52:
            tex = rng.choice(textures)
            ctx.draw_splitted_texture(tex, -30, -20, 60, 40, u(0, .4), u(.5, 1), u(0, .4), u(.5, 1))
        elif op < .66:
            ctx.draw_rect(-40, -25, 80, 50, u(0, 1), u(0, 1), u(0, 1), rng.choice([1.0, u(.2, 1)]))
        elif op < .76:
            ctx.draw_vertical_grd(-40, -40, 80, 80, u(0, 1), u(0, 1), u(0, 1), u(0, .5), u(0, 1), u(0, 1), u(0, 1), u(.5, 1))
        elif op < .86:
            ctx.draw_circle(0, 0, u(8, 45), u(0, 1), u(0, 1), u(0, 1), u(.2, 1))
        elif op < .95:
            ctx.draw_line(-50, u(-20, 20), 50, u(-20, 20), u(1, 12), u(0, 1), u(0, 1), u(0, 1), u(.3, 1))
        else:
            ctx.fill_color(u(0, 1), u(0, 1), u(0, 1), u(.05, .4))
        ctx.restore_state()
    ctx.clear_clip_rect()


def stream_polygons(ctx, textures, seed: int, n: int = 60) -> None:
    """N-gon fills (convex, concave and self-intersecting: the even-odd rule shows) under rotations, scales and colour transforms,
    mixed with reference-ABI draws (tests/cases.py pins the polygon extension with it)."""
    W, H = ctx.width, ctx.height
    rng = random.Random(2000 + seed)
    u = rng.uniform
    ctx.set_color(.1, .12, .2, 1)
    for k in range(n):
        ctx.save_state()
        ctx.translate(u(0, W), u(0, H))
        ctx.rotate(u(0, TWO_PI))
        s = u(.3, 1.6)
        ctx.scale(s, s * u(.6, 1.4))
        ctx.apply_color_transform(u(.5, 1.1), u(.5, 1), 1, rng.choice([1.0, u(.2, 1)]))
        op = rng.random()
        if op < .30:     # convex-ish m-gon
            m = rng.choice([3, 4, 5, 6, 8])
            pts = [(50 * math.cos(TWO_PI * j / m) * u(.6, 1), 50 * math.sin(TWO_PI * j / m) * u(.6, 1)) for j in range(m)]
        elif op < .55:   # random points: concave / self-intersecting
            pts = [(u(-60, 60), u(-60, 60)) for _ in range(rng.choice([3, 5, 7, 9]))]
        elif op < .65:   # star polygon {5/2}: the pentagon in the middle is OUTSIDE under the even-odd rule
            pts = [(45 * math.cos(TWO_PI * 2 * j / 5), 45 * math.sin(TWO_PI * 2 * j / 5)) for j in range(5)]
        elif op < .72:   # degenerate: repeated points, horizontal edges, a point far outside the canvas
            pts = [(-20, -10), (30, -10), (30, -10), (1e4, 40), (-20, 25)]
        else:
            pts = None
        if pts is not None:
            ctx.fill_polygon(pts, u(0, 1), u(0, 1), u(0, 1), rng.choice([1.0, u(.2, 1)]))
        elif op < .86:
            tex = rng.choice(textures)
            ctx.draw_texture(tex, -tex.width / 2, -tex.height / 2, tex.width, tex.height)
        elif op < .94:
            ctx.draw_line(-50, u(-20, 20), 50, u(-20, 20), u(1, 9), u(0, 1), u(0, 1), u(0, 1), u(.3, 1))
        else:
            ctx.draw_rect(-30, -20, 60, 40, u(0, 1), u(0, 1), u(0, 1), u(.2, 1))
        ctx.restore_state()


def stream_perspective(ctx, textures, seed: int, n: int = 40) -> None:
    """Perspective-warped quads (mild to strong projective terms, some with the horizon inside the canvas so that hw <= 0 pixels
    exist) mixed with reference-ABI draws; colour transforms with alpha exactly 1 and below."""
    W, H = ctx.width, ctx.height
    rng = random.Random(3000 + seed)
    u = rng.uniform
    ctx.set_color(.12, .1, .18, 1)
    for k in range(n):
        ctx.save_state()
        ctx.apply_color_transform(u(.5, 1.1), 1, u(.5, 1), rng.choice([1.0, u(.2, 1)]))
        op = rng.random()
        tex = rng.choice(textures)
        if op < .7:
            tx_, ty_ = u(0, W), u(0, H)
            ang = u(0, TWO_PI)
            c_, s_ = math.cos(ang), math.sin(ang)
            k_ = u(.5, 2.5)
            p = rng.choice([1e-3, 4e-3, 2e-2])   # 2e-2: the line hw = 0 crosses a 160-px canvas
            hinv = (c_ * k_, s_ * k_, -(c_ * k_ * tx_ + s_ * k_ * ty_), -s_ * k_, c_ * k_, (s_ * k_ * tx_ - c_ * k_ * ty_),
                    u(-p, p), u(-p, p), 1.0)
            ctx.draw_texture_perspective(tex, hinv, -40, -30, 80, 60)
        elif op < .85:
            ctx.translate(u(0, W), u(0, H))
            ctx.rotate(u(0, TWO_PI))
            ctx.draw_texture(tex, -tex.width / 2, -tex.height / 2, tex.width, tex.height)
        else:
            ctx.translate(u(0, W), u(0, H))
            ctx.draw_rect(-30, -20, 60, 40, u(0, 1), u(0, 1), u(0, 1), u(.2, 1))
        ctx.restore_state()


def stream_c2x(ctx, textures, n: int = 20000, seed: int = 2) -> None:
    """BASELINE config 2 as written, extensions included (product only): the C2 mix with bilinear sampling on half of the
    textured draws, N-gon fills in place of rects, and a clip rect that changes every 500 draws."""
    W, H = ctx.width, ctx.height
    rng = random.Random(seed)
    ctx.set_color(.1, .1, .1, 1)
    for k in range(n):
        if k % 500 == 0:
            if (k // 500) % 3 == 2:
                ctx.clear_clip_rect()
            else:
                ctx.set_clip_rect(int(rng.uniform(0, W * .3)), int(rng.uniform(0, H * .3)), int(W * .7), int(H * .7))
        kind = rng.random()
        ctx.save_state()
        ctx.translate(rng.uniform(0, W), rng.uniform(0, H))
        ctx.rotate(rng.uniform(0, TWO_PI))
        s = rng.uniform(0.1, 0.6)
        ctx.scale(s, s)
        ctx.apply_color_transform(1, 1, 1, 1.0 if rng.random() < 0.2 else rng.uniform(0.1, 1.0))
        if kind < 0.80:
            tex = textures[rng.randrange(len(textures))]
            ctx.set_sampling(1 if rng.random() < 0.5 else 0)
            if kind < 0.60:
                ctx.draw_texture(tex, -tex.width / 2, -tex.height / 2, tex.width, tex.height)
            else:
                u0, v0 = rng.uniform(0, .5), rng.uniform(0, .5)
                ctx.draw_splitted_texture(tex, -tex.width / 2, -tex.height / 2, tex.width, tex.height,
                                          u0, u0 + rng.uniform(.2, .5), v0, v0 + rng.uniform(.2, .5))
        elif kind < 0.90:
            m = rng.choice([3, 5, 6])
            pts = [(120 * math.cos(TWO_PI * j / m) * rng.uniform(.6, 1), 120 * math.sin(TWO_PI * j / m) * rng.uniform(.6, 1))
                   for j in range(m)]
            ctx.fill_polygon(pts, rng.random(), rng.random(), rng.random(), rng.uniform(.2, 1))
        elif kind < 0.95:
            ctx.draw_vertical_grd(-120, -120, 240, 240, rng.random(), rng.random(), rng.random(), rng.uniform(0, .5),
                                  rng.random(), rng.random(), rng.random(), rng.uniform(.5, 1))
        elif kind < 0.98:
            ctx.draw_circle(0, 0, rng.uniform(40, 160), rng.random(), rng.random(), rng.random(), rng.uniform(.2, 1))
        else:
            ctx.draw_line(-300, rng.uniform(-50, 50), 300, rng.uniform(-50, 50), rng.uniform(4, 30),
                          rng.random(), rng.random(), rng.random(), rng.uniform(.3, 1))
        ctx.restore_state()
    ctx.clear_clip_rect()
    ctx.set_sampling(0)


def stream_c3p(ctx, atlas, n: int = 50000, seed: int = 3, cells: int = 8) -> None:
    """BASELINE config 3, perspective variant (product only): n perspective-warped quads textured from the 2048^2 atlas.
    The inverse homography maps canvas pixels into a 256x256 source square; like DrawTexture, the square is mapped onto the
    whole texture, so every quad samples the full 16.8 MB atlas 8x minified (a texture-cache / L2 stress, as the config asks)."""
    W, H = ctx.width, ctx.height
    rng = random.Random(seed)
    ctx.set_color(0, 0, 0, 1)
    cell = atlas.width / cells
    for _ in range(n):
        cx, cy = rng.randrange(cells), rng.randrange(cells)
        tx_, ty_ = rng.uniform(0, W), rng.uniform(0, H)
        ang = rng.uniform(0, TWO_PI)
        k_ = 1.0 / rng.uniform(0.05, 0.5)
        c_, s_ = math.cos(ang) * k_, math.sin(ang) * k_
        ox, oy = (cx + .5) * cell, (cy + .5) * cell   # the source square's centre
        g, hh = rng.uniform(-2e-4, 2e-4), rng.uniform(-2e-4, 2e-4)   # mild projective term
        w0 = 1.0 - g * tx_ - hh * ty_
        hinv = (c_ + ox * g, s_ + ox * hh, -(c_ * tx_ + s_ * ty_) + ox * w0,
                -s_ + oy * g, c_ + oy * hh, (s_ * tx_ - c_ * ty_) + oy * w0,
                g, hh, w0)
        ctx.save_state()
        ctx.apply_color_transform(1, 1, 1, rng.uniform(0.2, 1.0))
        ctx.draw_texture_perspective(atlas, hinv, cx * cell, cy * cell, cell, cell)
        ctx.restore_state()
