"""GPU suite: the CUDA path, driven through the C ABI, against (1) the committed digests of the unmodified
reference build, (2) the C restatement run live on the same inputs, (3) size-independent properties at the
benchmark's full sizes.  Bit-exact everywhere (RGBA8/RGB8 readback and the f64 canvas)."""
import ctypes
import hashlib

import numpy as np
import pytest

import cases
from libnativecpurenderer_b200 import streams, trace

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,fn", cases.all_cases(reference_abi_only=True), ids=lambda v: v if isinstance(v, str) else "")
def test_product_matches_reference_golden(name, fn, gpu, golden, image_rgba):
    assert fn(gpu, image_rgba) == golden[name]


@pytest.mark.parametrize("name,fn", [c for c in cases.all_cases() if c[0].startswith("random_ap_")],
                         ids=lambda v: v if isinstance(v, str) else "")
def test_product_matches_reference_with_apply_pixel(name, fn, gpu, port, golden_apply_pixel, image_rgba):
    """The exported ApplyPixel (h:109; `inline` in the reference source, so its plain build has no such symbol) against digests of
    the reference's own function, exported by the shim build (oracle/ref_polygon_shim.cpp)."""
    got = fn(gpu, image_rgba)
    assert got == golden_apply_pixel[name]
    assert got == fn(port, image_rgba)


@pytest.mark.parametrize("seed", range(300, 312))
def test_product_matches_port_live(seed, gpu, port, image_rgba):
    run = cases.make_random_case(seed)
    assert run(gpu, image_rgba) == run(port, image_rgba)


def test_product_matches_reference_live(gpu, ref, image_rgba):
    for seed in range(400, 404):
        run = cases.make_random_case(seed)
        assert run(gpu, image_rgba) == run(ref, image_rgba)


def test_state_machine_is_bit_identical(gpu, port):
    import random

    for seed in range(4):
        outs = []
        for R in (gpu, port):
            rng = random.Random(seed)
            ctx = R.RenderContext(8, 8, True)
            for _ in range(60):
                k = rng.random()
                if k < .3:
                    ctx.translate(rng.uniform(-50, 50), rng.uniform(-50, 50))
                elif k < .6:
                    ctx.rotate(rng.uniform(-10, 10))
                elif k < .8:
                    ctx.scale(rng.uniform(.1, 3), rng.uniform(.1, 3))
                elif k < .9:
                    ctx.save_state()
                else:
                    ctx.restore_state()
            outs.append((ctx.get_transform(), ctx.get_inverse_transform()))
        assert outs[0] == outs[1]
    ctx = gpu.RenderContext(4, 4, True)
    assert ctx.restore_state() is False   # cpp:293


def test_trace_replay_equals_direct_calls(gpu, image_rgba):
    """NcrSubmitTrace (one FFI crossing) and csrc/ncr_replay (C calls) render what per-call ctypes renders."""
    from conftest import REPLAY_LIB

    tex_np = streams.make_c2_textures()
    tex = [gpu.Texture.from_numpy(t) for t in tex_np]
    direct = gpu.RenderContext(480, 270, True)
    streams.stream_c2(direct, tex, n=500)
    want = cases.digest(direct)

    rec = trace.TraceRecorder(480, 270, True)
    streams.stream_c2(rec, [trace.TexSlot(k, t.shape[1], t.shape[0]) for k, t in enumerate(tex_np)], n=500)
    arr = rec.as_array()
    a = gpu.RenderContext(480, 270, True)
    assert trace.submit_trace(a, arr, tex) == rec.n_records
    assert cases.digest(a) == want
    b = gpu.RenderContext(480, 270, True)
    trace.Replayer(REPLAY_LIB, gpu.path).run(b, arr, tex)
    assert cases.digest(b) == want


def test_readback_paths_agree(gpu, image_rgba):
    """GetBufferAsUInt8 fused into the composite, the standalone convert kernel, pinned and pageable destinations,
    GetBuffer and GetColor all describe the same canvas."""
    ctx = gpu.RenderContext(200, 120, False)
    tex = cases.tiny_textures(gpu, image_rgba)
    streams.stream_random(ctx, tex, 77, n=80)
    fused = bytes(ctx.get_buffer_as_uint8())          # flush with fused u8
    again = bytes(ctx.get_buffer_as_uint8())          # nothing pending: cached image
    f64 = ctx.get_buffer_np()
    ctx.flush()
    ctx.draw_rect(0, 0, 1, 1, 0, 0, 0, 0)             # invalidates the cached image without changing a pixel
    ctx.flush()
    converted = bytes(ctx.get_buffer_as_uint8())      # standalone convert kernel
    assert fused == again == converted
    n = ctx.get_buffer_size()
    pinned = gpu.lib.NcrAllocHost(n)
    assert pinned
    ctx.get_buffer_as_uint8_into(pinned)
    assert ctypes.string_at(pinned, n) == fused
    gpu.lib.NcrFreeHost(pinned)
    scaled = f64 * 255
    safe = np.where(np.abs(scaled) < 2 ** 31, scaled, 0)
    assert (np.trunc(safe).astype(np.int64) & 0xFF).astype(np.uint8).tobytes() == fused
    px = f64.reshape(120, 200, 3)
    assert ctx.get_color(17.9, 5.2)[:3] == tuple(px[5, 17])
    assert ctx.get_color(-3, 1e9)[:3] == tuple(px[119, 0])   # clamp, cpp:664-667


def test_shared_canvas_texture_reads_the_canvas_as_of_the_draw(gpu, port, image_rgba):
    outs = []
    for R in (gpu, port):
        src = R.RenderContext(32, 32, True)
        src.set_color(.2, .4, .6, 1)
        shared = src.as_texture_shared()
        dst = R.RenderContext(64, 64, True)
        dst.set_color(0, 0, 0, 1)
        dst.translate(1, 1)
        dst.rotate(.1)
        src.draw_circle(16, 16, 10, 1, 0, 0, .5)          # pending on src when the alias is drawn
        dst.draw_texture(shared, 0, 0, 40, 40)
        src.set_color(1, 1, 1, 1)                          # later change must not leak into the recorded draw
        dst.draw_texture(shared, 20, 20, 30, 30)
        outs.append(cases.digest(dst))
        assert (shared.width, shared.height) == (32, 32)
    assert outs[0] == outs[1]


def test_destroyed_texture_stays_valid_for_recorded_draws(gpu, port, image_rgba):
    """The reference never frees (cpp:356-360), so Python may drop a texture right after drawing it."""
    outs = []
    for R in (gpu, port):
        ctx = R.RenderContext(64, 64, True)
        ctx.set_color(0, 0, 0, 1)
        ctx.translate(3, 2)
        ctx.rotate(.2)
        t = R.Texture.from_numpy(image_rgba)
        ctx.draw_texture(t, 0, 0, 50, 50)
        del t
        outs.append(cases.digest(ctx))
    assert outs[0] == outs[1]


def test_resize_discards_pixels_and_keeps_state(gpu):
    ctx = gpu.RenderContext(16, 16, True)
    ctx.translate(2, 3)
    ctx.draw_rect(0, 0, 5, 5, 1, 1, 1, 1)
    ctx.resize(24, 8)
    assert ctx.get_buffer_size() == 24 * 8 * 4
    assert ctx.get_transform() == (1, 0, 0, 1, 2, 3)
    ctx.set_color(.5, .5, .5, .5)
    assert set(ctx.get_buffer_as_uint8()) == {127}


def test_empty_and_degenerate_inputs(gpu):
    ctx = gpu.RenderContext(0, 0, True)
    assert ctx.get_buffer_size() == 0 and bytes(ctx.get_buffer_as_uint8()) == b""
    ctx = gpu.RenderContext(1, 1, False)
    ctx.set_color(1, 0, 0, 1)
    ctx.flush()
    assert list(ctx.get_buffer_as_uint8()) == [255, 0, 0]
    ctx = gpu.RenderContext(33, 17, True)   # nothing recorded: readback of the zero-initialised canvas
    assert set(ctx.get_buffer_as_uint8()) == {0}
    ctx.draw_rect(5, 5, -1, 4, 1, 1, 1, 1)   # early-outs (cpp:853)
    ctx.draw_circle(5, 5, 0, 1, 1, 1, 1)
    ctx.draw_line(1, 1, 1, 1, 3, 1, 1, 1, 1)
    assert ctx.stats().n_cmds == 0
    assert set(ctx.get_buffer_as_uint8()) == {0}


# ---- full benchmark sizes: properties that do not need a CPU render -------------------------------------------
def _c2_full(gpu, n=20000):
    tex_np = streams.make_c2_textures()
    tex = [gpu.Texture.from_numpy(t) for t in tex_np]
    rec = trace.TraceRecorder(1920, 1080, True)
    streams.stream_c2(rec, [trace.TexSlot(k, t.shape[1], t.shape[0]) for k, t in enumerate(tex_np)], n=n)
    return tex, rec


def test_c2_full_size_is_deterministic_and_flush_invariant(gpu):
    """1080p, 20,000 draws: the image does not depend on how the stream is cut into flushes (one batch, or a flush
    after every 997 records), nor on re-execution from the resident command buffers."""
    tex, rec = _c2_full(gpu)
    arr = rec.as_array()
    one = gpu.RenderContext(1920, 1080, True)
    trace.submit_trace(one, arr, tex)
    want = hashlib.sha1(bytes(one.get_buffer_as_uint8())).hexdigest()
    ms = (ctypes.c_float * 8)()
    assert gpu.lib.NcrRerunLastFlush(one._ptr, 2, 1, ms) == 0
    assert hashlib.sha1(bytes(one.get_buffer_as_uint8())).hexdigest() == want

    # cut the same stream at record boundaries
    raw = arr.tobytes()
    offs, p, k = [0], 0, 0
    while p < len(raw):
        n = int.from_bytes(raw[p + 4:p + 8], "little")
        p += 8 + 8 * n
        k += 1
        if k % 997 == 0:
            offs.append(p)
    offs.append(len(raw))
    many = gpu.RenderContext(1920, 1080, True)
    for a, b in zip(offs, offs[1:]):
        if b > a:
            piece = np.frombuffer(raw[a:b], dtype=np.uint8).copy()
            buf = np.empty(len(piece) // 8 + 1, dtype=np.float64)
            buf.view(np.uint8)[:len(piece)] = piece
            trace.submit_trace(many, buf.view(np.uint8)[:len(piece)], tex)
            many.flush()
    assert hashlib.sha1(bytes(many.get_buffer_as_uint8())).hexdigest() == want


def test_c2_full_size_matches_port(gpu, port):
    """BASELINE config 2 at its full size (1080p RGBA, 20,000 mixed draws) against the C restatement, which needs
    a few seconds of one host core for it: bit-exact u8 and f64."""
    from conftest import REPLAY_LIB

    tex_np = streams.make_c2_textures()
    rec = trace.TraceRecorder(1920, 1080, True)
    streams.stream_c2(rec, [trace.TexSlot(k, t.shape[1], t.shape[0]) for k, t in enumerate(tex_np)])
    arr = rec.as_array()
    got = []
    for R in (gpu, port):
        ctx = R.RenderContext(1920, 1080, True)
        tex = [R.Texture.from_numpy(t) for t in tex_np]
        trace.Replayer(REPLAY_LIB, R.path).run(ctx, arr, tex)
        got.append(cases.digest(ctx))
    assert got[0] == got[1]


def test_c3_full_size_properties(gpu):
    """4K, 50,000 atlas sprites: deterministic across runs and flush-invariant (the atlas is 16.8 MB of RGBA8)."""
    atlas_np = streams.make_atlas()
    atlas = gpu.Texture.from_numpy(atlas_np)
    rec = trace.TraceRecorder(3840, 2160, True)
    streams.stream_c3(rec, trace.TexSlot(0, 2048, 2048), n=50000)
    arr = rec.as_array()
    hashes = []
    for _ in range(2):
        ctx = gpu.RenderContext(3840, 2160, True)
        trace.submit_trace(ctx, arr, [atlas])
        hashes.append(hashlib.sha1(bytes(ctx.get_buffer_as_uint8())).hexdigest())
        st = ctx.stats()
        assert st.n_cmds > 40000 and st.fine_entries > st.n_cmds
    assert hashes[0] == hashes[1]


# ---- bilinear extension: pinned to the reference's own (commented-out) four-tap sampler ---------------------------------------
@pytest.mark.parametrize("name,fn", cases.bilinear_cases(), ids=lambda v: v if isinstance(v, str) else "")
def test_bilinear_matches_the_references_own_four_tap_sampler(name, fn, gpu, golden_bilinear, image_rgba):
    """NcrSetSampling(ctx, 1) on the CUDA path against digests of the reference translation unit compiled with its commented-out
    four-tap code (cpp:575-620) switched on: identity and mapped paths, split draws, RGB and RGBA canvases, three flushes."""
    assert fn(gpu, image_rgba, switch=True) == golden_bilinear[name]


# ---- clip-rect extension: pinned to the unmodified reference (draw unclipped, put the outside pixels back) ---------------------
@pytest.mark.parametrize("name,fn", cases.clip_cases(), ids=lambda v: v if isinstance(v, str) else "")
def test_clip_rect_matches_the_reference_with_outside_pixels_put_back(name, fn, gpu, golden_clip, image_rgba):
    assert fn(gpu, image_rgba, native=True) == golden_clip[name]


# ---- N-gon fill extension: pinned to the reference's own DrawLine machinery (oracle/ref_polygon_shim.cpp) ---------------------
@pytest.mark.parametrize("name,fn", cases.polygon_cases(), ids=lambda v: v if isinstance(v, str) else "")
def test_polygon_fill_matches_the_references_drawline_machinery(name, fn, gpu, golden_polygon, image_rgba):
    assert fn(gpu, image_rgba) == golden_polygon[name]


# ---- perspective quads: spec map + the reference's own DrawTexture tail (oracle/ref_polygon_shim.cpp) ---------------------------
@pytest.mark.parametrize("name,fn", cases.perspective_cases(), ids=lambda v: v if isinstance(v, str) else "")
def test_perspective_quads_match_the_reference_tail_behind_the_spec_map(name, fn, gpu, golden_perspective, image_rgba):
    assert fn(gpu, image_rgba) == golden_perspective[name]


# ---- all extensions mixed (incl. perspective, which has no reference counterpart): product vs this repo's own C restatement ----
@pytest.mark.parametrize("seed", range(6))
def test_extensions_match_port_parity_unpinned(seed, gpu, port, image_rgba):
    """Clip rect, bilinear sampling, N-gon fill and perspective quads mixed in one stream (include/ncr_b200.h §2).  Each extension is pinned on
    its own above (the perspective map itself is this repo's spec: the reference has nothing projective); this test mixes them and
    adds RGB canvases, f64 and 3-channel textures, against the restatement.  RGB and RGBA canvases, u8 and f64 textures."""
    w, h, alpha = [(160, 90, True), (97, 61, False), (256, 144, True)][seed % 3]
    got = []
    for R in (gpu, port):
        ctx = R.RenderContext(w, h, alpha)
        tex = cases.tiny_textures(R, image_rgba)
        tex.append(R.Texture(7, 6, True, np.random.RandomState(5).rand(6, 7, 4).tobytes(), is_uint8=False))
        tex.append(R.Texture.from_numpy(np.random.RandomState(8).randint(0, 256, (12, 10, 3)).astype(np.uint8)))
        streams.stream_extensions(ctx, tex, seed, n=80)
        got.append(cases.digest(ctx))
    assert got[0] == got[1]


def test_clip_rect_limits_every_draw_kind(gpu):
    ctx = gpu.RenderContext(64, 48, True)
    ctx.set_color(0, 0, 0, 0)
    ctx.set_clip_rect(10, 8, 20, 16)
    ctx.fill_color(1, 1, 1, 1)
    ctx.draw_rect(0, 0, 64, 48, 1, 0, 0, 1)
    ctx.draw_line(0, 0, 64, 48, 9, 0, 1, 0, 1)
    ctx.clear_clip_rect()
    img = ctx.get_buffer_np().reshape(48, 64, 4)
    touched = img[..., 3] != 0
    assert touched[8:24, 10:30].all() and touched.sum() == 20 * 16


# ---- host runtime behaviour ------------------------------------------------------------------------------------------
def test_early_submits_do_not_change_the_image(image_rgba, port):
    """With tiny batch limits (NCR_MAX_PENDING_CMDS) every few draws are submitted on their own, asynchronously, while
    recording continues in the other staging buffer: the result must still be the restatement's, bit for bit."""
    import os
    import subprocess
    import sys
    import json

    from conftest import ROOT

    want = cases.make_random_case(511)(port, image_rgba)
    code = (
        "import sys, json; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import numpy as np, cases\n"
        "from libnativecpurenderer_b200.binding import Renderer\n"
        "img = np.load(%r)['rgba']\n"
        "print(json.dumps(cases.make_random_case(511)(Renderer(), img)))\n"
        % (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden", "image_rgba.npz"))
    )
    for limit in ("3", "17"):
        env = dict(os.environ, NCR_MAX_PENDING_CMDS=limit)
        res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
        assert res.returncode == 0, res.stderr
        assert json.loads(res.stdout.strip().splitlines()[-1]) == want


def test_contexts_on_concurrent_threads_are_independent(gpu, image_rgba):
    """One context per thread (ctypes releases the GIL): each thread's frames equal the single-threaded render."""
    import threading

    tex_np = streams.make_c2_textures()
    tex = [gpu.Texture.from_numpy(t) for t in tex_np]

    def render(seed):
        ctx = gpu.RenderContext(320, 180, True)
        out = []
        for f in range(3):
            streams.stream_c2(ctx, tex, n=150, seed=seed + f)
            out.append(hashlib.sha1(bytes(ctx.get_buffer_as_uint8())).hexdigest())
        return out

    want = {s: render(s) for s in (10, 20, 30, 40)}
    got = {}
    threads = [threading.Thread(target=lambda s=s: got.__setitem__(s, render(s))) for s in want]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert got == want


def test_stale_and_null_handles_are_ignored_not_dereferenced(gpu, image_rgba):
    lib = gpu.lib
    ctx = gpu.RenderContext(16, 16, True)
    tex = gpu.Texture.from_numpy(image_rgba)
    ptr, tptr = ctx._ptr, tex._ptr
    ctx.set_color(.5, .5, .5, .5)
    lib.DrawTexture(ptr, None, 0.0, 0.0, 4.0, 4.0)          # NULL texture: no-op
    lib.DestroyTexture(tptr)
    tex._ptr = 0
    lib.DrawTexture(ptr, tptr, 0.0, 0.0, 4.0, 4.0)          # destroyed texture: detected, no-op
    assert lib.GetTextureWidth(tptr) == 0
    assert set(ctx.get_buffer_as_uint8()) == {127}
    lib.DestroyRenderContext(ptr)
    ctx._ptr = 0
    assert lib.GetBufferSize(ptr) == 0                        # destroyed context: detected
    lib.DrawRect(ptr, 0.0, 0.0, 4.0, 4.0, 1.0, 1.0, 1.0, 1.0)
    lib.DestroyRenderContext(ptr)                             # double destroy: harmless
    assert lib.GetBufferSize(None) == 0


def _full_size_vs_cpu(gpu, cpu, workload):
    """The bench's exact workload (bench.build_workload) on both libraries through the C replayer; u8 + f64 digests."""
    import sys

    from conftest import REPLAY_LIB, ROOT

    sys.path.insert(0, ROOT)
    import bench

    w, h, alpha, tex_np, arr, draws, full = bench.build_workload(workload)
    got = []
    for R in (gpu, cpu):
        ctx = R.RenderContext(w, h, alpha)
        tex = [R.Texture.from_numpy(t) for t in tex_np]
        trace.Replayer(REPLAY_LIB, R.path).run(ctx, arr, tex)
        got.append(cases.digest(ctx))
    return got


def test_c3_full_size_matches_reference_4k(gpu, ref):
    """BASELINE config 3 (affine variant) at full size — 3840x2160 RGBA, 50,000 atlas sprites — against the UNMODIFIED
    reference build (about 20 s of one host core)."""
    got = _full_size_vs_cpu(gpu, ref, "c3")
    assert got[0] == got[1]


def test_c4_full_size_matches_reference_rgb_canvas(gpu, ref):
    """The milrenderer-shaped 1080p frame on a 3-channel canvas (mil:114 uses enable_alpha=False): identity-path
    background, FillColor dim, gradients, DrawLine bodies, notes, hit effects — against the unmodified reference."""
    got = _full_size_vs_cpu(gpu, ref, "c4")
    assert got[0] == got[1]


def test_fuzz_many_seeds_against_port(gpu, port, image_rgba):
    """Sixty more seeded streams (five canvas shapes incl. 3-channel and ragged sizes, three flushes each), u8 + f64."""
    bad = []
    for seed in range(1000, 1060):
        run = cases.make_random_case(seed, use_apply_pixel=(seed % 2 == 0))
        if run(gpu, image_rgba) != run(port, image_rgba):
            bad.append(seed)
    assert not bad, f"seeds that differ from the C restatement: {bad}"


def test_perspective_pixel_box_is_conservative(gpu, port):
    """The recorder bounds a perspective quad by the forward-mapped source rectangle; the restatement scans the whole canvas.
    Same pixels must come out (extension: parity unpinned against the reference, pinned between the two implementations)."""
    atlas_np = streams.make_atlas(cells=4, cell=32)
    got = []
    for R in (gpu, port):
        ctx = R.RenderContext(384, 216, True)
        atlas = R.Texture.from_numpy(atlas_np)
        streams.stream_c3p(ctx, atlas, n=400, cells=4)
        got.append(cases.digest(ctx))
    assert got[0] == got[1]


@pytest.mark.parametrize("shape", [(160, 90, True), (160, 90, False), (97, 61, False), (257, 129, True), (2, 2, False), (1, 1, True), (8, 3, True), (24, 5, False)])
def test_yuv420p_present_matches_port(shape, gpu, port, image_rgba):
    """Present path: NcrGetBufferAsYUV420P (flush + fused u8 image + ncr_yuv420p + 1.5 B/px readback) against
    the C restatement on the same random stream, incl. odd and tiny sizes (those follow this repo's clamped-neighbour
    definition; even sizes >= 8x8 are pinned to libswscale, see test_present_path_reproduces_libswscale)."""
    w, h, alpha = shape
    got = []
    for R in (gpu, port):
        ctx = R.RenderContext(w, h, alpha)
        tex = cases.tiny_textures(R, image_rgba)
        streams.stream_random(ctx, tex, 77, n=60)
        yuv = ctx.get_buffer_as_yuv420p()          # draws pending: composite (u8 image) + ncr_yuv420p in one flush
        assert yuv.size == w * h + 2 * ((w + 1) // 2) * ((h + 1) // 2)
        again = ctx.get_buffer_as_yuv420p()        # nothing pending, planes still current: just a copy
        ctx.fill_color(.3, .6, .9, .25)
        ctx.get_buffer_as_uint8()                  # flush without the planes ...
        alone = ctx.get_buffer_as_yuv420p()        # ... so these come from the standalone kernel
        got.append((yuv.tobytes(), again.tobytes(), alone.tobytes(), cases.digest(ctx)))
        assert got[-1][0] == got[-1][1] and got[-1][0] != got[-1][2]
    assert got[0] == got[1]


def test_yuv420p_full_size_and_video_cap_path(gpu):
    """1080p RGB chart frame (the milrenderer shape): planes are consistent with the u8 image of the same canvas, and
    PutRendererContextFrame (h:91) drives the same path."""
    w, h = 1920, 1080
    chart = streams.make_chart_textures()
    bg = np.ascontiguousarray(np.resize(streams.make_noise_texture(256, 7), (h, w, 4)))
    tex_np = [bg] + chart
    tex = [gpu.Texture.from_numpy(t) for t in tex_np]
    ctx = gpu.RenderContext(w, h, False)
    rec = trace.TraceRecorder(w, h, False)
    slots = [trace.TexSlot(k, t.shape[1], t.shape[0]) for k, t in enumerate(tex_np)]
    streams.stream_c4_frame(rec, slots[0], slots[1:], frame=7, n_notes=300)
    trace.submit_trace(ctx, rec.as_array(), tex)
    yuv = ctx.get_buffer_as_yuv420p()
    img = np.frombuffer(ctx.get_buffer_as_uint8(), dtype=np.uint8).reshape(h, w, 3)
    assert np.array_equal(yuv, cases.swscale_model(img))       # libswscale's arithmetic (pinned: tests/test_oracle.py)
    import sys

    from conftest import GOLDEN_DIR

    sys.path.insert(0, GOLDEN_DIR)
    import make_swscale_fixtures as mk

    libs = mk.load_swscale()
    if libs is not None:                                        # and the real library, where this machine has one
        assert yuv.tobytes() == mk.swscale_yuv420p(libs, img).tobytes()
    raw = ctypes.CDLL(gpu.path)   # the video entry points are not part of the render binding: plain ctypes, as pyb:425-478 does
    raw.CreateVideoCap.restype = ctypes.c_void_p
    raw.CreateVideoCap.argtypes = (ctypes.c_long, ctypes.c_long, ctypes.c_double)
    raw.PutRendererContextFrame.argtypes = (ctypes.c_void_p, ctypes.c_void_p)
    raw.DestroyVideoCap.argtypes = (ctypes.c_void_p,)
    cap = raw.CreateVideoCap(w, h, 60.0)
    assert cap
    d2h0 = ctx.stats().d2h_bytes
    raw.PutRendererContextFrame(cap, ctx._ptr)
    assert ctx.stats().d2h_bytes - d2h0 == w * h * 3 // 2   # 1.5 B/px leave the GPU, not 3
    raw.DestroyVideoCap(cap)


@pytest.mark.parametrize("present", ["u8", "yuv420p"])
def test_batch_render_delivers_every_frame_in_order(present, gpu, port):
    """NcrRenderFrames: 13 distinct chart frames on 3 worker contexts come back in frame order and each is
    byte-identical to the same frame rendered alone — on the product and on the C restatement."""
    from libnativecpurenderer_b200 import batch

    w, h = 320, 180
    chart = streams.make_chart_textures()
    bg = np.ascontiguousarray(np.resize(streams.make_noise_texture(64, 7), (h, w, 4)))
    tex_np = [bg] + chart
    slots = [trace.TexSlot(k, t.shape[1], t.shape[0]) for k, t in enumerate(tex_np)]
    traces = []
    for f in range(13):
        rec = trace.TraceRecorder(w, h, False)
        streams.stream_c4_frame(rec, slots[0], slots[1:], frame=17 * f, n_notes=40, n_fx=6)
        rec.save_state()          # state left behind by a frame must not leak into the next one on the same worker
        rec.translate(5, 5)
        rec.present()
        traces.append(rec.as_array())
    tex = [gpu.Texture.from_numpy(t) for t in tex_np]
    got = []
    assert batch.render_frames(gpu, w, h, False, traces, tex, on_frame=lambda i, px: got.append((i, px.tobytes())),
                               workers=3, present=present) == 13
    assert [i for i, _ in got] == list(range(13))
    with batch.FramePool(gpu, w, h, False, workers=4) as pool:   # a kept pool gives the same frames, call after call
        for _ in range(2):
            again = []
            assert pool.render(traces, tex, on_frame=lambda i, px: again.append((i, px.tobytes())), present=present) == 13
            assert again == got
    ptex = [port.Texture.from_numpy(t) for t in tex_np]
    for R, T in ((gpu, tex), (port, ptex)):
        for f in (0, 5, 12):
            ctx = R.RenderContext(w, h, False)
            trace.submit_trace(ctx, traces[f], T) if R is gpu else _replay_calls(ctx, traces[f], T, R)
            want = ctx.get_buffer_as_yuv420p().tobytes() if present == "yuv420p" else bytes(ctx.get_buffer_as_uint8())
            assert got[f][1] == want
    assert len({px for _, px in got}) == 13   # the frames really differ


def _replay_calls(ctx, arr, textures, R):
    from conftest import REPLAY_LIB

    trace.Replayer(REPLAY_LIB, R.path).run(ctx, arr, textures)


def test_8k_canvas_matches_port(gpu, port, image_rgba):
    """Maximum-size end of the path: a 7680x4320 RGB canvas (796 MB of f64 on the device, 129,600 tiles) with one draw of
    every kind, bit-exact against the C restatement."""
    w, h = 7680, 4320
    got = []
    for R in (gpu, port):
        ctx = R.RenderContext(w, h, False)
        tex = R.Texture.from_numpy(image_rgba)
        ctx.set_color(.2, .3, .4, 1)
        ctx.save_state()
        ctx.translate(w * .5, h * .5)
        ctx.rotate(.3)
        ctx.scale(20, 12)
        ctx.apply_color_transform(1, .9, .8, .7)
        ctx.draw_texture(tex, -64, -64, 128, 128)
        ctx.draw_splitted_texture(tex, 70, -64, 128, 128, .1, .9, .2, .8)
        ctx.restore_state()
        ctx.draw_rect(7000.5, 4000.5, 900, 900, 1, 0, 0, .5)          # clipped by the right/bottom edges
        ctx.draw_vertical_grd(0, 0, w, 400, 0, 0, 0, .8, 0, 0, 0, 0.0)
        ctx.draw_circle(1000, 3000, 700, 0, 1, 0, .4)
        ctx.draw_line(100, 4200, 7600, 100, 9, 1, 1, 1, .9)
        img = ctx.get_buffer_as_uint8()
        got.append(hashlib.sha1(bytes(img)).hexdigest())
        del ctx
    assert got[0] == got[1]


def test_100k_tiny_draws_one_flush(gpu, port):
    """Command-count end of the path: 100,000 small rects and circles in ONE flush (tile-list capacity is sized from the
    recorded boxes), bit-exact against the C restatement, u8 and f64."""
    import random

    got = []
    for R in (gpu, port):
        ctx = R.RenderContext(640, 360, True)
        ctx.set_color(0, 0, 0, 1)
        rng = random.Random(11)
        for k in range(100000):
            x, y = rng.uniform(-5, 645), rng.uniform(-5, 365)
            if k % 7:
                ctx.draw_rect(x, y, rng.uniform(.5, 6), rng.uniform(.5, 6), rng.random(), rng.random(), rng.random(), rng.uniform(.1, 1))
            else:
                ctx.draw_circle(x, y, rng.uniform(.5, 4), rng.random(), rng.random(), rng.random(), rng.uniform(.1, 1))
        got.append(cases.digest(ctx))
    assert got[0] == got[1]


def test_draw_texture_batch_equals_the_call_loop(gpu, port, image_rgba):
    """NcrDrawTextureBatch: n sprites in one FFI crossing == the save / apply_transform / apply_color_transform / draw /
    restore loop, bit for bit — against the product's own loop and against the C restatement's loop."""
    import math
    import random

    rng = random.Random(5)
    n = 400
    m, ct, xywh, uv = [], [], [], []
    for _ in range(n):
        a, s = rng.uniform(0, 6.28), rng.uniform(.2, 1.5)
        m.append([s * math.cos(a), s * math.sin(a), -s * math.sin(a), s * math.cos(a), rng.uniform(0, 480), rng.uniform(0, 270)])
        ct.append([1, rng.uniform(.5, 1), rng.uniform(.5, 1), 1.0 if rng.random() < .2 else rng.uniform(.1, 1)])
        xywh.append([-40, -30, 80, 60])
        u0, v0 = rng.uniform(0, .5), rng.uniform(0, .5)
        uv.append([u0, u0 + rng.uniform(.2, .5), v0, v0 + rng.uniform(.2, .5)])
    out = {}
    for name, R in (("gpu_batch", gpu), ("gpu_loop", gpu), ("port_loop", port)):
        ctx = R.RenderContext(480, 270, True)
        tex = R.Texture.from_numpy(image_rgba)
        ctx.set_color(.1, .2, .3, 1)
        ctx.translate(3, 4)          # an outer transform the batch composes with
        for split in (False, True):
            if name == "gpu_batch":
                assert ctx.draw_texture_batch(tex, xywh, m, ct, uv if split else None) == n
            else:
                for k in range(n):
                    ctx.save_state()
                    ctx.apply_transform(*m[k])
                    ctx.apply_color_transform(*ct[k])
                    if split:
                        ctx.draw_splitted_texture(tex, *xywh[k], *uv[k])
                    else:
                        ctx.draw_texture(tex, *xywh[k])
                    ctx.restore_state()
        out[name] = cases.digest(ctx)
    assert out["gpu_batch"] == out["gpu_loop"] == out["port_loop"]


def test_first_use_from_many_threads_at_once():
    """Device selection happens on first use; a fresh process whose first library calls come from 8 threads at the same
    time (what a frame pool or a threaded host does) must initialise once and give every thread a working context."""
    import subprocess
    import sys

    code = r"""
import sys, threading
sys.path.insert(0, %r)
from libnativecpurenderer_b200.binding import Renderer
R = Renderer()
out, go = [None] * 8, threading.Barrier(8)
def work(k):
    go.wait()
    ctx = R.RenderContext(64, 32, True)
    ctx.set_color(k / 8.0, 0, 0, 1)
    out[k] = bytes(ctx.get_buffer_as_uint8())[:4]
ts = [threading.Thread(target=work, args=(k,)) for k in range(8)]
[t.start() for t in ts]; [t.join() for t in ts]
assert out == [bytes([int(k / 8.0 * 255), 0, 0, 255]) for k in range(8)], out
print("ok")
""" % (str(__import__("pathlib").Path(__file__).resolve().parents[1]),)
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and res.stdout.strip().endswith("ok"), res.stdout + res.stderr


def test_frame_pool_reports_bad_frames(gpu, image_rgba):
    """A malformed trace in the middle of a batch fails the whole render (-1 -> RuntimeError) without hanging the pool, and
    the pool is usable afterwards."""
    from libnativecpurenderer_b200 import batch

    tex = [gpu.Texture.from_numpy(image_rgba)]
    def frame(k):
        rec = trace.TraceRecorder(96, 64, True)
        rec.set_color(0, 0, k / 10.0, 1)
        rec.draw_texture(trace.TexSlot(0, 128, 128), 5 + k, 5, 40, 40)
        rec.present()
        return rec.as_array()
    good = [frame(k) for k in range(9)]
    bad = list(good)
    broken = good[4].copy()
    broken.view(np.uint32)[11] = 0x7fffffff         # argument count of the 2nd record (DrawTexture): runs past the end of the trace
    bad[4] = broken
    with batch.FramePool(gpu, 96, 64, True, workers=3) as pool:
        with pytest.raises((RuntimeError, ValueError)):
            pool.render(bad, tex)
        seen = []
        assert pool.render(good, tex, on_frame=lambda i, px: seen.append(i)) == 9 and seen == list(range(9))


def test_present_only_flush_leaves_a_recoverable_canvas(gpu, port, image_rgba):
    """GetBufferAsUInt8 / YUV flushes do not write the f64 canvas back (present-only flush); every later
    consumer of the canvas — more draws without SetColor, GetBuffer, GetColor, canvas -> texture, a shared alias — must
    still see exactly what the reference's immediate-mode canvas holds."""
    outs = []
    for R in (gpu, port):
        out = []
        tex = cases.tiny_textures(R, image_rgba)
        ctx = R.RenderContext(150, 90, True)
        ctx.set_color(.1, .2, .3, 1)
        streams.stream_random(ctx, tex, 501, n=40)
        out.append(cases.sha(ctx.get_buffer_as_uint8()))           # present-only flush
        streams.stream_random(ctx, tex, 502, n=40)                 # draws on top of the (stale) canvas, no SetColor
        out.append(cases.sha(ctx.get_buffer_as_uint8()))
        out.append(cases.sha(ctx.get_buffer_np().tobytes()))        # f64 canvas after two presents
        ctx.draw_circle(40, 40, 25, 1, 0, 1, .5)
        out.append(cases.sha(ctx.get_buffer_as_uint8()))
        out.append(ctx.get_color(41.5, 39.0))                       # single-pixel read of a presented canvas
        ctx.fill_color(0, 1, 0, .25)
        out.append(cases.sha(ctx.get_buffer_as_uint8()))
        snap = ctx.as_texture()                                     # CreateTextureFromRenderContext after a present
        shared = ctx.as_texture_shared()
        dst = R.RenderContext(64, 64, True)
        dst.set_color(0, 0, 0, 1)
        dst.translate(2, 3)
        dst.rotate(.3)
        dst.draw_texture(snap, 0, 0, 50, 40)
        dst.draw_texture(shared, 10, 10, 40, 50)
        out.append(cases.digest(dst))
        ctx.set_color(.5, .5, .5, .5)                               # a stale canvas that is overwritten, never read
        ctx.draw_rect(5, 5, 60, 30, 1, 0, 0, .5)
        out.append(cases.sha(ctx.get_buffer_as_uint8()))
        out.append(cases.sha(ctx.get_buffer_as_uint8()))            # nothing pending: cached image
        out.append(cases.sha(ctx.get_buffer_np().tobytes()))
        outs.append(out)
    assert outs[0] == outs[1]


def test_video_frames_never_write_the_canvas_back(gpu, image_rgba):
    """A milrenderer-shaped loop (SetColor ... GetBufferAsUInt8 per frame, mil:866-1036) runs one composite per frame and
    never has to bring a canvas up to date."""
    ctx = gpu.RenderContext(320, 180, False)
    tex = cases.tiny_textures(gpu, image_rgba)
    before = ctx.stats()
    for f in range(6):
        ctx.set_color(0, 0, 0, 0)
        streams.stream_random(ctx, tex, 600 + f, n=30)
        ctx.get_buffer_as_uint8()
    st = ctx.stats()
    assert st.materialized == before.materialized == 0
    assert st.flushes - before.flushes == 6
    assert st.kernel_launches - before.kernel_launches == 12   # bin_fine + composite per frame (31 commands: no coarse pass)


def test_u8_truncation_edge_values_on_the_gpu(gpu, port):
    """(iu8)(v*255) of out-of-range canvas values (cpp:52-57: cvttsd2si semantics — truncation toward zero, low byte of a
    32-bit conversion, NaN / overflow -> 0x80000000 -> 0): through the composite's fused u8 image, through the standalone
    convert kernel, through the fused and standalone YUV paths, and blended on top of (NaN and inf propagate through
    dst*(1-a)+src*a exactly as in IEEE f64).  Expected bytes are the reference's own (tests/test_oracle.py checks the same
    list against the unmodified reference build)."""
    vals = [2.0, -0.01, 1e10, 300.7 / 255, float("nan"), -1e10, 1.0, 0.999999, 0.625, 1 / 3]
    outs = []
    for R in (gpu, port):
        out = []
        for alpha in (True, False):
            ctx = R.RenderContext(20, 10, alpha)
            ctx.set_color(0, 0, 0, 0)
            ctx.set_pixel(0, 0, *vals[0:4])
            ctx.set_pixel(1, 0, *vals[4:8])
            ctx.set_pixel(2, 0, vals[8], vals[9], 0, 0)
            ctx.set_pixel(17, 9, float("inf"), -float("inf"), -0.0, 1e300)
            fused = bytes(ctx.get_buffer_as_uint8())
            out.append(list(fused[:12]))
            out.append(cases.sha(fused))
            out.append(cases.sha(ctx.get_buffer_as_yuv420p().tobytes()))
            if R is gpu: ctx.flush()                        # (the CPU checker is immediate-mode: nothing to flush)
            ctx.draw_rect(19, 9, 1, 1, 0, 0, 0, 0)        # invalidates the cached images without changing those pixels
            if R is gpu: ctx.flush()
            out.append(cases.sha(ctx.get_buffer_as_uint8()))                       # standalone convert kernel
            out.append(cases.sha(ctx.get_buffer_as_yuv420p().tobytes()))          # standalone YUV kernel
            ctx.fill_color(.25, .5, .75, .5)                                       # blend over NaN / inf / huge values
            ctx.draw_rect(0, 0, 3, 1, 1, 1, 1, 1)                                  # a == 1: plain store over them
            ctx.draw_rect(16, 8, 4, 2, .5, .5, .5, .999)
            out.append(cases.digest(ctx))
        outs.append(out)
    assert outs[0][0] == [254, 254, 0, 44, 0, 0, 255, 254, 159, 85, 0, 0]
    assert outs[0] == outs[1]


def test_contexts_on_explicit_devices_and_multi_device_frame_pool(gpu, image_rgba):
    """One host process, several GPUs (what milrenderer.py needs on an 8-GPU box): contexts are placed on explicit devices,
    a texture created once is copied to a device the first time a context there draws it, and a frame pool spread over
    the devices delivers the same frames, in order, as a single-device pool.  With one visible device the pool is spread
    over [0, 0]; with two or more (gpurun --gpus 2) the second device really renders."""
    from libnativecpurenderer_b200 import batch

    n_dev = gpu.lib.NcrDeviceCount()
    assert n_dev >= 1
    devices = [0, 1 % n_dev]
    tex = gpu.Texture.from_numpy(image_rgba)           # lives on the default device
    want = None
    for d in sorted(set(devices)):
        p = gpu.lib.NcrCreateRenderContextOnDevice(96, 64, True, d)
        assert p and gpu.lib.NcrContextDevice(p) == d
        ctx = gpu.context_from_ptr(p, 96, 64, True)
        streams.stream_k1(ctx, tex, n=40)
        got = cases.digest(ctx)
        want = want or got
        assert got == want
    assert not gpu.lib.NcrCreateRenderContextOnDevice(8, 8, True, n_dev)   # out of range: NULL + error text
    assert "device" in gpu.last_error()

    w, h = 320, 180
    chart = streams.make_chart_textures()
    bg = np.ascontiguousarray(np.resize(streams.make_noise_texture(64, 7), (h, w, 4)))
    tex_np = [bg] + chart
    slots = [trace.TexSlot(k, t.shape[1], t.shape[0]) for k, t in enumerate(tex_np)]
    traces = []
    for f in range(11):
        rec = trace.TraceRecorder(w, h, False)
        streams.stream_c4_frame(rec, slots[0], slots[1:], frame=13 * f, n_notes=40, n_fx=6)
        rec.present()
        traces.append(rec.as_array())
    texs = [gpu.Texture.from_numpy(t) for t in tex_np]
    one, many = [], []
    with batch.FramePool(gpu, w, h, False, workers=2) as pool:
        assert pool.render(traces, texs, on_frame=lambda i, px: one.append((i, px.tobytes()))) == 11
    with batch.FramePool(gpu, w, h, False, workers=4, devices=devices) as pool:
        for present in ("u8", "yuv420p"):
            many.clear()
            assert pool.render(traces, texs, on_frame=lambda i, px: many.append((i, px.tobytes())), present=present) == 11
            assert [i for i, _ in many] == list(range(11))
            if present == "u8":
                assert many == one


# ---- every BASELINE configuration at its full size against digests of the UNMODIFIED reference build -----------------------
def _bench_module():
    import sys

    from conftest import ROOT

    sys.path.insert(0, ROOT)
    import bench

    return bench


@pytest.mark.parametrize("workload", ["c1", "c2", "c3", "c4", "c5", "bg"])
def test_full_size_workload_matches_the_reference_digest(workload, gpu):
    """bench.py's workloads (BASELINE configs 1-5 + the low-overdraw frame) at their full sizes — incl. C2's 20,000 draws and
    the 4K RGB chart frame of config 5 — rendered through the C ABI; the RGB(A)8 frame must hash to the digest of the
    unmodified reference build (tests/golden/bench_golden.json, generated by tests/golden/make_bench_golden.py)."""
    from conftest import REPLAY_LIB

    bench = _bench_module()
    want = bench.load_golden()["workloads"][workload]["u8"]
    w, h, alpha, tex_np, arr, draws, full = bench.build_workload(workload)
    ctx = gpu.RenderContext(w, h, alpha)
    tex = [gpu.Texture.from_numpy(t) for t in tex_np]
    trace.Replayer(REPLAY_LIB, gpu.path).run(ctx, arr, tex)
    assert cases.sha(ctx.get_buffer_as_uint8()) == want
    f64 = ctx.get_buffer_np()                                   # the stale canvas, brought up to date on demand
    scaled = f64 * 255
    safe = np.where(np.abs(scaled) < 2 ** 31, scaled, 0)
    assert cases.sha((np.trunc(safe).astype(np.int64) & 0xFF).astype(np.uint8).tobytes()) == want


@pytest.mark.parametrize("name,frames", [("c4", [0, 123, 359, 719, 3600, 7199]), ("c5", [0, 123, 239, 479, 959, 1919])])
def test_chart_video_frames_match_the_reference_digests(name, frames, gpu):
    """Six frame indices of the 7,200-frame 1080p chart (config 4) and of the 4K chart (config 5), each rendered through the
    frame pool exactly as the video legs of bench.py render them (u8 and YUV present), against the reference's digests."""
    from libnativecpurenderer_b200 import batch

    bench = _bench_module()
    want = bench.load_golden()["video"][name]
    w, h, alpha, _ = bench.WORKLOADS[name]
    tex = [gpu.Texture.from_numpy(t) for t in bench.chart_textures(w, h)]
    traces = [bench.build_video_frame(name, f)[4] for f in frames]
    with batch.FramePool(gpu, w, h, alpha, workers=3) as pool:
        for present, key in (("u8", "u8"), ("yuv420p", "yuv420p")):
            got = {}
            assert pool.render(traces, tex, on_frame=lambda i, px: got.__setitem__(i, cases.sha(px.tobytes())), present=present) == len(frames)
            assert [got[i] for i in range(len(frames))] == [want[str(f)][key] for f in frames]


def test_present_path_reproduces_libswscale(gpu):
    """SURVEY 8-f1, pinned: the YUV 4:2:0 planes the product hands to the encoder are byte-identical to what a real libswscale
    (9.1.100; committed fixtures, tests/golden/make_swscale_fixtures.py) produces for PutRendererContextFrame's call — through
    the flush that has draws pending (composite + ncr_yuv420p) and through the one that has none (standalone conversion)."""
    fx = cases.swscale_fixtures()
    k = 0
    while f"img_{k}" in fx:
        img = fx[f"img_{k}"]
        ctx = cases.canvas_holding_u8_image(gpu, img)
        pending = ctx.get_buffer_as_yuv420p().tobytes()            # draws pending: present-only flush
        assert pending == fx[f"yuv_{k}"].tobytes(), f"fixture {k} {img.shape}"
        assert bytes(ctx.get_buffer_as_uint8()) == img.tobytes()
        ctx.flush()
        ctx.draw_rect(0, 0, 1, 1, 0, 0, 0, 0)                       # invalidates the cached planes without changing a pixel
        ctx.flush()
        assert ctx.get_buffer_as_yuv420p().tobytes() == pending    # nothing pending: converted from the canvas
        k += 1
    assert k >= 5


def test_non_default_device_leaves_device_0_untouched(gpu):
    """with NCR_DEVICE=1 the recording path of a fresh worker thread (frame pool, replayer threads) used to allocate
    pinned staging with device 0 current, creating a primary context on GPU 0 for every rank.  A process that renders on device 1
    — pool workers, replayer threads and a per-call context — must not appear among GPU 0's compute processes (NVML).
    Needs two visible devices (gpurun --gpus 2)."""
    import subprocess
    import sys

    from conftest import REPLAY_LIB, ROOT

    if gpu.lib.NcrDeviceCount() < 2:
        pytest.skip("one visible device")
    code = f"""
import os, sys
os.environ["NCR_DEVICE"] = "1"
sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {ROOT + '/tests'!r})
import numpy as np, pynvml
from libnativecpurenderer_b200 import batch, streams, trace
from libnativecpurenderer_b200.binding import Renderer
pynvml.nvmlInit()
used0_before = pynvml.nvmlDeviceGetMemoryInfo(pynvml.nvmlDeviceGetHandleByIndex(0)).used
R = Renderer()
w, h = 320, 180
chart = streams.make_chart_textures()
bg = np.ascontiguousarray(np.resize(streams.make_noise_texture(64, 7), (h, w, 4)))
tex_np = [bg] + chart
slots = [trace.TexSlot(k, t.shape[1], t.shape[0]) for k, t in enumerate(tex_np)]
traces = []
for f in range(9):
    rec = trace.TraceRecorder(w, h, False); streams.stream_c4_frame(rec, slots[0], slots[1:], frame=f, n_notes=60, n_fx=6); rec.present(); traces.append(rec.as_array())
tex = [R.Texture.from_numpy(t) for t in tex_np]
assert batch.render_frames(R, w, h, False, traces, tex, workers=4) == 9
trace.Replayer({REPLAY_LIB!r}, R.path).run_threads(4, w, h, False, traces[0], tex, repeats=2)
ctx = R.RenderContext(64, 64, True); assert R.lib.NcrContextDevice(ctx._ptr) == 1
streams.stream_k1(ctx, R.Texture.from_numpy(np.zeros((8, 8, 4), np.uint8)), n=20); ctx.get_buffer_as_uint8()
pids0 = [p.pid for p in pynvml.nvmlDeviceGetComputeRunningProcesses(pynvml.nvmlDeviceGetHandleByIndex(0))]
pids1 = [p.pid for p in pynvml.nvmlDeviceGetComputeRunningProcesses(pynvml.nvmlDeviceGetHandleByIndex(1))]
grown0 = pynvml.nvmlDeviceGetMemoryInfo(pynvml.nvmlDeviceGetHandleByIndex(0)).used - used0_before
print("ON0", os.getpid() in pids0, "ON1", os.getpid() in pids1, "GROWN0_MB", grown0 >> 20)
"""
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("ON0")][-1].split()
    assert line[1] == "False", res.stdout                 # this process is not a compute process of GPU 0 ...
    assert int(line[5]) < 64, res.stdout                  # ... and GPU 0's memory did not grow by a CUDA context (hundreds of MB)
    # (line[3] is True where NVML reports container-local pids; in a pid namespace the memory check alone carries the test)


def test_present_path_scaling_branch_reproduces_libswscale(gpu, port):
    """cap size != canvas size (cpp:241-256: sws_scale resizes while converting): NcrGetBufferAsYUV420PScaled and
    PutRendererContextFrame with a VideoCap of another size give the planes of the real libswscale — committed fixtures (shrink,
    enlarge, odd destination sizes, one axis only), the live library at video sizes where it loads, and the C restatement on
    odd / tiny shapes outside the pinned domain."""
    import sys

    from conftest import GOLDEN_DIR

    fx = cases.swscale_fixtures()
    ctxs = {}
    for n, (k, dw, dh) in enumerate(fx["scaled"]):
        k, dw, dh = int(k), int(dw), int(dh)
        if k not in ctxs:
            ctxs[k] = cases.canvas_holding_u8_image(gpu, fx[f"img_{k}"])
        assert ctxs[k].get_buffer_as_yuv420p_scaled(dw, dh).tobytes() == fx[f"syuv_{n}"].tobytes(), (n, k, dw, dh)
    # PutRendererContextFrame with a cap of another size drives the same path (h:91)
    raw = ctypes.CDLL(gpu.path)
    raw.CreateVideoCap.restype = ctypes.c_void_p
    raw.CreateVideoCap.argtypes = (ctypes.c_long, ctypes.c_long, ctypes.c_double)
    raw.PutRendererContextFrame.argtypes = (ctypes.c_void_p, ctypes.c_void_p)
    cap = raw.CreateVideoCap(48, 32, 30.0)
    d2h0 = ctxs[0].stats().d2h_bytes
    raw.PutRendererContextFrame(cap, ctxs[0]._ptr)
    assert ctxs[0].stats().d2h_bytes - d2h0 == 48 * 32 * 3 // 2
    # live library, video sizes: 720p -> 1080p and 1080p -> 720p
    sys.path.insert(0, GOLDEN_DIR)
    import make_swscale_fixtures as mk

    libs = mk.load_swscale()
    rs = np.random.RandomState(5)
    if libs is not None:
        for (h, w, c), (dw, dh) in (((720, 1280, 3), (1920, 1080)), ((1080, 1920, 4), (1280, 720)), ((270, 480, 3), (854, 481))):
            img = rs.randint(0, 256, (h, w, c)).astype(np.uint8)
            ctx = gpu.RenderContext(w, h, c == 4)
            ctx.set_color(0, 0, 0, 0)
            tex = np.zeros((h + 1, w + 1, 4), dtype=np.uint8)
            tex[:h, :w, :c] = img
            tex[..., 3] = 255
            ctx.draw_texture(gpu.Texture.from_numpy(tex), 0, 0, w + 1, h + 1)     # RGBA8 texels k/255 with a == 1: stored as they are
            got_img = np.frombuffer(ctx.get_buffer_as_uint8(), dtype=np.uint8).reshape(h, w, c)
            assert ctx.get_buffer_as_yuv420p_scaled(dw, dh).tobytes() == mk.swscale_yuv420p(libs, got_img, dw, dh).tobytes(), (w, h, dw, dh)
    # odd / tiny shapes (this repo's definition outside the pinned domain): product == restatement
    for (w, h, alpha), (dw, dh) in (((97, 61, False), (50, 33)), ((33, 20, True), (70, 41)), ((8, 3, True), (8, 3)), ((24, 5, False), (12, 9))):
        outs = []
        for R in (gpu, port):
            ctx = R.RenderContext(w, h, alpha)
            streams.stream_random(ctx, cases.tiny_textures(R, np.zeros((4, 4, 4), np.uint8)), 91, n=40)
            outs.append(ctx.get_buffer_as_yuv420p_scaled(dw, dh).tobytes())
        assert outs[0] == outs[1], (w, h, dw, dh)
