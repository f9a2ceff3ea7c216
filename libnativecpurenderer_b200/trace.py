"""Command-stream ("trace") recorder.

``TraceRecorder`` has the drawing interface of ``binding.RenderContext`` (the reference's
pyb:51-300 method names) but appends binary records instead of calling a library.  The result is
replayed without per-call FFI cost by

* ``NcrSubmitTrace`` (product, include/ncr_b200.h §2), or
* ``libnativecpurenderer_b200/csrc/ncr_replay.cpp`` (any library exporting the reference C ABI; measurement/tests only).

Record layout: ``uint32 op, uint32 n, float64 args[n]`` (csrc/ncr_trace.h).  Texture arguments are
slot numbers into the texture table given at replay time; use ``TexSlot`` in place of a ``Texture``.
"""
from __future__ import annotations

import ctypes
import math
import struct
from dataclasses import dataclass

import numpy as np

# csrc/ncr_trace.h
T_SAVE, T_RESTORE, T_SET_TRANSFORM, T_APPLY_TRANSFORM, T_SCALE, T_TRANSLATE, T_ROTATE = 1, 2, 3, 4, 5, 6, 7
T_SET_CT, T_APPLY_CT, T_SET_COLOR, T_FILL_COLOR, T_DRAW_TEXTURE, T_DRAW_SPLIT, T_DRAW_RECT = 8, 9, 10, 11, 12, 13, 14
T_DRAW_LINE, T_DRAW_CIRCLE, T_DRAW_GRD, T_SET_PIXEL, T_APPLY_PIXEL, T_PRESENT = 15, 16, 17, 18, 19, 20
T_CLIP_SET, T_CLIP_CLEAR, T_SAMPLING, T_FILL_POLY, T_DRAW_PERSP = 32, 33, 34, 35, 36


@dataclass(frozen=True)
class TexSlot:
    """Stand-in for a texture while recording: slot index plus the size generators may need."""

    slot: int
    width: int
    height: int

    @property
    def _ptr(self):  # so accidental use with a live context fails loudly
        raise TypeError("TexSlot is only valid with TraceRecorder")


class TraceRecorder:
    def __init__(self, width: int, height: int, enable_alpha: bool):
        self.width = width
        self.height = height
        self.enable_alpha = enable_alpha
        self._chunks: list[bytes] = []
        self.n_records = 0
        self.n_draws = 0

    # -- encoding --
    def _rec(self, op: int, *args: float) -> None:
        self._chunks.append(struct.pack(f"<II{len(args)}d", op, len(args), *args))
        self.n_records += 1

    def tobytes(self) -> bytes:
        return b"".join(self._chunks)

    def as_array(self) -> np.ndarray:
        """8-byte aligned copy of the stream (records are multiples of 8 bytes)."""
        raw = self.tobytes()
        arr = np.frombuffer(raw, dtype=np.uint8).copy()
        out = np.empty(len(raw) // 8 + 1, dtype=np.float64)  # float64 storage guarantees alignment
        out.view(np.uint8)[: len(raw)] = arr
        return out.view(np.uint8)[: len(raw)]

    # -- state --
    def save_state(self): self._rec(T_SAVE)
    def restore_state(self): self._rec(T_RESTORE)
    def set_transform(self, a, b, c, d, e, f): self._rec(T_SET_TRANSFORM, a, b, c, d, e, f)
    def apply_transform(self, a, b, c, d, e, f): self._rec(T_APPLY_TRANSFORM, a, b, c, d, e, f)
    def scale(self, sx, sy): self._rec(T_SCALE, sx, sy)
    def translate(self, tx, ty): self._rec(T_TRANSLATE, tx, ty)
    def rotate(self, angle): self._rec(T_ROTATE, angle)
    def rotate_degree(self, deg): self.rotate(deg * math.pi / 180)
    def set_color_transform(self, r, g, b, a): self._rec(T_SET_CT, r, g, b, a)
    def apply_color_transform(self, r, g, b, a): self._rec(T_APPLY_CT, r, g, b, a)

    # -- pixels --
    def set_color(self, r, g, b, a): self._rec(T_SET_COLOR, r, g, b, a)
    def fill_color(self, r, g, b, a): self._draw(T_FILL_COLOR, r, g, b, a)
    def set_pixel(self, x, y, r, g, b, a): self._draw(T_SET_PIXEL, x, y, r, g, b, a)
    def apply_pixel(self, x, y, r, g, b, a): self._draw(T_APPLY_PIXEL, x, y, r, g, b, a)

    def _draw(self, op, *args):
        self._rec(op, *args)
        self.n_draws += 1

    # -- primitives --
    def draw_texture(self, tex: TexSlot, x, y, w, h): self._draw(T_DRAW_TEXTURE, tex.slot, x, y, w, h)

    def draw_splitted_texture(self, tex: TexSlot, x, y, width, height, u_start, u_end, v_start, v_end):
        self._draw(T_DRAW_SPLIT, tex.slot, x, y, width, height, u_start, u_end, v_start, v_end)

    def draw_rect(self, x, y, width, height, r, g, b, a): self._draw(T_DRAW_RECT, x, y, width, height, r, g, b, a)
    def draw_line(self, x0, y0, x1, y1, width, r, g, b, a): self._draw(T_DRAW_LINE, x0, y0, x1, y1, width, r, g, b, a)
    def draw_circle(self, x, y, radius, r, g, b, a): self._draw(T_DRAW_CIRCLE, x, y, radius, r, g, b, a)

    def draw_vertical_grd(self, x, y, width, height, *stops):
        self._draw(T_DRAW_GRD, x, y, width, height, *stops)

    def draw_vertical_mut_grd(self, x, y, width, height, steps):
        for (p0, c0), (p1, c1) in zip(steps, steps[1:]):
            self.draw_vertical_grd(x, y + height * p0, width, height * (p1 - p0), *c0[:4], *c1[:4])

    def present(self):
        """End of frame: the replayer reads the canvas back as u8 here."""
        self._rec(T_PRESENT)

    # -- extensions (product only) --
    def set_clip_rect(self, x, y, w, h): self._rec(T_CLIP_SET, x, y, w, h)
    def clear_clip_rect(self): self._rec(T_CLIP_CLEAR)
    def set_sampling(self, mode): self._rec(T_SAMPLING, mode)

    def fill_polygon(self, points, r, g, b, a):
        self._draw(T_FILL_POLY, r, g, b, a, *[float(v) for p in points for v in p])

    def draw_texture_perspective(self, tex: TexSlot, inv_h, x, y, w, h):
        self._draw(T_DRAW_PERSP, tex.slot, *[float(v) for v in inv_h], x, y, w, h)


def texture_table(textures) -> ctypes.Array:
    """``void*[]`` of live ``Texture`` handles, in slot order, for the replayers."""
    return (ctypes.c_void_p * len(textures))(*[t._ptr for t in textures])


def submit_trace(ctx, trace: np.ndarray, textures) -> int:
    """Product fast path: replay ``trace`` (from ``TraceRecorder.as_array``) on a live context."""
    table = texture_table(textures)
    n = ctx._lib.NcrSubmitTrace(ctx._ptr, ctypes.c_void_p(trace.ctypes.data), trace.nbytes, table, len(textures))
    if n < 0:
        raise ValueError("malformed trace")
    return n


class Replayer:
    """ctypes face of lib/libncr_replay.so bound to one target library (tests / bench only)."""

    def __init__(self, replay_lib_path: str, target_lib_path: str):
        self.lib = ctypes.CDLL(replay_lib_path)
        self.lib.ncr_replay_open.restype = ctypes.c_void_p
        self.lib.ncr_replay_open.argtypes = (ctypes.c_char_p,)
        self.lib.ncr_replay_run.restype = ctypes.c_double
        self.lib.ncr_replay_run.argtypes = (ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p,
                                            ctypes.c_long, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p)
        self.lib.ncr_replay_run_threads.restype = ctypes.c_double
        self.lib.ncr_replay_run_threads.argtypes = (ctypes.c_void_p, ctypes.c_int, ctypes.c_long, ctypes.c_long, ctypes.c_int,
                                                    ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_int)
        self.lib.ncr_replay_run_threads_ex.restype = ctypes.c_double
        self.lib.ncr_replay_run_threads_ex.argtypes = self.lib.ncr_replay_run_threads.argtypes + (ctypes.c_void_p, ctypes.c_long)
        self.lib.ncr_replay_set_present.restype = ctypes.c_int
        self.lib.ncr_replay_set_present.argtypes = (ctypes.c_void_p, ctypes.c_int)
        self.api = self.lib.ncr_replay_open(target_lib_path.encode())
        if not self.api:
            raise OSError(f"cannot bind {target_lib_path}")

    def set_present(self, mode: str) -> None:
        """What a PRESENT record reads back: "u8" (GetBufferAsUInt8, the reference ABI) or "yuv420p" (NcrGetBufferAsYUV420P)."""
        if self.lib.ncr_replay_set_present(self.api, {"u8": 0, "yuv420p": 1}[mode]) != 0:
            raise OSError("target library has no NcrGetBufferAsYUV420P")

    def run(self, ctx, trace: np.ndarray, textures, frame_address: int | None = None, repeats: int = 1) -> float:
        table = texture_table(textures)
        secs = self.lib.ncr_replay_run(self.api, ctx._ptr, ctypes.c_void_p(trace.ctypes.data), trace.nbytes, table,
                                       len(textures), ctypes.c_void_p(frame_address) if frame_address else None, repeats, None)
        if secs < 0:
            raise ValueError("malformed or unsupported trace")
        return secs

    def run_threads(self, n_threads: int, width: int, height: int, alpha: bool, trace: np.ndarray, textures,
                    repeats: int = 1, warm_repeats: int = 0, frames_out: np.ndarray | None = None) -> float:
        """``frames_out``: uint8 array of shape (n_threads, frame_bytes) that receives every worker's last frame (after the clock stops)."""
        table = texture_table(textures)
        out_p, stride = (ctypes.c_void_p(frames_out.ctypes.data), frames_out.shape[1]) if frames_out is not None else (None, 0)
        secs = self.lib.ncr_replay_run_threads_ex(self.api, n_threads, width, height, int(alpha),
                                                  ctypes.c_void_p(trace.ctypes.data), trace.nbytes, table, len(textures), repeats,
                                                  warm_repeats, out_p, stride)
        if secs < 0:
            raise ValueError("replay failed")
        return secs
