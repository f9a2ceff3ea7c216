"""Frame-parallel batch render: Python face of ``NcrFramePool*`` / ``NcrRenderFrames`` (include/ncr_b200.h).

The reference sketches this as ``MultiThreadedVideoRenderContextPreparer`` (reference
src/libNativeCPURendererPybind.py:302-367: record the calls of N frames, replay them on a block of contexts) and leaves
``renderer()`` empty.  Here frames are recorded with :class:`trace.TraceRecorder` (same method names as ``RenderContext``),
rendered by a pool of native worker threads — one context / CUDA stream each — and delivered in frame order."""
from __future__ import annotations

import ctypes
from typing import Callable, Iterable, Sequence

import numpy as np

from . import trace as _trace

_SINK = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_long, ctypes.POINTER(ctypes.c_ubyte), ctypes.c_long)
_PRESENT = {"u8": 0, "yuv420p": 1}


def _declare(lib) -> None:
    if getattr(lib, "_ncr_batch_declared", False):
        return
    P, L, I = ctypes.c_void_p, ctypes.c_long, ctypes.c_int
    lib.NcrCreateFramePool.restype, lib.NcrCreateFramePool.argtypes = P, (L, L, I, I)
    lib.NcrCreateFramePoolOnDevices.restype, lib.NcrCreateFramePoolOnDevices.argtypes = P, (L, L, I, I, P, I)
    lib.NcrDestroyFramePool.restype, lib.NcrDestroyFramePool.argtypes = None, (P,)
    lib.NcrFramePoolWorkers.restype, lib.NcrFramePoolWorkers.argtypes = I, (P,)
    lib.NcrFramePoolRender.restype, lib.NcrFramePoolRender.argtypes = L, (P, P, P, L, P, L, I, _SINK, P)
    lib.NcrRenderFrames.restype, lib.NcrRenderFrames.argtypes = L, (L, L, I, P, P, L, P, L, I, I, _SINK, P)
    lib._ncr_batch_declared = True


def _call(renderer, traces: Sequence[np.ndarray], textures, on_frame, invoke) -> int:
    textures = list(textures)
    n = len(traces)
    ptrs = (ctypes.c_void_p * max(n, 1))(*[t.ctypes.data for t in traces])
    sizes = (ctypes.c_long * max(n, 1))(*[t.nbytes for t in traces])
    table = _trace.texture_table(textures)
    err: list[BaseException] = []

    def _sink(_user, index, pixels, nbytes):
        if on_frame is None or err:
            return
        try:
            on_frame(int(index), np.ctypeslib.as_array(pixels, shape=(int(nbytes),)))
        except BaseException as e:   # never unwind through the C frames
            err.append(e)

    rc = invoke(ptrs, sizes, n, table, len(textures), _SINK(_sink))
    if err:
        raise err[0]
    if rc == -2:
        raise ValueError("a frame does not start by overwriting the canvas (set_color): frames would not be independent")
    if rc < 0:
        raise RuntimeError(f"batch render failed: {renderer.last_error()}")
    return int(rc)


class FramePool:
    """``workers`` render contexts (one CUDA stream each) with their device and pinned buffers, kept between renders.

    ``devices`` spreads the workers over several GPUs of the box (worker k on ``devices[k % len(devices)]``): ONE host
    process — which is what the reference's ``milrenderer.py`` is — then drives all of them, with frames still delivered
    in order; textures are created once and copied to each device on first use there."""

    def __init__(self, renderer, width: int, height: int, alpha: bool, workers: int = 8, devices: Sequence[int] | None = None):
        _declare(renderer.lib)
        self._r = renderer
        if devices:
            arr = (ctypes.c_int * len(devices))(*devices)
            self._p = renderer.lib.NcrCreateFramePoolOnDevices(width, height, int(alpha), workers, arr, len(devices))
        else:
            self._p = renderer.lib.NcrCreateFramePool(width, height, int(alpha), workers)
        if not self._p:
            raise RuntimeError(f"NcrCreateFramePool failed: {renderer.last_error()}")
        self.workers = renderer.lib.NcrFramePoolWorkers(self._p)

    def render(self, traces: Sequence[np.ndarray], textures: Iterable,
               on_frame: Callable[[int, np.ndarray], None] | None = None, present: str = "u8") -> int:
        """``on_frame(index, pixels)`` runs in frame order with a uint8 view valid only during the call."""
        lib, p, mode = self._r.lib, self._p, _PRESENT[present]
        return _call(self._r, traces, textures, on_frame,
                     lambda ptrs, sizes, n, table, nt, cb: lib.NcrFramePoolRender(p, ptrs, sizes, n, table, nt, mode, cb, None))

    def close(self) -> None:
        if self._p:
            self._r.lib.NcrDestroyFramePool(self._p)
            self._p = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def render_frames(renderer, width: int, height: int, alpha: bool, traces: Sequence[np.ndarray], textures: Iterable,
                  on_frame: Callable[[int, np.ndarray], None] | None = None, workers: int = 8, present: str = "u8") -> int:
    """One-shot: render ``traces`` (one recorded frame each; every frame must start with ``set_color``) on ``workers``
    contexts created for this call.  ``present`` is "u8" (``GetBufferAsUInt8`` image) or "yuv420p" (planes)."""
    _declare(renderer.lib)
    lib, mode = renderer.lib, _PRESENT[present]
    return _call(renderer, traces, textures, on_frame,
                 lambda ptrs, sizes, n, table, nt, cb: lib.NcrRenderFrames(width, height, int(alpha), ptrs, sizes, n, table, nt,
                                                                           workers, mode, cb, None))
