"""Frame-parallel batch render: Python face of ``NcrRenderFrames`` (include/ncr_b200.h).

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


def render_frames(renderer, width: int, height: int, alpha: bool, traces: Sequence[np.ndarray], textures: Iterable,
                  on_frame: Callable[[int, np.ndarray], None] | None = None, workers: int = 8, present: str = "u8") -> int:
    """Render ``traces`` (one recorded frame each; every frame must start with ``set_color``) on ``workers`` contexts.

    ``on_frame(index, pixels)`` is called in frame order with a uint8 view that is only valid during the call (copy it to
    keep it).  ``present`` is "u8" (the ``GetBufferAsUInt8`` image) or "yuv420p" (``NcrGetBufferAsYUV420P`` planes).
    Returns the number of frames rendered; raises on device errors or frames that are not independent."""
    lib = renderer.lib
    fn = lib.NcrRenderFrames
    fn.restype = ctypes.c_long
    fn.argtypes = (ctypes.c_long, ctypes.c_long, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p,
                   ctypes.c_long, ctypes.c_int, ctypes.c_int, _SINK, ctypes.c_void_p)
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

    cb = _SINK(_sink)
    rc = fn(width, height, int(alpha), ptrs, sizes, n, table, len(textures), workers, {"u8": 0, "yuv420p": 1}[present], cb, None)
    if err:
        raise err[0]
    if rc == -2:
        raise ValueError("a frame does not start by overwriting the canvas (set_color): frames would not be independent")
    if rc < 0:
        raise RuntimeError(f"NcrRenderFrames failed: {renderer.last_error()}")
    return int(rc)
