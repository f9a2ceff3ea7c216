"""ctypes host mirror of the reference's Python binding for the draw/composite path.

The reference binding (reference src/libNativeCPURendererPybind.py) loads
``./libNativeCPURenderer.so`` from the current directory and declares argtypes inside
every method.  This module exposes the same classes, method names, argument meaning and
error behaviour (``RenderContext`` pyb:51-300, ``Texture`` pyb:369-440, ``Helpers``
pyb:11-49, ``get_version`` pyb:661-666), but binds one prototype table per loaded
library, so that the SAME host code can drive

* the product: ``libnativecpurenderer_b200/lib/libNativeCPURenderer.so`` (CUDA, sm_100a),
* the unmodified reference build ``oracle/_ref/libNativeCPURenderer.so`` (test oracle),
* our C restatement ``oracle/libncr_oracle.so`` (test oracle),

which is how the parity tests read like the reference's own smoke script (pyb:668-719).
The reference's unchanged binding also works against the product library; see
INTEGRATION.md.  There is no CPU fallback here: loading the product library on a box
without a usable GPU raises at the first call that needs the device.
"""
from __future__ import annotations

import ctypes
import math
import os
import random
from ctypes import c_bool, c_double, c_long, c_void_p, c_char_p, c_float, c_int, c_ulonglong, c_uint

_D = c_double
_P = c_void_p

# name -> (restype, argtypes).  C signatures: include/ncr_b200.h (reference h:83-152).
_RENDER_ABI = {
    "GetBufferSize": (c_long, (_P,)),
    "CreateRenderContext": (_P, (c_long, c_long, c_bool)),
    "DestroyRenderContext": (None, (_P,)),
    "ResizeRenderContext": (None, (_P, c_long, c_long)),
    "SaveContextState": (None, (_P,)),
    "RestoreContextState": (c_bool, (_P,)),
    "GetBuffer": (None, (_P, _P)),
    "GetBufferAsUInt8": (None, (_P, _P)),
    "CreateTexture": (_P, (c_long, c_long, c_bool, _P)),
    "CreateTextureUInt8": (_P, (c_long, c_long, c_bool, _P)),
    "DestroyTexture": (None, (_P,)),
    "CreateTextureFromRenderContext": (_P, (_P,)),
    "CreateTextureFromRenderContextShared": (_P, (_P,)),
    "SetTransform": (None, (_P,) + (_D,) * 6),
    "ApplyTransform": (None, (_P,) + (_D,) * 6),
    "Scale": (None, (_P, _D, _D)),
    "Translate": (None, (_P, _D, _D)),
    "Rotate": (None, (_P, _D)),
    "GetTransform": (None, (_P, _P)),
    "GetInverseTransform": (None, (_P, _P)),
    "SetPixel": (c_bool, (_P, c_long, c_long) + (_D,) * 4),
    "SetColorTransform": (None, (_P,) + (_D,) * 4),
    "ApplyColorTransform": (None, (_P,) + (_D,) * 4),
    "SetColor": (None, (_P,) + (_D,) * 4),
    "GetColor": (None, (_P, _D, _D, _P, _P, _P, _P)),
    "FillColor": (None, (_P,) + (_D,) * 4),
    "DrawTexture": (None, (_P, _P) + (_D,) * 4),
    "DrawSplittedTexture": (None, (_P, _P) + (_D,) * 8),
    "DrawRect": (None, (_P,) + (_D,) * 8),
    "DrawLine": (None, (_P,) + (_D,) * 9),
    "DrawCircle": (None, (_P,) + (_D,) * 7),
    "DrawVerticalGrd": (None, (_P,) + (_D,) * 12),
    "ResampleTexture": (_P, (_P, c_long, c_long)),
    "GetTextureWidth": (c_long, (_P,)),
    "GetTextureHeight": (c_long, (_P,)),
    "GetTextureEnableAlpha": (c_bool, (_P,)),
    "CreateMilthmHitEffectTexture": (_P, (_P,) + (_D,) * 5),
    "GetVersion": (c_long, ()),
}

# Exported by the product and by our C restatement, but not by the reference's -O3 build
# (inline there): bound when present.
_OPTIONAL_ABI = {
    "ApplyPixel": (c_bool, (_P, c_long, c_long) + (_D,) * 4),
    "TransformPoint": (None, (_P, _D, _D, _P, _P)),
    "GetMilthmHitEffectPixel": (None, (_D, _D, _D, _D, _P)),
}

# Additive entry points of the product library (include/ncr_b200.h, "extensions").
_EXT_ABI = {
    "NcrFlush": (c_int, (_P,)),
    "NcrLastError": (c_char_p, ()),
    "NcrDeviceName": (c_char_p, ()),
    "NcrAllocHost": (_P, (c_ulonglong,)),
    "NcrFreeHost": (None, (_P,)),
    "NcrSubmitTrace": (c_long, (_P, _P, c_long, _P, c_long)),
    "NcrMeasureD2HRate": (c_double, (c_ulonglong, c_int, c_int)),
    "NcrDeviceCount": (c_int, ()),
    "NcrCreateRenderContextOnDevice": (_P, (c_long, c_long, c_bool, c_int)),
    "NcrContextDevice": (c_int, (_P,)),
    "NcrCreateFramePoolOnDevices": (_P, (c_long, c_long, c_int, c_int, _P, c_int)),
    "NcrGetBufferAsYUV420PScaled": (c_long, (_P, c_long, c_long, _P)),
    "NcrRerunLastFlush": (c_int, (_P, c_int, c_int, _P)),
    "NcrRerunLastFlushEx": (c_int, (_P, c_int, c_int, _P, c_int, c_int)),
    "NcrGetStats": (None, (_P, _P)),
    "NcrSetStatsMode": (None, (_P, c_int)),
    "NcrSetClipRect": (None, (_P, c_long, c_long, c_long, c_long)),
    "NcrClearClipRect": (None, (_P,)),
    "NcrSetSampling": (None, (_P, c_int)),
    "NcrFillPolygon": (None, (_P, _P, c_long) + (_D,) * 4),
    "NcrDrawTexturePerspective": (None, (_P, _P, _P) + (_D,) * 4),
    "NcrKernelLaunchCount": (c_ulonglong, ()),
    "NcrMeasureF64Rate": (c_double, ()),
    "NcrDrawTextureBatch": (c_long, (_P, _P, c_long, _P, _P, _P, _P)),
    "NcrYUV420PSize": (c_long, (_P,)),
    "NcrGetBufferAsYUV420P": (c_long, (_P, _P)),
}


class NcrStats(ctypes.Structure):
    """Mirror of ``struct NcrStats`` in include/ncr_b200.h."""

    _fields_ = [
        ("n_cmds", c_ulonglong),
        ("coarse_entries", c_ulonglong),
        ("fine_entries", c_ulonglong),
        ("blended_pixels", c_ulonglong),
        ("h2d_bytes", c_ulonglong),
        ("d2h_bytes", c_ulonglong),
        ("flushes", c_ulonglong),
        ("kernel_launches", c_ulonglong),
        ("ms_bin_coarse", c_float),
        ("ms_bin_fine", c_float),
        ("ms_composite", c_float),
        ("ms_total", c_float),
        ("materialized", c_ulonglong),
        ("interior_entries", c_ulonglong),
    ]


def default_library_path() -> str:
    override = os.environ.get("NCR_LIBRARY")   # A/B builds of the product during development
    if override:
        return override
    here = os.path.dirname(os.path.abspath(__file__))
    return os.path.join(here, "lib", "libNativeCPURenderer.so")


class Renderer:
    """One loaded library exporting the reference C ABI, with classes bound to it.

    ``Renderer(path).RenderContext`` / ``.Texture`` / ``.Helpers`` have the reference
    binding's interface.  ``Renderer()`` loads the product library and raises
    ``OSError`` if it has not been built (``python -m libnativecpurenderer_b200.build``).
    """

    def __init__(self, path: str | None = None):
        self.path = os.path.abspath(path or default_library_path())
        if not os.path.exists(self.path):
            raise OSError(f"{self.path} not found: build it first (python -m libnativecpurenderer_b200.build)")
        self.lib = ctypes.CDLL(self.path)
        self.missing: list[str] = []
        for table, required in ((_RENDER_ABI, True), (_OPTIONAL_ABI, False), (_EXT_ABI, False)):
            for name, (restype, argtypes) in table.items():
                try:
                    fn = getattr(self.lib, name)
                except AttributeError:
                    if required:
                        self.missing.append(name)
                    continue
                fn.restype = restype
                fn.argtypes = argtypes
        self.is_product = hasattr(self.lib, "NcrFlush")
        outer = self

        class _BoundContext(RenderContext):
            _r = outer

        class _BoundTexture(Texture):
            _r = outer

        class _BoundHelpers(Helpers):
            _r = outer

        self.RenderContext = _BoundContext
        self.Texture = _BoundTexture
        self.Helpers = _BoundHelpers

    def get_version(self) -> int:
        return self.lib.GetVersion()

    def last_error(self) -> str:
        if not self.is_product:
            return ""
        msg = self.lib.NcrLastError()
        return msg.decode() if msg else ""

    def context_from_ptr(self, ptr: int, width: int, height: int, enable_alpha: bool) -> "RenderContext":
        """Wraps a context created through an additive entry point (e.g. ``NcrCreateRenderContextOnDevice``); owned by the wrapper."""
        ctx = self.RenderContext.__new__(self.RenderContext)
        ctx.width, ctx.height, ctx.enable_alpha = width, height, enable_alpha
        ctx._lib = self.lib
        ctx._ptr = ptr
        ctx._can_release = True
        return ctx

    def context_on_device(self, width: int, height: int, enable_alpha: bool, device: int) -> "RenderContext":
        ptr = self.lib.NcrCreateRenderContextOnDevice(width, height, enable_alpha, device)
        if not ptr:
            raise RuntimeError(f"NcrCreateRenderContextOnDevice failed: {self.last_error()}")
        return self.context_from_ptr(ptr, width, height, enable_alpha)

    def texture_from_ptr(self, ptr: int) -> "Texture":
        tex = self.Texture.__new__(self.Texture)
        tex._ptr = ptr
        tex._update_props()
        return tex


def _as_void_p(buf) -> c_void_p:
    """Address of a writable/readable Python buffer (bytearray, numpy array, ctypes array)."""
    if hasattr(buf, "ctypes"):  # numpy
        return c_void_p(buf.ctypes.data)
    if isinstance(buf, (bytes, bytearray, memoryview)):
        return ctypes.cast((ctypes.c_char * len(buf)).from_buffer(buf), c_void_p)
    return ctypes.cast(buf, c_void_p)


class RenderContext:
    """Canvas + transform / colour state (reference pyb:51-300, cpp:7-57, 277-309, 386-492)."""

    _r: Renderer = None  # bound by Renderer

    def __init__(self, width: int, height: int, enable_alpha: bool):
        self.width = width
        self.height = height
        self.enable_alpha = enable_alpha
        self._lib = self._r.lib
        self._ptr = self._lib.CreateRenderContext(width, height, enable_alpha)
        if not self._ptr:
            raise RuntimeError(f"CreateRenderContext failed: {self._r.last_error()}")
        self._can_release = True

    def __del__(self):
        if getattr(self, "_can_release", False) and getattr(self, "_ptr", 0):
            self._lib.DestroyRenderContext(self._ptr)
            self._ptr = 0

    # --- readback (cpp:3-5, 311-316, 52-57) ---
    def get_buffer_size(self) -> int:
        return self._lib.GetBufferSize(self._ptr)

    def get_buffer(self):
        out = (c_double * self.get_buffer_size())()
        self._lib.GetBuffer(self._ptr, ctypes.byref(out))
        return list(out)

    def get_buffer_np(self):
        import numpy as np

        out = np.empty(self.get_buffer_size(), dtype=np.float64)
        self._lib.GetBuffer(self._ptr, _as_void_p(out))
        return out

    def get_buffer_as_uint8(self) -> bytearray:
        out = bytearray(self.get_buffer_size())
        if len(out):
            self._lib.GetBufferAsUInt8(self._ptr, _as_void_p(out))
        return out

    def get_buffer_as_uint8_into(self, address: int) -> None:
        """Readback into caller-owned memory (e.g. pinned memory from ``NcrAllocHost``)."""
        self._lib.GetBufferAsUInt8(self._ptr, c_void_p(address))

    def as_pilimg(self):
        from PIL import Image

        mode = "RGBA" if self.enable_alpha else "RGB"
        return Image.frombytes(mode, (self.width, self.height), bytes(self.get_buffer_as_uint8()))

    def resize(self, width: int, height: int):
        self._lib.ResizeRenderContext(self._ptr, width, height)
        self.width = width
        self.height = height

    # --- state machine (cpp:277-309, 386-492, 623-641) ---
    def save_state(self):
        self._lib.SaveContextState(self._ptr)

    def restore_state(self):
        return self._lib.RestoreContextState(self._ptr)

    def set_transform(self, a, b, c, d, e, f):
        self._lib.SetTransform(self._ptr, a, b, c, d, e, f)

    def apply_transform(self, a, b, c, d, e, f):
        self._lib.ApplyTransform(self._ptr, a, b, c, d, e, f)

    def scale(self, sx, sy):
        self._lib.Scale(self._ptr, sx, sy)

    def translate(self, tx, ty):
        self._lib.Translate(self._ptr, tx, ty)

    def rotate(self, angle):
        self._lib.Rotate(self._ptr, angle)

    def rotate_degree(self, deg):
        self.rotate(deg * math.pi / 180)

    def get_transform(self):
        out = (c_double * 6)()
        self._lib.GetTransform(self._ptr, ctypes.byref(out))
        return tuple(out)

    def get_inverse_transform(self):
        out = (c_double * 6)()
        self._lib.GetInverseTransform(self._ptr, ctypes.byref(out))
        return tuple(out)

    def set_color_transform(self, r, g, b, a):
        self._lib.SetColorTransform(self._ptr, r, g, b, a)

    def apply_color_transform(self, r, g, b, a):
        self._lib.ApplyColorTransform(self._ptr, r, g, b, a)

    # --- pixel writes (cpp:494-549, 643-691) ---
    def set_pixel(self, x: int, y: int, r, g, b, a):
        return self._lib.SetPixel(self._ptr, x, y, r, g, b, a)

    def apply_pixel(self, x: int, y: int, r, g, b, a):
        return self._lib.ApplyPixel(self._ptr, x, y, r, g, b, a)

    def set_color(self, r, g, b, a):
        self._lib.SetColor(self._ptr, r, g, b, a)

    def get_color(self, x: float, y: float):
        # The C signature takes f64 coordinates (h:113); the reference binding passes c_long (pyb:258),
        # which is an ABI mismatch there.  This mirror follows the header.
        out = [c_double() for _ in range(4)]
        self._lib.GetColor(self._ptr, x, y, *[ctypes.byref(o) for o in out])
        return tuple(o.value for o in out)

    def fill_color(self, r, g, b, a):
        self._lib.FillColor(self._ptr, r, g, b, a)

    # --- primitives (cpp:720-948, 1285-1316) ---
    def draw_texture(self, tex: "Texture", x, y, w, h):
        self._lib.DrawTexture(self._ptr, tex._ptr, x, y, w, h)

    def draw_splitted_texture(self, tex: "Texture", x, y, width, height, u_start, u_end, v_start, v_end):
        self._lib.DrawSplittedTexture(self._ptr, tex._ptr, x, y, width, height, u_start, u_end, v_start, v_end)

    def draw_rect(self, x, y, width, height, r, g, b, a):
        self._lib.DrawRect(self._ptr, x, y, width, height, r, g, b, a)

    def draw_line(self, x0, y0, x1, y1, width, r, g, b, a):
        self._lib.DrawLine(self._ptr, x0, y0, x1, y1, width, r, g, b, a)

    def draw_circle(self, x, y, radius, r, g, b, a):
        self._lib.DrawCircle(self._ptr, x, y, radius, r, g, b, a)

    def draw_vertical_grd(self, x, y, width, height, top_r, top_g, top_b, top_a, bottom_r, bottom_g, bottom_b, bottom_a):
        self._lib.DrawVerticalGrd(self._ptr, x, y, width, height, top_r, top_g, top_b, top_a,
                                  bottom_r, bottom_g, bottom_b, bottom_a)

    def draw_vertical_mut_grd(self, x, y, width, height, steps):
        # pyb:272-280: consecutive (position, rgba) stops become stacked two-colour gradients.
        for (p0, c0), (p1, c1) in zip(steps, steps[1:]):
            self.draw_vertical_grd(x, y + height * p0, width, height * (p1 - p0), *c0[:4], *c1[:4])

    # --- canvas -> texture (cpp:362-384) ---
    def as_texure(self):  # (sic) the reference spells it this way, pyb:282
        return self._r.texture_from_ptr(self._lib.CreateTextureFromRenderContext(self._ptr))

    as_texture = as_texure

    def as_texture_shared(self):
        tex = self._r.texture_from_ptr(self._lib.CreateTextureFromRenderContextShared(self._ptr))
        tex._can_release = False
        return tex

    # --- product-only extensions (include/ncr_b200.h) ---
    def flush(self) -> None:
        if self._lib.NcrFlush(self._ptr) != 0:
            raise RuntimeError(self._r.last_error())

    def set_clip_rect(self, x: int, y: int, w: int, h: int):
        self._lib.NcrSetClipRect(self._ptr, x, y, w, h)

    def clear_clip_rect(self):
        self._lib.NcrClearClipRect(self._ptr)

    def set_sampling(self, mode: int):
        self._lib.NcrSetSampling(self._ptr, mode)

    def fill_polygon(self, points, r, g, b, a):
        flat = [float(v) for p in points for v in p]
        arr = (c_double * len(flat))(*flat)
        self._lib.NcrFillPolygon(self._ptr, arr, len(flat) // 2, r, g, b, a)

    def get_buffer_as_yuv420p(self, out=None):
        """Present path: planar Y, U, V of the canvas as one uint8 array (pinned to libswscale, see include/ncr_b200.h)."""
        import numpy as np

        n = self._lib.NcrYUV420PSize(self._ptr)
        if out is None:
            out = np.empty(n, dtype=np.uint8)
        got = self._lib.NcrGetBufferAsYUV420P(self._ptr, _as_void_p(out))
        if got != n:
            raise RuntimeError("NcrGetBufferAsYUV420P failed")
        return out

    def get_buffer_as_yuv420p_scaled(self, dst_w: int, dst_h: int):
        """The planes at another size: what PutRendererContextFrame hands the encoder when the VideoCap's size differs from the canvas's."""
        import numpy as np

        n = dst_w * dst_h + 2 * ((dst_w + 1) // 2) * ((dst_h + 1) // 2)
        out = np.empty(n, dtype=np.uint8)
        got = self._lib.NcrGetBufferAsYUV420PScaled(self._ptr, dst_w, dst_h, _as_void_p(out))
        if got != n:
            raise RuntimeError(f"NcrGetBufferAsYUV420PScaled returned {got}, expected {n}: {self._r.last_error()}")
        return out

    def get_buffer_as_yuv420p_into(self, address: int) -> int:
        return self._lib.NcrGetBufferAsYUV420P(self._ptr, c_void_p(address))

    def draw_texture_batch(self, tex: "Texture", xywh, transforms=None, color_transforms=None, uv=None) -> int:
        """n sprites in one FFI crossing (NcrDrawTextureBatch): per sprite save_state, apply_transform(transforms[k]),
        apply_color_transform(color_transforms[k]), draw_texture / draw_splitted_texture(uv[k]), restore_state."""
        import numpy as np

        def arr(a, cols):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1, cols)
            if a.shape[0] != n:
                raise ValueError("array lengths differ")
            return a

        xywh = np.ascontiguousarray(xywh, dtype=np.float64).reshape(-1, 4)
        n = xywh.shape[0]
        m, ct, uvs = arr(transforms, 6), arr(color_transforms, 4), arr(uv, 4)
        ptr = lambda a: None if a is None else c_void_p(a.ctypes.data)  # noqa: E731
        return self._lib.NcrDrawTextureBatch(self._ptr, tex._ptr, n, ptr(m), ptr(ct), ptr(xywh), ptr(uvs))

    def draw_texture_perspective(self, tex: "Texture", inv_h, x, y, w, h):
        arr = (c_double * 9)(*[float(v) for v in inv_h])
        self._lib.NcrDrawTexturePerspective(self._ptr, tex._ptr, arr, x, y, w, h)

    def stats(self) -> NcrStats:
        st = NcrStats()
        self._lib.NcrGetStats(self._ptr, ctypes.byref(st))
        return st

    def set_stats_mode(self, mode: int):
        self._lib.NcrSetStatsMode(self._ptr, mode)


class Texture:
    """Immutable image (reference pyb:369-440, cpp:318-360, 950-988)."""

    _r: Renderer = None

    def __init__(self, width: int, height: int, enableAlpha: bool, data, is_uint8: bool = True):
        if width * height * (4 if enableAlpha else 3) * (1 if is_uint8 else 8) != len(data):
            raise ValueError("data size not match")
        self.width = width
        self.height = height
        self.enableAlpha = enableAlpha
        lib = self._r.lib
        raw = bytearray(data)
        fn = lib.CreateTextureUInt8 if is_uint8 else lib.CreateTexture
        self._ptr = fn(width, height, enableAlpha, _as_void_p(raw) if len(raw) else None)
        if not self._ptr:
            raise RuntimeError(f"texture creation failed: {self._r.last_error()}")

    def __del__(self):
        ptr = getattr(self, "_ptr", 0)
        if ptr:
            self._r.lib.DestroyTexture(ptr)
            self._ptr = 0

    def _update_props(self):
        lib = self._r.lib
        self.width = lib.GetTextureWidth(self._ptr)
        self.height = lib.GetTextureHeight(self._ptr)
        self.enableAlpha = lib.GetTextureEnableAlpha(self._ptr)

    def resample(self, width: int, height: int) -> "Texture":
        return self._r.texture_from_ptr(self._r.lib.ResampleTexture(self._ptr, width, height))

    @classmethod
    def from_pilimg(cls, img):
        from PIL import Image

        if not isinstance(img, Image.Image):
            raise TypeError("img must be a PIL.Image.Image")
        if img.mode not in ("RGB", "RGBA"):
            img = img.convert("RGBA")
        return cls(img.width, img.height, img.mode == "RGBA", img.tobytes())

    @classmethod
    def from_numpy(cls, arr):
        """(h, w, 3|4) uint8 or float64 array."""
        h, w, ch = arr.shape
        return cls(w, h, ch == 4, arr.tobytes(), is_uint8=(arr.dtype.itemsize == 1))


class Helpers:
    """reference pyb:11-49."""

    _r: Renderer = None

    @classmethod
    def create_milthm_hit_effect_textures(cls, mask: Texture, n: int, seed: float | None = None):
        seed = random.random() if seed is None else seed
        fn = cls._r.lib.CreateMilthmHitEffectTexture
        return [cls._r.texture_from_ptr(fn(mask._ptr, seed, i / (n - 1), 0x96 / 0xFF, 0x90 / 0xFF, 0xFD / 0xFF))
                for i in range(n)]
