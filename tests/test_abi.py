"""CPU suite, part 2: the drop-in boundary.  The product library must load without a GPU, export every
symbol include/ncr_b200.h declares (which covers every function of the reference header h:83-152), and
fail loudly — not fall back — when no CUDA device is usable."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ncr_b200.h")

REFERENCE_EXPORTS = """GetBufferSize CreateRenderContext DestroyRenderContext CreateVideoCap InitializeVideoCap DestroyVideoCap
PutRendererContextFrame ReleaseVideoCap SaveContextState RestoreContextState GetBuffer GetBufferAsUInt8 CreateTexture
CreateTextureUInt8 DestroyTexture CreateTextureFromRenderContext SetTransform ApplyTransform Scale Translate Rotate
TransformPoint GetTransform GetInverseTransform SetPixel ApplyPixel SetColorTransform ApplyColorTransform SetColor GetColor
FillColor DrawTexture DrawRect DrawLine DrawCircle ResampleTexture GetTextureWidth GetTextureHeight GetTextureEnableAlpha
GetAudioClipBufferSizeFromData GetAudioClipBufferSize CreateAudioClipFromBuffer CreateAudioClipFromInt16Buffer
CreateSilentAudioClip DestroyAudioClip CloneAudioClip ApplyResampleAudioClip ResampleAudioClipLike OverlayAudioClip
OverlayAudioClipSecond SaveAudioClipAsWav GetAudioClipSampleRate GetAudioClipChannels GetAudioClipNumFrames
GetAudioClipDuration GetWapperedBytesDataPtr GetWapperedBytesDataSize ApplyVolumeGain PutAudioIntoVideoCap GetVersion
ApplyCutAudioClip ApplySpeedAudioClip DrawVerticalGrd DrawSplittedTexture CreateTextureFromRenderContextShared
ResizeRenderContext GetMilthmHitEffectPixel CreateMilthmHitEffectTexture""".split()   # reference h:84-151, 68 functions


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b([A-Z][A-Za-z0-9]+)\s*\(", text)) - {"NCR"})


def product_lib():
    from libnativecpurenderer_b200.binding import default_library_path

    return ctypes.CDLL(default_library_path())


def test_header_declares_every_reference_function():
    declared = set(declared_functions())
    assert len(REFERENCE_EXPORTS) == 68
    assert not [n for n in REFERENCE_EXPORTS if n not in declared]


def test_library_exports_every_declared_symbol():
    lib = product_lib()
    names = declared_functions()
    assert len(names) >= 80
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_tables_match_the_library():
    from libnativecpurenderer_b200.binding import Renderer

    r = Renderer()
    assert r.missing == [] and r.is_product
    assert r.get_version() == 1


def test_library_has_sm100a_code_only():
    from libnativecpurenderer_b200.binding import default_library_path

    out = subprocess.run(["cuobjdump", "-lelf", default_library_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, out


def test_host_only_entry_points_work_without_a_device():
    """Audio/WAV helpers are plain host code (reference cpp:990-1283) and must not touch CUDA."""
    lib = product_lib()
    lib.CreateSilentAudioClip.restype = ctypes.c_void_p
    lib.CreateSilentAudioClip.argtypes = (ctypes.c_long, ctypes.c_long, ctypes.c_long)
    lib.SaveAudioClipAsWav.restype = ctypes.c_void_p
    lib.SaveAudioClipAsWav.argtypes = (ctypes.c_void_p,)
    lib.GetWapperedBytesDataSize.restype = ctypes.c_long
    lib.GetWapperedBytesDataSize.argtypes = (ctypes.c_void_p,)
    lib.GetWapperedBytesDataPtr.restype = ctypes.c_void_p
    lib.GetWapperedBytesDataPtr.argtypes = (ctypes.c_void_p,)
    clip = lib.CreateSilentAudioClip(8000, 2, 100)
    wav = lib.SaveAudioClipAsWav(clip)
    n = lib.GetWapperedBytesDataSize(wav)
    assert n == 44 + 100 * 2 * 2
    raw = ctypes.string_at(lib.GetWapperedBytesDataPtr(wav), n)
    assert raw[:4] == b"RIFF" and raw[8:16] == b"WAVEfmt " and raw[36:40] == b"data"


def test_no_gpu_fails_loudly_not_silently():
    """In a process that cannot see a CUDA device the product returns NULL and says why; nothing renders on the CPU."""
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from libnativecpurenderer_b200.binding import Renderer\n"
        "r = Renderer()\n"
        "p = r.lib.CreateRenderContext(4, 4, True)\n"
        "print('PTR', p, '|', r.last_error())\n" % ROOT
    )
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stderr
    assert "PTR None |" in res.stdout
    assert "no usable CUDA device" in res.stdout


REF_PYB = "/root/reference/src/libNativeCPURendererPybind.py"


@pytest.mark.skipif(not os.path.exists(REF_PYB), reason="reference checkout not present (GPU box)")
def test_reference_binding_imports_against_the_product_unchanged(tmp_path):
    """The reference's own ctypes binding (pyb:9 loads ./libNativeCPURenderer.so with RTLD_NOW) must import against the
    product library from a directory that holds nothing but that file, and reach host-only entry points through it."""
    from libnativecpurenderer_b200.binding import default_library_path

    os.symlink(default_library_path(), tmp_path / "libNativeCPURenderer.so")
    code = (
        "import sys; sys.path.insert(0, '/root/reference/src')\n"
        "import libNativeCPURendererPybind as CPURenderer\n"
        "print('VERSION', CPURenderer.get_version())\n"
        "clip = CPURenderer.AudioClip.slient(8000, 2, 50)\n"
        "print('WAV', len(clip.save_as_wav()), clip.duration)\n"
        "missing = [n for n in %r if not hasattr(CPURenderer.lib, n)]\n"
        "print('MISSING', missing)\n" % (REFERENCE_EXPORTS,)
    )
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=tmp_path, timeout=300)
    assert res.returncode == 0, res.stderr
    assert "VERSION 1" in res.stdout and "MISSING []" in res.stdout
    assert "WAV 244 0.00625" in res.stdout


def test_batch_render_validates_frames_before_touching_a_device():
    """NcrRenderFrames: an empty batch is a no-op and a frame that does not start by overwriting the canvas
    is refused (-2) — both decided on the host, before any context is created."""
    from libnativecpurenderer_b200 import batch, trace
    from libnativecpurenderer_b200.binding import Renderer

    R = Renderer()
    assert batch.render_frames(R, 64, 48, True, [], []) == 0
    rec = trace.TraceRecorder(64, 48, True)
    rec.translate(3, 4)
    rec.fill_color(1, 0, 0, .5)   # reads the previous canvas: not an independent frame
    rec.present()
    with pytest.raises(ValueError):
        batch.render_frames(R, 64, 48, True, [rec.as_array()], [])
