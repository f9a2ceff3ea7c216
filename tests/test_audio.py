"""CPU suite, part 4: the audio-clip helpers the product re-exports as plain host code
(csrc/host_misc.cpp) against the UNMODIFIED reference build (oracle/_ref; reference cpp:990-1283).  The observable is
the WAV byte stream SaveAudioClipAsWav produces after each operation, plus the clip's getters."""
import ctypes
import os

import numpy as np
import pytest

from conftest import REF_LIB
from libnativecpurenderer_b200 import build

P, L, D = ctypes.c_void_p, ctypes.c_long, ctypes.c_double


def _load(path):
    lib = ctypes.CDLL(path)
    sig = {
        "CreateAudioClipFromBuffer": (P, (L, L, L, P)), "CreateAudioClipFromInt16Buffer": (P, (L, L, L, P)),
        "CreateSilentAudioClip": (P, (L, L, L)), "CloneAudioClip": (P, (P,)), "ApplyResampleAudioClip": (None, (P, L, L)),
        "ResampleAudioClipLike": (None, (P, P)), "OverlayAudioClip": (L, (P, P, L, ctypes.c_bool)),
        "OverlayAudioClipSecond": (L, (P, P, D, ctypes.c_bool)), "SaveAudioClipAsWav": (P, (P,)),
        "GetAudioClipSampleRate": (L, (P,)), "GetAudioClipChannels": (L, (P,)), "GetAudioClipNumFrames": (L, (P,)),
        "GetAudioClipDuration": (D, (P,)), "GetAudioClipBufferSize": (L, (P,)), "GetWapperedBytesDataPtr": (P, (P,)),
        "GetWapperedBytesDataSize": (L, (P,)), "ApplyVolumeGain": (None, (P, D)), "ApplyCutAudioClip": (None, (P, L, L)),
        "ApplySpeedAudioClip": (None, (P, D)),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


def _observe(lib, clip):
    wav = lib.SaveAudioClipAsWav(clip)
    n = lib.GetWapperedBytesDataSize(wav)
    return (lib.GetAudioClipSampleRate(clip), lib.GetAudioClipChannels(clip), lib.GetAudioClipNumFrames(clip),
            lib.GetAudioClipDuration(clip), lib.GetAudioClipBufferSize(clip), ctypes.string_at(lib.GetWapperedBytesDataPtr(wav), n))


def _script(lib):
    rs = np.random.RandomState(3)
    a = np.ascontiguousarray(rs.uniform(-0.9, 0.9, 4410 * 2))
    b16 = np.ascontiguousarray(rs.randint(-20000, 20000, 3000).astype(np.int16))
    out = []
    ca = lib.CreateAudioClipFromBuffer(44100, 2, 4410, a.ctypes.data)
    cb = lib.CreateAudioClipFromInt16Buffer(22050, 1, 3000, b16.ctypes.data)
    out += [_observe(lib, ca), _observe(lib, cb)]
    cc = lib.CloneAudioClip(ca)
    lib.ApplyVolumeGain(cc, 0.37)
    out.append(_observe(lib, cc))
    lib.ApplyResampleAudioClip(cc, 48000, 1)          # rate up, stereo -> mono
    out.append(_observe(lib, cc))
    cd = lib.CloneAudioClip(cb)
    lib.ResampleAudioClipLike(cd, ca)                  # rate up, mono -> stereo
    out.append(_observe(lib, cd))
    rc1 = lib.OverlayAudioClip(ca, cd, 1000, False)    # same format
    rc2 = lib.OverlayAudioClip(ca, cb, 500, True)      # auto-resampled source
    rc3 = lib.OverlayAudioClipSecond(ca, cd, 0.05, False)
    rc4 = lib.OverlayAudioClip(ca, cb, 10, False)      # format mismatch without auto-resample: error code
    out += [(rc1, rc2, rc3, rc4), _observe(lib, ca)]
    lib.ApplyCutAudioClip(ca, 300, 3900)
    out.append(_observe(lib, ca))
    lib.ApplySpeedAudioClip(ca, 1.25)
    out.append(_observe(lib, ca))
    silent = lib.CreateSilentAudioClip(8000, 2, 64)
    lib.OverlayAudioClip(silent, lib.CreateSilentAudioClip(8000, 2, 200), 32, False)   # source longer than the target's tail
    out.append(_observe(lib, silent))
    return out


def test_audio_helpers_match_the_reference_build():
    if not os.path.exists(REF_LIB):
        pytest.skip("oracle/_ref not built (reference sources absent)")
    got = _script(_load(build.LIB))
    want = _script(_load(REF_LIB))
    assert len(got) == len(want)
    for k, (g, w) in enumerate(zip(got, want)):
        assert g == w, f"step {k} differs"
