"""Shared fixtures.  `-m "not gpu"`: oracle vs golden vectors, host logic, ABI surface (runs on CPU in minutes).
`-m gpu`: parity of the CUDA path against the oracles, through the C ABI."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libNativeCPURenderer.so")
PORT_LIB = os.path.join(ROOT, "oracle", "libncr_oracle.so")
REPLAY_LIB = os.path.join(ROOT, "libnativecpurenderer_b200", "lib", "libncr_replay.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the product library and the CPU checkers exist (nvcc/gcc only; no GPU needed to build)."""
    from libnativecpurenderer_b200 import build

    build.build_product()
    build.build_replayer()
    if not os.path.exists(PORT_LIB):
        build.build_oracles()


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_bilinear():
    """Digests of the reference translation unit with its own commented-out four-tap sampler (cpp:575-620) switched on
    (oracle/Makefile -> _ref/libNativeCPURenderer_bilinear.so; tests/golden/make_golden.py)."""
    with open(os.path.join(GOLDEN_DIR, "golden_bilinear.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_clip():
    """Digests of the UNMODIFIED reference drawing under an emulated clip rect (cases.ClipEmulated; tests/golden/make_golden.py)."""
    with open(os.path.join(GOLDEN_DIR, "golden_clip.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_polygon():
    """Digests of N-gon fills done by the reference's own DrawLine machinery (oracle/ref_polygon_shim.cpp; make_golden.py)."""
    with open(os.path.join(GOLDEN_DIR, "golden_polygon.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_perspective():
    """Digests of perspective quads from the shim build (spec's projective map + the reference's own bounds / sampler / ApplyPixel)."""
    with open(os.path.join(GOLDEN_DIR, "golden_perspective.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_apply_pixel():
    """Digests of the random streams that call ApplyPixel directly, from the reference's own (inline) ApplyPixel exported by the shim
    build (oracle/ref_polygon_shim.cpp; make_golden.py)."""
    with open(os.path.join(GOLDEN_DIR, "golden_apply_pixel.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def ref_polygon():
    from libnativecpurenderer_b200.binding import Renderer

    path = os.path.join(os.path.dirname(REF_LIB), "libNativeCPURenderer_polygon.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libNativeCPURenderer_polygon.so not built (reference sources absent)")
    return Renderer(path)


@pytest.fixture(scope="session")
def ref_bilinear():
    from libnativecpurenderer_b200.binding import Renderer

    path = os.path.join(os.path.dirname(REF_LIB), "libNativeCPURenderer_bilinear.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libNativeCPURenderer_bilinear.so not built (reference sources absent)")
    return Renderer(path)


@pytest.fixture(scope="session")
def image_rgba():
    """The reference's test_files/image.png as raw (128,128,4) uint8 (tests/golden/make_golden.py)."""
    return np.load(os.path.join(GOLDEN_DIR, "image_rgba.npz"))["rgba"]


@pytest.fixture(scope="session")
def port():
    from libnativecpurenderer_b200.binding import Renderer

    return Renderer(PORT_LIB)


@pytest.fixture(scope="session")
def ref():
    from libnativecpurenderer_b200.binding import Renderer

    if not os.path.exists(REF_LIB):
        pytest.skip("oracle/_ref not built (reference sources absent)")
    return Renderer(REF_LIB)


@pytest.fixture(scope="session")
def gpu():
    """The product library on a real device.  No fallback: a missing library or device fails the test."""
    from libnativecpurenderer_b200.binding import Renderer

    r = Renderer()
    probe = r.lib.CreateRenderContext(1, 1, True)
    assert probe, f"CUDA path unavailable: {r.last_error()}"
    r.lib.DestroyRenderContext(probe)
    return r
