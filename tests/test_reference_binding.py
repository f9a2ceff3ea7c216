"""The reference's own, unchanged ctypes binding (src/libNativeCPURendererPybind.py, staged by oracle/Makefile into the
git-ignored oracle/_ref/pyb/) driving (a) the unmodified reference build — CPU, validates the harness against the committed
digests — and (b) the product library on the GPU: the real drop-in client, per-call ctypes, loading ./libNativeCPURenderer.so
from its working directory (pyb:9)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import GOLDEN_DIR, REF_LIB, ROOT

PYB_DIR = os.path.join(ROOT, "oracle", "_ref", "pyb")
PYB = os.path.join(PYB_DIR, "libNativeCPURendererPybind.py")
DRIVER = os.path.join(ROOT, "tests", "pyb_driver.py")

needs_pyb = pytest.mark.skipif(not os.path.exists(PYB), reason="oracle/_ref/pyb not staged (run `make -C oracle ref` where /root/reference exists)")


def run_through_reference_binding(library: str, workdir, cases, extra_env=None) -> dict:
    os.symlink(library, os.path.join(workdir, "libNativeCPURenderer.so"))
    env = dict(os.environ, PYTHONPATH=PYB_DIR + os.pathsep + os.environ.get("PYTHONPATH", ""))
    env.update(extra_env or {})
    res = subprocess.run([sys.executable, DRIVER, os.path.join(GOLDEN_DIR, "image_rgba.npz"), *cases], capture_output=True,
                         text=True, cwd=workdir, env=env, timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("PYB_RESULT ")][-1]
    return json.loads(line[len("PYB_RESULT "):])


@needs_pyb
def test_reference_binding_with_reference_build_reproduces_the_goldens(tmp_path, golden):
    if not os.path.exists(REF_LIB):
        pytest.skip("oracle/_ref not built")
    # the symlinked library's $ORIGIN is the scratch directory: point the loader at the FFmpeg stub next to the real file
    out = run_through_reference_binding(REF_LIB, str(tmp_path), ["k2"], {"LD_LIBRARY_PATH": os.path.dirname(REF_LIB)})
    assert out["version"] == 1
    assert out["k2"] == golden["k2"]


@needs_pyb
@pytest.mark.gpu
def test_reference_binding_drives_the_product_on_the_gpu(tmp_path, golden):
    """K1 (BASELINE config 1: 1,000 quads via the ctypes binding) and the binding's smoke loop K2, rendered by the product
    through the reference's unchanged Python file; digests equal the unmodified reference build's."""
    from libnativecpurenderer_b200.binding import default_library_path

    out = run_through_reference_binding(default_library_path(), str(tmp_path), ["k1", "k2"])
    assert out["version"] == 1
    assert out["k1"]["u8"] == golden["k1"]["u8"]
    assert out["k1"]["pil_size"] == [1920, 1080]
    assert out["k2"] == golden["k2"]
