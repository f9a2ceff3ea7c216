"""Host logic of the present path's scaling branch, checked without a GPU: the filter tables built by the product
(csrc/swscale_filter.h — what ncr_sws_horizontal / ncr_sws_vertical consume) against an independent Python restatement of
libswscale's initFilter (bilinear), which tests/test_oracle.py pins to the real library through the C restatement's output."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT


def _log2(v):
    return max(int(v).bit_length() - 1, 0) if v > 0 else 0


def _cdiv(a, b):
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b > 0) else -q


def python_init_filter(src, dst, align, one):
    """libswscale/utils.c initFilter, SWS_BILINEAR, srcPos == dstPos == 128 (every plane of RGB -> YUV420P)."""
    x_inc = ((src << 16) + (dst >> 1)) // dst
    fone = 1 << (54 - min(_log2(src // dst), 8))
    if abs(x_inc - 0x10000) < 10:
        size, filt, pos = 1, [[fone] for _ in range(dst)], list(range(dst))
    else:
        size = 3 if x_inc <= (1 << 16) else 1 + (2 * src + dst - 1) // dst
        size = max(min(size, src - 2), 1)
        centre = ((128 * x_inc) >> 7) - ((128 * 0x10000) >> 7)
        filt, pos = [], []
        for _ in range(dst):
            xx = _cdiv(centre - (size - 2) * (1 << 16), 1 << 17)
            pos.append(xx)
            row = []
            for _j in range(size):
                d = abs(xx * (1 << 17) - centre) << 13
                if x_inc > (1 << 16):
                    d = _cdiv(d * dst, src)
                row.append(max((1 << 30) - d, 0) * (fone >> 30))
                xx += 1
            filt.append(row)
            centre += 2 * x_inc
    longest = 0
    for i in range(dst - 1, -1, -1):
        cut = 0
        for _j in range(size):
            cut += abs(filt[i][0])
            if cut > 0.002 * fone or (i < dst - 1 and pos[i] >= pos[i + 1]):
                break
            filt[i] = filt[i][1:] + [0]
            pos[i] += 1
        length, cut = size, 0
        for j in range(size - 1, 0, -1):
            cut += abs(filt[i][j])
            if cut > 0.002 * fone:
                break
            length -= 1
        longest = max(longest, length)
    if longest == 1 and align == 2:
        align = 1
    out_size = (longest + (align - 1)) & ~(align - 1)
    coef = []
    for i in range(dst):
        row = [(filt[i][j] if j < size else 0) for j in range(out_size)]
        if pos[i] < 0:
            for j in range(1, out_size):
                left = max(j + pos[i], 0)
                row[left] += row[j]
                row[j] = 0
            pos[i] = 0
        if pos[i] + out_size > src:
            shift = pos[i] + min(out_size - src, 0)
            acc = 0
            for j in range(out_size - 1, -1, -1):
                if pos[i] + j >= src:
                    acc += row[j]
                    row[j] = 0
            for j in range(out_size - 1, -1, -1):
                row[j] = 0 if j < shift else row[j - shift]
            pos[i] -= shift
            row[src - 1 - pos[i]] += acc
        total = (sum(row) + one // 2) // one or 1
        err = 0
        for j in range(out_size):
            v = row[j] + err
            q = _cdiv(v + (total >> 1) if v >= 0 else v - (total >> 1), total)
            coef.append(q)
            err = v - q * total
    return out_size, pos, coef


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    out = tmp_path_factory.mktemp("sws") / "libsws_probe.so"
    subprocess.run(["g++", "-shared", "-fPIC", "-O2", "-std=c++17", "-o", str(out), os.path.join(ROOT, "tests", "sws_filter_probe.cpp")],
                   check=True, capture_output=True)
    lib = ctypes.CDLL(str(out))
    lib.ncr_probe_sws_filter.restype = ctypes.c_int
    lib.ncr_probe_sws_filter.argtypes = (ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int)
    return lib


@pytest.mark.parametrize("src,dst", [(1920, 1280), (1080, 720), (1280, 1920), (720, 1080), (960, 640), (3840, 1920), (2160, 1080),
                                     (96, 64), (64, 96), (100, 37), (50, 200), (97, 131), (65, 71), (480, 854), (270, 481), (1920, 1920),
                                     (1080, 540), (16, 16), (24, 5), (8, 3), (3, 8), (2, 2), (5, 1)])
def test_product_filter_tables_equal_the_libswscale_restatement(probe, src, dst):
    for align, one in ((4, 1 << 14), (2, 1 << 12)):   # horizontal / vertical scaler of x86 libswscale
        size, pos, coef = python_init_filter(src, dst, align, one)
        got_pos = np.zeros(dst, dtype=np.int32)
        got_coef = np.zeros(dst * 64 + 64, dtype=np.int32)
        got_size = probe.ncr_probe_sws_filter(src, dst, align, one, got_pos.ctypes.data, got_coef.ctypes.data, got_coef.size)
        assert got_size == size
        assert got_pos.tolist() == pos
        assert got_coef[: dst * size].tolist() == coef
        if size > 1 or abs(((src << 16) + (dst >> 1)) // dst - 0x10000) >= 10:
            assert all(abs(sum(coef[i * size:(i + 1) * size]) - one) <= 1 for i in range(dst))   # rows are normalised
