"""CPU suite, part 3: the N>1 host logic (frame -> rank assignment, max-over-ranks aggregation) under a real
world_size-2 process group (gloo), plus the trace recorder / replayer host code against the C restatement."""
import os
import socket
import sys

import numpy as np
import pytest

from libnativecpurenderer_b200 import sharding, streams, trace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_frame_partitions_are_disjoint_and_complete():
    for n in (0, 1, 7, 36000):
        for world in (1, 2, 4, 8):
            for mode in ("interleave", "block"):
                got = sorted(f for r in range(world) for f in sharding.frames_for_rank(n, r, world, mode))
                assert got == list(range(n))
                sizes = [len(sharding.frames_for_rank(n, r, world, mode)) for r in range(world)]
                assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.frames_for_rank(10, 2, 2)


def _worker(rank, world, port, out):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = sharding.frames_for_rank(10, rank, world)
    secs = 2.0 if rank == 0 else 4.0   # rank 1 is slower: the job's time is the max
    dist.barrier()
    total = sharding.aggregate_throughput(len(frames), secs, dist)
    out.put((rank, len(frames), total))
    dist.destroy_process_group()


def test_aggregate_throughput_world_size_2_gloo():
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [5, 5]
    assert all(abs(r[2] - 10 / 4.0) < 1e-12 for r in res)   # 10 frames / slowest rank's 4 s


def test_trace_round_trip_against_the_port(port, image_rgba):
    """Recorder -> bytes -> C replayer -> library renders exactly what direct calls render."""
    from conftest import REPLAY_LIB
    import cases

    tex_np = streams.make_c2_textures()
    tex = [port.Texture.from_numpy(t) for t in tex_np]
    direct = port.RenderContext(320, 200, True)
    streams.stream_c2(direct, tex, n=300)
    rec = trace.TraceRecorder(320, 200, True)
    streams.stream_c2(rec, [trace.TexSlot(k, t.shape[1], t.shape[0]) for k, t in enumerate(tex_np)], n=300)
    rec.present()
    arr = rec.as_array()
    assert arr.ctypes.data % 8 == 0 and rec.n_draws == 300
    via = port.RenderContext(320, 200, True)
    frame = np.zeros(320 * 200 * 4, dtype=np.uint8)
    rp = trace.Replayer(REPLAY_LIB, port.path)
    rp.run(via, arr, tex, frame_address=frame.ctypes.data)
    assert cases.digest(via) == cases.digest(direct)
    assert frame.tobytes() == bytes(direct.get_buffer_as_uint8())   # PRESENT record read the frame back
    assert rp.run_threads(2, 320, 200, True, arr, tex, repeats=1, warm_repeats=1) > 0
    with pytest.raises(ValueError):
        rp.run(via, arr[:-3], tex)   # truncated stream is rejected, not executed past the end
