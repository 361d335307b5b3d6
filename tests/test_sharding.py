"""Multi-GPU host logic on CPU: world_size-2 gloo run of the stream sharding + transcript gather
(nemotron-speech.cpp_b200/sharding.py). Each rank decodes ITS streams with the CPU oracle standing in for the
engine; rank 0 must end up with exactly the single-process result."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _decode(ids):
    import oracle as O
    import synth
    m = O.Model(synth.cached_model("f32", 2, R=0))
    out = {}
    for s in ids:
        st = O.Stream(m, 0)
        st.push(synth.synth_pcm(300 + s, 0.9))
        out[s] = st.tokens().tolist()
    return out


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import torch.distributed as dist
    import nsb200  # noqa: F401  (registers the package under its importable name)
    from nemotron_speech_cpp_b200 import sharding
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        merged = sharding.run_sharded(list(range(5)), rank, world, _decode)
        dist.barrier()
        if rank == 0:
            q.put(merged)
    finally:
        dist.destroy_process_group()


def test_owner_mapping_is_sticky_and_balanced():
    sys.path.insert(0, ROOT)
    import nsb200  # noqa: F401
    from nemotron_speech_cpp_b200 import sharding
    for G in (1, 2, 4, 8):
        shares = [sharding.local_streams(range(1024), r, G) for r in range(G)]
        assert sorted(sum(shares, [])) == list(range(1024))
        assert max(map(len, shares)) - min(map(len, shares)) <= 1
        assert all(sharding.owner_of(s, G) == r for r, sh in enumerate(shares) for s in sh)
    with pytest.raises(ValueError):
        sharding.owner_of(0, 0)


def test_two_rank_gloo_gather_equals_single_process(built):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single = _decode(list(range(5)))
    assert merged == single
    assert sum(len(v) for v in single.values()) > 0


def test_bench_workload_mixes_recordings_on_every_rank():
    """bench.py builds its streams from 8 synthetic recordings keyed by the GLOBAL stream id; under `stream s -> rank s mod G` every rank of
    a 1 / 2 / 4 / 8-GPU run must still see all 8 recordings in equal shares (an 8-GPU run once gave every stream of a rank the same
    recording: whole batches burst together and the max over ranks measured the burstiest recording, not the workload)."""
    import collections
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from nemotron_speech_cpp_b200 import sharding
    for world in (1, 2, 4, 8):
        for per_gpu in (64, 128):
            total = per_gpu * world
            for rank in range(world):
                mine = sharding.local_streams(range(total), rank, world)
                assert len(mine) == per_gpu
                cnt = collections.Counter(bench.stream_recording(g) for g in mine)
                assert len(cnt) == 8 and max(cnt.values()) == min(cnt.values()), (world, rank, cnt)
                pairs = {(bench.stream_recording(g), g // 8) for g in mine}
                assert len(pairs) == per_gpu                      # no two streams of a rank are the same (recording, shift)
