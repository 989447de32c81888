"""N>1 path on CPU: two gloo ranks shard a sequence by mini-GOP, run the (oracle) ME on their own pictures and agree on
the step time / total; the union of the shards must equal the single-process result, with no exchange on the data path."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H

sharding = importlib.import_module("svt-vp9_b200.sharding")


def test_encode_order_and_refs():
    order = sharding.minigop_encode_order(4)
    assert sorted(order) == list(range(1, 17)) and order[0] == 16 and order[1] == 8
    seen = {0}
    for o in order:                       # every picture is encoded after both of its references
        r0, r1 = sharding.minigop_refs(o)
        assert r0 in seen and r1 in seen, (o, r0, r1)
        seen.add(o)
    assert [sharding.temporal_layer(o) for o in (16, 8, 4, 12, 2, 1)] == [0, 1, 2, 2, 3, 4]


@pytest.mark.parametrize("n_pictures,world", [(33, 2), (40, 3), (17, 4), (1, 2), (100, 8)])
def test_shards_partition_the_sequence(n_pictures, world):
    owned = []
    for r in range(world):
        for first, count in sharding.shard_minigops(n_pictures, world, r):
            assert (first - 1) % 16 == 0 and 1 <= count <= 16
            owned += list(range(first, first + count))
    assert sorted(owned) == list(range(1, n_pictures))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _me_digest(clip_planes, structs, display, P_of):
    """Oracle ME of one picture (display number) against its two references -> crc of the result bytes."""
    import zlib
    base = ((display - 1) // 4) * 4
    off = display - base
    r0, r1 = sharding.minigop_refs(off, 2)
    P = P_of(sharding.temporal_layer(off, 2) , (base + r0, base + r1))
    n_sb = ((P.luma_width + 63) // 64) * ((P.luma_height + 63) // 64)
    out = np.zeros(n_sb * 85, dtype=H.api.ME_RESULT_DTYPE)
    H.oracle().svtb200_oracle_me_picture(P, structs[display], structs[base + r0], structs[base + r1], out.ctypes.data, None, 0, n_sb)
    return zlib.crc32(out.tobytes())


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = importlib.import_module("svt-vp9_b200.frames")
    presets = importlib.import_module("svt-vp9_b200.presets")
    w, h, n_pictures = 192, 128, 13
    clip = H.synth.make_clip(w, h, n_pictures, seed=5, k=3, g=4, a=5, b=3, m=32, noise=2)
    planes = [frames.pa_planes(f[0]) for f in clip]
    structs = [frames.plane_structs(p) for p in planes]
    P_of = lambda tl, poc: presets.me_params_1080p_m8(w, h, tl + 2, 2, poc)
    # mini-GOPs of 4 pictures (2 levels) so that a 13-picture clip gives each of the 2 ranks work
    runs = [(1 + g * 4, min(4, n_pictures - 1 - g * 4)) for g in range(rank, (n_pictures - 1 + 3) // 4, world)]
    mine = {}
    for first, count in runs:
        for d in range(first, first + count):
            mine[d] = _me_digest(planes, structs, d, P_of)
    ms, total = sharding.agree_step(10.0 + rank, len(mine))
    q.put((rank, mine, ms, total))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_cover_the_clip_and_match_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    merged = {}
    for rank, mine, ms, total in got:
        assert ms == 11.0 and total == 12            # max over ranks; all 12 inter pictures processed exactly once
        assert not (set(mine) & set(merged))
        merged.update(mine)
    assert sorted(merged) == list(range(1, 13))
    # the same pictures in one process give the same results: no rank needed anything from the other
    frames = importlib.import_module("svt-vp9_b200.frames")
    presets = importlib.import_module("svt-vp9_b200.presets")
    clip = H.synth.make_clip(192, 128, 13, seed=5, k=3, g=4, a=5, b=3, m=32, noise=2)
    planes = [frames.pa_planes(f[0]) for f in clip]
    structs = [frames.plane_structs(p) for p in planes]
    P_of = lambda tl, poc: presets.me_params_1080p_m8(192, 128, tl + 2, 2, poc)
    for d in (1, 4, 6, 12):
        assert _me_digest(planes, structs, d, P_of) == merged[d]
