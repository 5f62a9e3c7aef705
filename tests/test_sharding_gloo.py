"""N > 1 host-side logic on CPU with the gloo backend, world_size 2: window
sharding covers every item exactly once, the weight broadcast makes replicas identical, the
max-over-ranks reduction used for timing works. No GPU needed."""
import contextlib
import io
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neuralrgbd_b200 import sharding


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from neuralrgbd_b200 import camera, synth
    from neuralrgbd_b200.models.KVNET import KVNET
    cam = camera.make_cam_intrinsics(585., 585., 320., 240., [64, 64])
    with contextlib.redirect_stdout(io.StringIO()):
        m = KVNET(64, cam, synth.d_candidates(8), 10., 64, None, t_win_r=2)
    torch.manual_seed(100 + rank)                       # replicas start different
    for p in m.parameters():
        p.data.normal_()
    n = sharding.broadcast_module(m, src=0)
    digest = float(sum(p.double().sum() for p in m.parameters()))
    mine = sharding.shard_indices(11, rank, world)
    tmax = sharding.max_over_ranks(1.0 + rank)
    q.put((rank, n, digest, mine, tmax))
    dist.destroy_process_group()


def test_broadcast_sharding_and_max_world2():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, n0, d0, m0, t0), (r1, n1, d1, m1, t1) = res
    assert n0 == n1 and 4_000_000 < n0 < 6_000_000         # every parameter broadcast exactly once (D=8 net: 4.62 M)
    assert d0 == d1                                        # identical replicas after the broadcast
    assert sorted(m0 + m1) == list(range(11)) and not set(m0) & set(m1)
    assert t0 == t1 == 2.0


def test_chunking_restarts_cover_trajectory():
    chunks = sharding.chunk_trajectory(n_frames=30, t_win_r=2, n_chunks=8)
    refs = [i for c in chunks for i in range(c['ref_begin'], c['ref_end'])]
    assert refs == list(range(2, 28))
    for c in chunks:
        assert c['frame_begin'] >= 0 and c['frame_end'] <= 30
        assert c['frame_begin'] == c['ref_begin'] - 2 and c['frame_end'] == c['ref_end'] + 2
    assert sharding.chunk_trajectory(4, 2, 3) == []
