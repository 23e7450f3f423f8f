"""CPU, world_size 2, gloo: the N>1 path (contiguous frame ranges, absolute first_frame, ordered gather).
The per-frame work is a stand-in that encodes the absolute frame index, so ordering mistakes are visible."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from color_modem_b200.shard import frame_range, process_sharded


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total, results):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        seen = {}

        def work(first_frame, n):
            seen['range'] = (first_frame, n)
            frames = torch.arange(first_frame, first_frame + n, dtype=torch.int64)
            return torch.stack([frames, frames * frames], dim=1)        # [n, 2], depends on the ABSOLUTE index

        out = process_sharded(total, work, gather=True)
        assert seen['range'] == (frame_range(total, rank, world)[0],
                                 frame_range(total, rank, world)[1] - frame_range(total, rank, world)[0])
        if rank == 0:
            expect = torch.arange(total, dtype=torch.int64)
            assert out.shape == (total, 2)
            assert torch.equal(out[:, 0], expect) and torch.equal(out[:, 1], expect * expect)
            results.put('ok')
        else:
            assert out is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_frame_sharding_and_gather():
    ctx = mp.get_context('spawn')
    results = ctx.Queue()
    for total in (7, 600):
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, total, results)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        assert results.get(timeout=10) == 'ok'
