"""N>1 host-side path on CPU: world_size-2 gloo processes shard a frame stream, produce per-frame
payloads and gather them back in frame order (no collective on the data path)."""
import os
import sys

import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_frames, chunk, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from yolo_puncture_b200.sharding import gather_in_frame_order, shard_indices
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_indices(n_frames, rank, world, chunk)
    payload = [("frame", i, rank) for i in mine]
    full = gather_in_frame_order(payload, n_frames, rank, world, chunk)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, full))


def test_shard_indices_cover_stream_once():
    from yolo_puncture_b200.sharding import shard_counts, shard_indices
    for n, world, chunk in [(10, 2, 4), (64, 8, 8), (7, 4, 2), (1, 2, 1)]:
        all_idx = sorted(i for r in range(world) for i in shard_indices(n, r, world, chunk))
        assert all_idx == list(range(n))
        assert sum(shard_counts(n, world, chunk)) == n
    assert shard_indices(10, 0, 2, 4) == [0, 1, 2, 3, 8, 9] and shard_indices(10, 1, 2, 4) == [4, 5, 6, 7]


def test_two_rank_gloo_gather_in_frame_order():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_frames, chunk, world, port = 11, 3, 2, 29731
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, chunk, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, full in outs:
        assert [f[1] for f in full] == list(range(n_frames))
        assert [f[2] for f in full] == [(i // chunk) % world for i in range(n_frames)]
