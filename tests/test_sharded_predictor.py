"""Frame-sharded predictor: host-side logic on CPU with world_size-2 gloo processes and a stand-in engine;
on a box with >= 2 GPUs, real replicas, the CUDA-IPC mask mailbox and the one-process-two-devices case."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _FakeBoxes:
    def __init__(self, data):
        self.data = data

    def cpu(self):
        return self

    def numpy(self):
        return self

    def __len__(self):
        return len(self.data)


class _FakeResult:
    def __init__(self, idx):
        self.boxes = _FakeBoxes(np.full((idx % 3, 6), float(idx), np.float32))
        self.masks = None


class _FakeYolo:
    device = None

    def predict(self, frames, **kw):
        return [_FakeResult(int(f)) for f in frames]


def _cpu_worker(rank, world, port, n_frames, chunk, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from yolo_puncture_b200.sharded import ShardedPredictor
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sp = ShardedPredictor(_FakeYolo(), rank, world, chunk)
    mine = sp.frames_of(n_frames)
    ordered, local = sp.predict(mine, n_frames)  # the stand-in "frames" are their own indices
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, mine, [(n, b.tolist()) for n, b in ordered], len(local)))


def test_sharded_predict_two_ranks_gloo_returns_global_frame_order():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_frames, chunk, world, port = 13, 4, 2, 29741
    procs = [ctx.Process(target=_cpu_worker, args=(r, world, port, n_frames, chunk, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert outs[0][1] == [0, 1, 2, 3, 8, 9, 10, 11] and outs[1][1] == [4, 5, 6, 7, 12]
    for rank, mine, ordered, n_local in outs:
        assert n_local == len(mine) and len(ordered) == n_frames
        for i, (n, boxes) in enumerate(ordered):  # every rank holds every frame's payload, in frame order
            assert n == i % 3 and all(v == float(i) for row in boxes for v in row)


def _gpu_worker(rank, world, port, q, backend="gloo"):
    try:
        _gpu_worker_body(rank, world, port, q, backend)
    except Exception as e:  # never leave the parent waiting on the queue
        import traceback
        q.put((rank, False, [f"{type(e).__name__}: {e}", traceback.format_exc()[-600:]], 0))


def _gpu_worker_body(rank, world, port, q, backend):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from yolo_puncture_b200 import YOLO, index_masks, synth
    from yolo_puncture_b200.sharded import ShardedPredictor
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group(backend, rank=rank, world_size=world, **({"device_id": torch.device("cuda", rank)} if backend == "nccl" else {}))
    yolo = YOLO("yolov8n-seg", device=rank)
    n_frames, chunk = 6, 3
    sp = ShardedPredictor(yolo, rank, world, chunk)
    sp.attach_mailbox(640, 640, n_frames, consumer_rank=0)
    mine = sp.frames_of(n_frames)
    frames = [synth.synth_frame(i) for i in mine]
    ordered, local = sp.predict(frames, n_frames, handoff=True, conf=0.25, iou=0.7, retina_masks=True)
    # what this rank pushed, as checksums; rank 0 then compares them with what sits in ITS memory
    sums = {i: int(m.sum().item()) for i, (m, _) in zip(mine, index_masks(local, True, 100))}
    all_sums = [None] * world
    dist.all_gather_object(all_sums, sums)
    ok = True
    if rank == 0:
        box = sp.mailbox.tensor()
        torch.cuda.synchronize()
        for d in all_sums:
            for i, v in d.items():
                ok = ok and int(box[i].sum().item()) == v
    # the fixed-size NCCL gather (taken on an NCCL default group) must return what the generic object gather returns
    from yolo_puncture_b200.sharding import gather_in_frame_order, summarize_results
    ref = gather_in_frame_order(summarize_results(local), n_frames, rank, world, chunk)
    same = all(a[0] == b[0] and np.array_equal(a[1], b[1]) for a, b in zip(ordered, ref))
    ids_ok = all(len(o) == 3 and all(d["index"] < o[0] for d in o[2]) for o in ordered)
    dist.barrier()
    sp.mailbox.close()
    dist.destroy_process_group()
    q.put((rank, ok and same and ids_ok, [n for n, *_ in ordered], sum(sums.values())))


@pytest.mark.gpu
@pytest.mark.parametrize("backend,port", [("gloo", 29743), ("nccl", 29745)])
def test_mask_mailbox_peer_push_between_two_gpus(backend, port):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gpu_worker, args=(r, 2, port, q, backend)) for r in range(2)]
    for p in procs:
        p.start()
    outs = []
    try:
        for _ in range(2):
            outs.append(q.get(timeout=240))
    finally:
        for p in procs:
            p.join(20)
            if p.is_alive():  # a rank stuck in a collective after its peer failed: do not hold the box
                p.terminate()
    outs.sort(key=lambda o: o[0])
    assert len(outs) == 2, outs
    assert outs[0][1] and outs[1][1], outs
    assert outs[0][2] == outs[1][2] and len(outs[0][2]) == 6 and sum(outs[0][2]) > 0  # same ordered counts on both ranks
    assert outs[0][3] + outs[1][3] > 0


@pytest.mark.gpu
def test_one_process_two_devices():
    """per-device launch state - the same process drives cuda:0 and then cuda:1 and the caller's current device
    is left alone."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from yolo_puncture_b200 import YOLO, synth
    frames = synth.synth_frames(2)
    torch.cuda.set_device(0)
    y0 = YOLO("yolov8n-seg", device=0)
    a = y0.predict(frames, conf=0.25, retina_masks=True)
    y1 = YOLO("yolov8n-seg", device=1)
    assert torch.cuda.current_device() == 0
    b = y1.predict(frames, conf=0.25, retina_masks=True)
    assert torch.cuda.current_device() == 0
    y0.to("cuda:1")
    c = y0.predict(frames, conf=0.25, retina_masks=True)
    assert torch.cuda.current_device() == 0
    for ra, rb, rc in zip(a, b, c):
        assert rb.boxes.data.device.index == 1 and rc.boxes.data.device.index == 1
        assert torch.equal(ra.boxes.data.cpu(), rb.boxes.data.cpu()) and torch.equal(ra.boxes.data.cpu(), rc.boxes.data.cpu())
        if ra.masks is not None:
            assert torch.equal(ra.masks.raw.cpu(), rb.masks.raw.cpu())
