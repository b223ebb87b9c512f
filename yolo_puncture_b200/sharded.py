"""Frame-sharded detector across the GPUs of one box, with the point-to-point mask hand-off (BASELINE
configs C4 / C5).

One process per GPU (`torchrun`): every rank owns an engine replica on GPU `LOCAL_RANK` and the frames
`sharding.shard_indices(n, rank, world, chunk)` of each stream chunk.  Nothing is exchanged on the data path: the ranks run
`YOLO.predict()` on their own frames, the host gathers the small per-frame payloads (counts + boxes) into frame order
(`sharding.gather_in_frame_order`), and the masks stay on the GPU that produced them -

- unless a single consumer needs them: the reference's tracker (DEVA, `yolo_seg/yolo_with_deva.py:133-159`) is sequential and
  lives on ONE GPU.  `MaskMailbox` is a buffer in that GPU's memory, exported to the other ranks with CUDA IPC; every rank
  builds the int64 index masks of its frames (`handoff.index_masks`, reference `auto_segment`, yolo_with_deva.py:54-88) and
  copies them into the mailbox slots of those frames with `cudaMemcpyAsync` between peer devices (NVLink / NVSwitch),
  no NCCL and no host bounce.  The ordered host gather that follows doubles as the "slots are filled" notification.
"""

import ctypes as C
import os

import torch

from ._lib import check, lib
from .handoff import index_masks
from .sharding import gather_in_frame_order, shard_indices, summarize_results


class _DevBuffer:
    """A raw device allocation seen by torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3,
                                         "strides": None}


class MaskMailbox:
    """(slots, H, W) int64 index masks in the consumer GPU's memory.

    consumer rank:  mb = MaskMailbox.create(device, slots, H, W); handle = mb.handle   (64 bytes, send to the producers)
    producer rank:  mb = MaskMailbox.open(device, handle, slots, H, W); mb.push(slot, index_mask)
    """

    def __init__(self, device, ptr, slots, H, W, owner, handle=None):
        self.device, self.ptr, self.slots, self.H, self.W, self.owner, self.handle = device, ptr, slots, H, W, owner, handle
        self.slot_bytes = H * W * 8
        self._tensor = None

    @classmethod
    def create(cls, device, slots, H, W):
        ptr, handle = C.c_void_p(), (C.c_ubyte * 64)()
        check(lib().ypb_mailbox_create(int(device), slots * H * W * 8, C.byref(ptr), handle))
        return cls(int(device), ptr.value, slots, H, W, True, bytes(handle))

    @classmethod
    def open(cls, device, handle, slots, H, W):
        ptr = C.c_void_p()
        buf = (C.c_ubyte * 64).from_buffer_copy(bytes(handle))
        check(lib().ypb_mailbox_open(int(device), buf, C.byref(ptr)))
        return cls(int(device), ptr.value, slots, H, W, False, bytes(handle))

    def tensor(self):
        """The mailbox as a (slots, H, W) int64 tensor - on the consumer rank this is what the tracker reads."""
        if self._tensor is None:
            with torch.cuda.device(self.device):
                self._tensor = torch.as_tensor(_DevBuffer(self.ptr, (self.slots, self.H, self.W), "<i8"), device=f"cuda:{self.device}")
        return self._tensor

    def push(self, slot, index_mask, stream=None):
        """Copy one (H, W) int64 device tensor (this rank's GPU) into mailbox slot `slot` (the consumer's GPU)."""
        if index_mask.dtype != torch.int64 or tuple(index_mask.shape) != (self.H, self.W) or not index_mask.is_cuda:
            raise ValueError("push needs an (H, W) int64 device tensor")
        if not 0 <= slot < self.slots:
            raise IndexError(slot)
        src = index_mask.contiguous()
        st = (stream or torch.cuda.current_stream(src.device)).cuda_stream
        check(lib().ypb_peer_copy(C.c_void_p(st), C.c_void_p(self.ptr + slot * self.slot_bytes), C.c_void_p(src.data_ptr()),
                                  self.slot_bytes))

    def push_many(self, first_slot, index_maps, stream=None):
        """(n, H, W) contiguous int64 device tensor -> slots [first_slot, first_slot + n) in one copy."""
        n = int(index_maps.shape[0])
        if index_maps.dtype != torch.int64 or tuple(index_maps.shape[1:]) != (self.H, self.W) or not index_maps.is_contiguous():
            raise ValueError("push_many needs a contiguous (n, H, W) int64 device tensor")
        if first_slot < 0 or first_slot + n > self.slots:
            raise IndexError((first_slot, n))
        st = (stream or torch.cuda.current_stream(index_maps.device)).cuda_stream
        check(lib().ypb_peer_copy(C.c_void_p(st), C.c_void_p(self.ptr + first_slot * self.slot_bytes),
                                  C.c_void_p(index_maps.data_ptr()), n * self.slot_bytes))

    def close(self):
        if self.ptr:
            self._tensor = None
            if self.owner:
                check(lib().ypb_mailbox_destroy(self.device, C.c_void_p(self.ptr)))
            else:
                check(lib().ypb_mailbox_close(self.device, C.c_void_p(self.ptr)))
            self.ptr = 0


class ShardedPredictor:
    """One rank of a frame-sharded `predict` over a stream chunk of `n_frames` frames (chunk = frames per rank and round).

    predict(local_frames, n_frames, ...) -> (ordered, local_results): `ordered[i]` = (n_i, boxes_i (n_i, 6) numpy) for every
    frame i of the chunk in global order, identical on all ranks; `local_results` = this rank's `Results` (masks on its GPU).
    With a mailbox, `handoff=True` also builds the index masks of the local frames and pushes them to the consumer GPU."""

    def __init__(self, yolo, rank=None, world=None, chunk=None, group=None):
        self.yolo = yolo
        self.rank = int(os.environ.get("RANK", "0")) if rank is None else rank
        self.world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else world
        self.chunk = chunk
        self.group = group
        self.mailbox = None

    def frames_of(self, n_frames):
        return shard_indices(n_frames, self.rank, self.world, self.chunk or max(1, n_frames // self.world))

    def attach_mailbox(self, H, W, slots, consumer_rank=0):
        """Collective over the group: the consumer creates the mailbox, everybody else maps it (CUDA IPC)."""
        import torch.distributed as dist
        dev = self.yolo.device.index if self.yolo.device is not None else torch.cuda.current_device()
        if self.world == 1:
            self.mailbox = MaskMailbox.create(dev, slots, H, W)
            return self.mailbox
        box = [None]
        if self.rank == consumer_rank:
            self.mailbox = MaskMailbox.create(dev, slots, H, W)
            box[0] = self.mailbox.handle
        dist.broadcast_object_list(box, src=consumer_rank, group=self.group)
        if self.rank != consumer_rank:
            self.mailbox = MaskMailbox.open(dev, box[0], slots, H, W)
        return self.mailbox

    def predict(self, local_frames, n_frames, handoff=False, min_area=100, **predict_kw):
        mine = self.frames_of(n_frames)
        if len(mine) != len(local_frames):
            raise ValueError(f"rank {self.rank}: {len(local_frames)} frames for {len(mine)} owned indices")
        res = self.yolo.predict(local_frames, **predict_kw) if local_frames else []
        payload = summarize_results(res)
        if handoff and res:
            pairs = index_masks(res, suppress_small_mask=True, min_area=min_area)
            payload = [p + (info,) for p, (_, info) in zip(payload, pairs)]
            if self.mailbox is not None:
                maps = [m for m, _ in pairs]
                # runs of consecutive frame indices are views of one (n, H, W) buffer: one peer copy per run
                i = 0
                while i < len(mine):
                    j = i + 1
                    while j < len(mine) and mine[j] == mine[j - 1] + 1 and \
                            maps[j].data_ptr() == maps[j - 1].data_ptr() + maps[j - 1].numel() * 8:
                        j += 1
                    if j - i > 1:
                        run = torch.as_strided(maps[i], (j - i, self.mailbox.H, self.mailbox.W),
                                               (self.mailbox.H * self.mailbox.W, self.mailbox.W, 1))
                        self.mailbox.push_many(mine[i] % self.mailbox.slots, run)
                    else:
                        self.mailbox.push(mine[i] % self.mailbox.slots, maps[i])
                    i = j
                torch.cuda.current_stream().synchronize()  # the copies have landed before the gather announces them
        chunk = self.chunk or max(1, n_frames // self.world)
        if self.world > 1 and self.group is None and self._nccl_ok():
            ordered = self._gather_rows_nccl(payload, n_frames, chunk, handoff and bool(res))
        else:
            ordered = gather_in_frame_order(payload, n_frames, self.rank, self.world, chunk, group=self.group)
        return ordered, res

    # ------------------------------------------------------------------ fixed-size gather of the per-frame rows
    @staticmethod
    def _nccl_ok():
        import torch.distributed as dist
        return dist.is_initialized() and dist.get_backend() == "nccl" and torch.cuda.is_available()

    def _gather_rows_nccl(self, payload, n_frames, chunk, with_ids):
        """The ordered gather of (count, boxes[, kept ids]) per frame as ONE fixed-size all_gather of a small tensor (a few
        hundred KB per rank) instead of pickled objects: results only - counts, boxes, ids - never activations or masks.
        Every rank ends up with the same ordered list `gather_in_frame_order` would return."""
        import torch.distributed as dist
        from .engine import MAX_DET
        from .sharding import shard_counts
        dev = self.yolo.device
        width = 7 if with_ids else 6
        cap = max(shard_counts(n_frames, self.world, chunk))
        key = (cap, width)
        if getattr(self, "_gkey", None) != key:
            self._gkey = key
            self._ghost = torch.zeros((cap, 1 + MAX_DET * width), dtype=torch.float32).pin_memory()
            self._gdev = torch.empty((self.world, cap, 1 + MAX_DET * width), dtype=torch.float32, device=dev)
        host = self._ghost
        hn = host.numpy()
        hn[:, 0] = 0
        for j, item in enumerate(payload):
            n, boxes = item[0], item[1]
            hn[j, 0] = n
            if n:
                rows = hn[j, 1:1 + n * width].reshape(n, width)
                rows[:, :6] = boxes
                if with_ids:
                    rows[:, 6] = 0
                    if item[2]:  # kept detections: id 1, 2, ... at their row of the frame's Results
                        rows[[d["index"] for d in item[2]], 6] = [d["id"] for d in item[2]]
        with torch.cuda.device(dev):
            mine = host.to(dev, non_blocking=True)
            dist.all_gather_into_tensor(self._gdev.view(self.world * cap, -1), mine)
            allh = self._gdev.cpu().numpy()
        where = [None] * n_frames
        for r in range(self.world):
            for j, idx in enumerate(shard_indices(n_frames, r, self.world, chunk)):
                where[idx] = (r, j)
        return _OrderedRows(allh, where, width, with_ids)


class _OrderedRows:
    """The gathered per-frame rows in global frame order, unpacked on access: `rows[i]` -> (n_i, boxes (n_i, 6)[, info]).
    (Unpacking every frame of every rank eagerly cost more host time at 8 ranks than the detector itself.)"""

    def __init__(self, table, where, width, with_ids):
        self._t, self._where, self._w, self._ids = table, where, width, with_ids

    def __len__(self):
        return len(self._where)

    def _one(self, i):
        r, j = self._where[i]
        n = int(self._t[r, j, 0])
        rows = self._t[r, j, 1:1 + n * self._w].reshape(n, self._w)
        if not self._ids:
            return (n, rows.copy())
        kept = rows[:, 6] > 0
        info = [{"id": int(q[6]), "score": float(q[4]), "category_id": int(q[5]), "index": int(k)}
                for k, q in zip(kept.nonzero()[0], rows[kept])]
        return (n, rows[:, :6].copy(), info)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._one(k) for k in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        return self._one(i)

    def __iter__(self):
        return (self._one(i) for i in range(len(self)))

    def counts(self):
        """Detections per frame, in frame order (one vectorised read)."""
        import numpy as np
        return np.array([int(self._t[r, j, 0]) for r, j in self._where])
