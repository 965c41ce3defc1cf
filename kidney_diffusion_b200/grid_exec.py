"""Transports of the patch-grid sampler: how a border strip (or a previous-stage patch) produced on one rank reaches the rank
that consumes it.  Every message has exactly one producer and one consumer, both known from the
static plan (grid_plan.py), and is identified by a key (stage, patch index, kind).

* PeerMailbox   -- the B200 path.  Each rank owns a mailbox in its HBM with one slot + one flag word per incoming message,
                   exported through CUDA IPC and mapped by every other rank of the node.  `post` = a copy kernel storing the
                   strip straight into the consumer's slot over NVLink, followed by a release-store of the flag; `fetch` =
                   a one-thread kernel on the consumer's stream that spins on the flag (acquire) and returns a view of the
                   slot.  One-sided and asynchronous: the host never waits, ranks never rendezvous, nothing is matched.
* TaggedTransport -- torch.distributed isend / recv with one tag per message (gloo: the CPU tests of the multi-rank logic).
* NCCL fallback -- grid.py keeps the round-synchronous `batch_isend_irecv` exchange of round 1 for boxes where IPC mapping
                   is unavailable (KD_GRID_TRANSPORT=nccl forces it).
"""
from __future__ import annotations

import ctypes
import os

import torch

KINDS = {"above": 0, "side": 1, "corner": 2, "lowres": 3, "done": 4}
ALIGN = 256


def strip_shape(kind, S, ov):
    return {"above": (3, ov, S), "side": (3, S, ov), "corner": (3, ov, ov), "lowres": (3, S, S)}[kind]


class TaggedTransport:
    """Point-to-point messages over torch.distributed with a unique tag per message (gloo)."""

    name = "torch.distributed isend/recv (tagged)"

    def __init__(self, dist, rank, n_patches):
        self.dist, self.rank, self.n = dist, rank, max(1, n_patches)
        self.pending = []
        self.bytes_sent = 0

    def _tag(self, key):
        stage, k, kind = key
        return (stage * 8 + KINDS[kind]) * self.n + k

    def post(self, dst, key, view):
        t = view.contiguous()
        self.bytes_sent += t.numel() * 4
        self.pending.append((self.dist.isend(t, dst, tag=self._tag(key)), t))

    def fetch(self, src, key, shape, device):
        buf = torch.empty(shape, device=device, dtype=torch.float32)
        self.dist.recv(buf, src, tag=self._tag(key))
        return buf

    def finish(self):
        for req, _ in self.pending:
            req.wait()
        self.pending = []

    def close(self):
        self.finish()


class _RawCuda:
    """Wraps a raw device allocation for torch.as_tensor (__cuda_array_interface__, zero copy)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = dict(shape=(nbytes // 4,), typestr="<f4", data=(ptr, False), version=2, strides=None)


class PeerMailbox:
    """incoming: for every rank, the ordered list of (key, shape) it will receive -- identical on all ranks."""

    name = "CUDA-IPC peer mailbox (one-sided NVLink stores + device-side flags)"

    def __init__(self, dist, rank, world, device, incoming, timeout_s=None):
        from . import _lib, ops

        self.dist, self.rank, self.world, self.device = dist, rank, world, device
        self.ops, self.lib = ops, _lib.load()
        self.timeout_s = float(os.environ.get("KD_GRID_TIMEOUT_S", "900")) if timeout_s is None else timeout_s
        self.bytes_sent = 0
        # layout of every rank's mailbox: [flags: one u32 per message, padded][slots...]
        self.layout = []
        for r in range(world):
            msgs = incoming.get(r, [])
            off = -(-4 * max(1, len(msgs)) // ALIGN) * ALIGN
            slots = {}
            for idx, (key, shape) in enumerate(msgs):
                nbytes = 4 * int(torch.Size(shape).numel())
                slots[key] = (idx, off, shape)
                off += -(-nbytes // ALIGN) * ALIGN
            self.layout.append((slots, max(off, ALIGN)))
        self.base = [None] * world
        self.own = None
        self.status = None
        self.epoch = 1  # value a raised flag carries; a cached mailbox is reused with the next epoch instead of being re-zeroed

    def allocate(self):
        """Phase 1 (local): allocate + export this rank's mailbox."""
        from . import _lib

        own = ctypes.c_void_p()
        _lib.check(self.lib.kd_peer_alloc(self.layout[self.rank][1], ctypes.byref(own)), "kd_peer_alloc")
        self.base[self.rank] = own.value
        handle = (ctypes.c_uint8 * 64)()
        _lib.check(self.lib.kd_peer_export(own, handle), "kd_peer_export")
        self.handle = bytes(handle)
        self.own = torch.as_tensor(_RawCuda(self.base[self.rank], self.layout[self.rank][1]), device=self.device)
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)

    def exchange_handles(self):
        """Phase 2a (collective, cannot fail locally): everyone learns everyone's IPC handle (64 opaque bytes; works on any
        backend -- NCCL in production, gloo when the tests put two ranks on one GPU)."""
        self.everyone = [None] * self.world
        self.dist.all_gather_object(self.everyone, self.handle)

    def map_peers(self):
        """Phase 2b (local): map the other ranks' mailboxes into this process."""
        from . import _lib

        for r in range(self.world):
            if r == self.rank:
                continue
            h = (ctypes.c_uint8 * 64)(*self.everyone[r])
            p = ctypes.c_void_p()
            _lib.check(self.lib.kd_peer_open(h, ctypes.byref(p)), f"kd_peer_open(rank {r})")
            self.base[r] = p.value

    def post(self, dst, key, view):
        idx, off, shape = self.layout[dst][0][key]
        assert tuple(view.shape) == tuple(shape), (key, tuple(view.shape), shape)
        self.bytes_sent += view.numel() * 4
        self.ops.strip_push(view, view.stride(0), view.stride(1), self.base[dst] + off, self.base[dst] + 4 * idx, self.epoch)

    def fetch(self, src, key, shape, device):
        idx, off, shp = self.layout[self.rank][0][key]
        self.ops.flag_wait(self.base[self.rank] + 4 * idx, self.epoch, self.timeout_s, self.status)
        n = int(torch.Size(shp).numel())
        return self.own[off // 4: off // 4 + n].view(shp)

    def finish(self):
        torch.cuda.synchronize(self.device)
        if int(self.status.item()) != 0:
            raise RuntimeError(f"rank {self.rank}: {int(self.status.item())} border strips did not arrive within {self.timeout_s:.0f} s "
                               "(peer rank failed or the plan's order was violated)")

    def recycle(self):
        """Collective, after finish(): the next run may reuse every slot once ALL ranks are done reading this run's."""
        self.dist.barrier()
        self.epoch += 1
        self.bytes_sent = 0

    def close(self):
        """Collective: nobody may free a mailbox another rank still has mapped (or is still writing into)."""
        from . import _lib

        self.own = None
        torch.cuda.synchronize(self.device)
        self.dist.barrier()
        for r in range(self.world):
            if r != self.rank and self.base[r] is not None:
                _lib.check(self.lib.kd_peer_close(ctypes.c_void_p(self.base[r])), "kd_peer_close")
                self.base[r] = None
        self.dist.barrier()
        if self.base[self.rank] is not None:
            _lib.check(self.lib.kd_peer_free(ctypes.c_void_p(self.base[self.rank])), "kd_peer_free")
            self.base[self.rank] = None


def try_peer_mailbox(dist, rank, world, device, incoming):
    """PeerMailbox when every rank of the group can set it up, else None (the decision is collective)."""
    if os.environ.get("KD_GRID_TRANSPORT", "auto").lower() == "nccl":
        return None
    def agree(ok):
        votes = [None] * world
        dist.all_gather_object(votes, bool(ok))
        return all(votes)

    def attempt(fn):
        try:
            fn()
            return True
        except Exception as e:  # noqa: BLE001 -- any failure (IPC not permitted, no peer access, ...) selects the fallback on ALL ranks
            print(f"[kidney_b200.grid] rank {rank}: CUDA-IPC mailbox unavailable ({e}); falling back to NCCL rounds", flush=True)
            return False

    box = PeerMailbox(dist, rank, world, device, incoming)
    ok = agree(attempt(box.allocate))
    if ok:
        box.exchange_handles()
        ok = agree(attempt(box.map_peers))
    if ok:
        return box
    try:
        box.close()
    except Exception:  # noqa: BLE001
        pass
    return None


_BOX_CACHE = {}


def cached_peer_mailbox(dist, rank, world, device, incoming, max_entries=4, cache=None):
    """Plan mailboxes are reused between runs with the same message layout (allocation + IPC mapping cost ~0.1 s); all ranks
    call this with identical arguments in the same order, so cache hits, misses and evictions are collective."""
    cache = _BOX_CACHE if cache is None else cache
    key = (world, str(device), repr(sorted((r, tuple(m)) for r, m in incoming.items())))
    box = cache.get(key)
    if box is not None:
        return box
    while len(cache) >= max_entries:
        cache.pop(next(iter(cache))).close()
    box = try_peer_mailbox(dist, rank, world, device, incoming)
    if box is not None:
        cache[key] = box
    return box


_CANVAS_CACHE = {}  # the stitch canvases (3.2 GB per rank at 16 384^2) are kept for the next image of the same size
