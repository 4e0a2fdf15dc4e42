"""N-axis sharding: who exchanges what, and how the shards reach each other's mailbox.

Every shard owns a contiguous range of items and a replica of the small global state (dirichlets, tables,
pi_hat).  Per acquisition step exactly two exchanges cross GPUs, both INSIDE the fused step kernels
(``csrc/xchg.cuh``, ``csrc/step.cu``) as peer-memory stores over NVLink -- no NCCL call sits on the step path:

  1. arg-max     every shard stores {its best record, p_h(idx) of its best candidates} into every peer's mailbox;
                 all shards merge to the same global record and so know the chosen item AND its hard-prediction row
  2. marginals   every shard stores its sum_n pi_hat_xi[n, :] (int64 fixed point: exact, so pi_hat is
                 bit-identical for every shard count); all shards add them up in rank order

The API path (host picks / host labels) adds a third, the owner's p_h(idx) for a host-chosen idx, and ships the tie
lists the same way.  Construction adds one SUM all-reduce of the (H, C, C) soft-confusion sums (coda.py:42) -- bulk
data, so that one goes through NCCL (``TorchComm``) or, with one process driving all GPUs, through peer copies.

Two ways to form a group:
  ``ProcessGroup``    one process per GPU (torchrun): mailboxes are mapped with CUDA IPC handles
  ``InProcessGroup``  one Python process drives all shards (``main.py`` unchanged): plain peer access; the shards may
                      even share one GPU (each on its own stream), which is how the 1-GPU test tier covers sharding
"""
from __future__ import annotations

import ctypes as ct

import torch

from . import _native as nat


class LocalComm:
    world = 1
    rank = 0

    def allreduce_sum_(self, t):
        return t

    def allreduce_min_(self, t):
        return t

    def allgather(self, t):
        return t.unsqueeze(0)

    def barrier(self):
        pass


class TorchComm:
    """torch.distributed process group (NCCL on GPUs, gloo in the CPU tests): construction-time reductions."""

    def __init__(self, group=None):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def allreduce_sum_(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def allreduce_min_(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN, group=self.group)
        return t

    def allgather(self, t):
        out = torch.empty((self.world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out.view(-1), t.contiguous().view(-1), group=self.group)
        return out

    def barrier(self):
        self.dist.barrier(group=self.group)


def default_comm():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return TorchComm()
    return LocalComm()


# ---------------------------------------------------------------------------------------
# mailboxes
# ---------------------------------------------------------------------------------------
class Mailbox:
    """One shard's mailbox (device memory from the library's own cudaMalloc, so it can be IPC-exported)."""

    def __init__(self, device, world, H, C, rep_words):
        self.lib = nat.load()
        self.device = torch.device(device)
        self.dims = (int(world), int(H), int(C), int(rep_words))
        self.bytes = int(self.lib.coda_b200_xchg_box_bytes(*self.dims))
        ptr = ct.c_void_p()
        with torch.cuda.device(self.device):
            nat.check(self.lib.coda_b200_xchg_alloc(self.bytes, ct.byref(ptr)), "xchg_alloc")
        self.ptr = int(ptr.value)
        self.epoch = torch.zeros(8, dtype=torch.int64, device=self.device)     # epochs [0..4), wait ns [4..8)
        torch.cuda.synchronize(self.device)     # the zero fill ran on the current stream; shards may use their own
        self.opened = []

    def export(self) -> bytes:
        buf = ct.create_string_buffer(64)
        with torch.cuda.device(self.device):
            nat.check(self.lib.coda_b200_ipc_export(self.ptr, buf), "ipc_export")
        return buf.raw

    def open_peer(self, handle: bytes) -> int:
        out = ct.c_void_p()
        with torch.cuda.device(self.device):
            nat.check(self.lib.coda_b200_ipc_open(ct.create_string_buffer(handle, 64), ct.byref(out)), "ipc_open")
        self.opened.append(int(out.value))
        return int(out.value)

    def view(self, rank, peer_ptrs) -> nat.XchgStruct:
        world, H, Cc, rep_words = self.dims
        x = nat.XchgStruct()
        x.world, x.rank = world, rank
        for r, ptr in enumerate(peer_ptrs):
            x.box[r] = ptr
        x.epoch = self.epoch.data_ptr()
        x.H, x.C, x.rep_words = H, Cc, rep_words
        return x

    def close(self):
        lib = self.lib
        try:
            with torch.cuda.device(self.device):
                for q in self.opened:
                    lib.coda_b200_ipc_close(q)
                self.opened = []
                if self.ptr:
                    lib.coda_b200_xchg_free(self.ptr)
                    self.ptr = 0
        except Exception:
            pass


class SoloGroup:
    """world == 1: no exchange."""
    world = 1

    def __init__(self):
        self.comm = LocalComm()

    def rank_of(self, engine):
        return 0

    def attach(self, engines):
        for e in engines:
            e.xchg = None

    def allreduce_sum_(self, tensors):
        return tensors

    def barrier(self):
        pass


class ProcessGroup:
    """One process per GPU (torchrun).  ``comm`` does the construction-time all-reduce; mailboxes go through CUDA IPC."""

    def __init__(self, comm: TorchComm):
        self.comm = comm
        self.world, self.rank = comm.world, comm.rank
        self.box = None

    def rank_of(self, engine):
        return self.rank

    def attach(self, engines):
        (eng,) = engines
        self.box = Mailbox(eng.dev, self.world, eng.H, eng.C, eng.rep_words)
        mine = torch.frombuffer(bytearray(self.box.export()), dtype=torch.uint8).to(eng.dev)
        allh = self.comm.allgather(mine).cpu()                               # (world, 64)
        ptrs = []
        for r in range(self.world):
            ptrs.append(self.box.ptr if r == self.rank else self.box.open_peer(bytes(allh[r].numpy().tobytes())))
        eng.xchg = self.box.view(self.rank, ptrs)
        eng._mailbox = self.box
        self.comm.barrier()                                                  # every mailbox exists and is zeroed

    def allreduce_sum_(self, tensors):
        for t in tensors:
            self.comm.allreduce_sum_(t)
        return tensors

    def barrier(self):
        self.comm.barrier()


class InProcessGroup:
    """All shards driven by this process: engines[r] is rank r, on any mix of devices (peer access is enabled
    between distinct devices; shards on the same device just use different streams)."""

    def __init__(self, world):
        self.world = int(world)
        self.comm = LocalComm()
        self.boxes = []
        self._engines = []

    def rank_of(self, engine):
        return self._engines.index(engine)

    def attach(self, engines):
        assert len(engines) == self.world
        self._engines = list(engines)
        lib = nat.load()
        devs = sorted({e.dev.index for e in engines})
        for a in devs:
            for b in devs:
                if a != b:
                    with torch.cuda.device(a):
                        nat.check(lib.coda_b200_peer_enable(b), "peer_enable")
        self.boxes = [Mailbox(e.dev, self.world, e.H, e.C, e.rep_words) for e in engines]
        ptrs = [b.ptr for b in self.boxes]
        for r, e in enumerate(engines):
            e.xchg = self.boxes[r].view(r, ptrs)
            e._mailbox = self.boxes[r]

    def allreduce_sum_(self, tensors):
        """tensors[r] lives on engine r's device: every one ends up holding the sum (construction only)."""
        for t, e in zip(tensors, self._engines):
            e.sync()
        tot = tensors[0].clone()
        for t in tensors[1:]:
            tot += t.to(tot.device)
        for t in tensors:
            t.copy_(tot.to(t.device))
        for d in {t.device for t in tensors}:
            torch.cuda.synchronize(d)
        return tensors

    def barrier(self):
        pass


# ---------------------------------------------------------------------------------------
# host-side mirrors of the device merge rules (used by the CPU/gloo tests and the slow paths)
# ---------------------------------------------------------------------------------------
IDX_NONE = (1 << 63) - 1


def merge_records(recs):
    """recs: iterable of (vA, iA, cntA, vB, iB).  Max value, lowest index on equal values,
    counts summed -- the rule of best2_merge (csrc/common.cuh) without the runner-up."""
    va, ia, ca, vb, ib = float("-inf"), IDX_NONE, 0, float("-inf"), IDX_NONE
    for (a, i, c, b, j) in recs:
        if a > va or (a == va and i < ia):
            va, ia = a, i
        if b > vb or (b == vb and j < ib):
            vb, ib = b, j
        ca += c
    return va, ia, ca, vb, ib


def merge_best2(items):
    """Host mirror of best2_merge: items = iterable of (v, i, v2); returns the merged (v, i, v2) where v2 is the
    best value among all OTHER items (the runner-up the isclose tie test needs)."""
    v, i, v2 = float("-inf"), IDX_NONE, float("-inf")
    for (ov, oi, ov2) in items:
        if oi == IDX_NONE:
            continue
        if i == IDX_NONE:
            v, i, v2 = ov, oi, ov2
        elif ov > v or (ov == v and oi < i):
            v2 = max(v2, ov2, v)
            v, i = ov, oi
        else:
            v2 = max(v2, ov2, ov)
    return v, i, v2


def choose_among_ties(tie_idx, rng):
    """coda.py:308: random.choice over the tied candidates in ascending index order.  ``random.choice``
    consumes exactly one ``_randbelow(len)``, so choosing a position is RNG-equivalent."""
    order = sorted(int(i) for i in tie_idx)
    return order[rng.choice(range(len(order)))]
