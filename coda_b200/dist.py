"""N-axis sharding: the three tiny exchanges of one acquisition step.

Every rank owns a contiguous range of items and a replica of the small global state
(dirichlets, tables, pi_hat).  Per step:
  1. arg-max      all-gather of one 5-word record per rank  -> merged on every rank
                  (+ all-gather of the tie lists, isclose is evaluated against the GLOBAL best)
  2. label        the owner of the chosen item shares p_h(idx) (H ints) -- as a SUM all-reduce
                  with zeros from non-owners, so no rank needs to know who the owner is
  3. marginals    SUM all-reduce of sum_n pi_hat_xi[n, :] in int64 fixed point: exact, so pi_hat is
                  bit-identical for every shard count
Construction adds one SUM all-reduce of the (H, C, C) soft-confusion sums (coda.py:42).
Messages are <= 4 KB: latency-bound; NCCL over NVLink via torch.distributed is the plumbing.
The same code runs on the gloo backend with CPU tensors for the host-logic tests.
"""
from __future__ import annotations

import torch


class LocalComm:
    world = 1
    rank = 0

    def allreduce_sum_(self, t):
        return t

    def allreduce_min_(self, t):
        return t

    def allgather(self, t):
        return t.unsqueeze(0)

    def share_jvec_(self, jvec, sel):
        return jvec

    def barrier(self):
        pass


class TorchComm:
    """torch.distributed process group (NCCL on GPUs, gloo in the CPU tests)."""

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

    def share_jvec_(self, jvec, sel):
        """Owner holds p_h(idx) in ``jvec``; every other rank's ``jvec`` is zero (``label_row`` wrote it that way)."""
        self.dist.all_reduce(jvec, op=self.dist.ReduceOp.SUM, group=self.group)
        return jvec

    def barrier(self):
        self.dist.barrier(group=self.group)


def default_comm():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return TorchComm()
    return LocalComm()


# ---------------------------------------------------------------------------------------
# host-side mirrors of the device merge rules (used by the CPU/gloo tests and the slow paths)
# ---------------------------------------------------------------------------------------
IDX_NONE = (1 << 63) - 1


def merge_records(recs):
    """recs: iterable of (vA, iA, cntA, vB, iB).  Max value, lowest index on equal values,
    counts summed -- the rule of k_select_merge (select.cu)."""
    va, ia, ca, vb, ib = float("-inf"), IDX_NONE, 0, float("-inf"), IDX_NONE
    for (a, i, c, b, j) in recs:
        if a > va or (a == va and i < ia):
            va, ia = a, i
        if b > vb or (b == vb and j < ib):
            vb, ib = b, j
        ca += c
    return va, ia, ca, vb, ib


def choose_among_ties(tie_idx, rng):
    """coda.py:308: random.choice over the tied candidates in ascending index order.  ``random.choice``
    consumes exactly one ``_randbelow(len)``, so choosing a position is RNG-equivalent."""
    order = sorted(int(i) for i in tie_idx)
    return order[rng.choice(range(len(order)))]
