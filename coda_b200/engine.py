"""Device-side state and kernel sequencing for one shard of the CODA acquisition path.

PyTorch is used for device memory, streams and (through ``coda_b200.dist``) NCCL -- the
arithmetic of the hot path is in the C-ABI library (``include/coda_b200.h``).

Modes (what is kept between steps; results are the same):
  ``incremental``   the normalised P(best | hypothetical) row of every pair is cached; a label of
                    class t only invalidates the pairs of class t (coda.py:317 touches row t only)
                    and the marginal refresh is the rank-1 column update of coda.py:319.
  ``recompute``     every step recomputes all pairs from the tables (no row cache).
  ``recompute_all`` additionally rebuilds all class tables and re-runs the full slab pass of
                    ``update_pi_hat`` every step -- the reference's literal per-step work.
"""
from __future__ import annotations

import contextlib
import math
import os

import numpy as np
import torch

from . import _native as nat
from .dist import LocalComm

TIE_CAP = 256
MODES = ("incremental", "recompute", "recompute_all")
TABLE_BATCH_BYTES = 512 << 20


def _ptr(t):
    return t.data_ptr() if t is not None else None


class Engine:
    def __init__(self, preds: torch.Tensor, *, alpha: float, learning_rate: float, multiplier: float,
                 uniform_prior: bool, hyp_w: float = 1.0, mode: str = "incremental", n_offset: int = 0,
                 n_global: int | None = None, comm=None):
        if mode not in MODES:
            raise ValueError(f"mode must be one of {MODES}")
        if not (isinstance(preds, torch.Tensor) and preds.is_cuda):
            raise RuntimeError("coda_b200: dataset.preds must live on a CUDA (sm_100a) device; "
                               "there is no CPU path in this package")
        if preds.dtype != torch.float32 or preds.dim() != 3:
            raise TypeError("coda_b200: preds must be a float32 (H, N, C) tensor (coda/datasets.py:14)")
        if not preds.is_contiguous():
            raise ValueError("coda_b200: preds must be contiguous (H, N, C)")
        nat.require_device()
        self.lib = nat.load()
        # sector gathers of the rank-1 refresh: ask for 64-byte L2 fills (the default 128 doubles their DRAM traffic;
        # streaming kernels measured the same at 64 and 128)
        nat.check(self.lib.coda_b200_set_l2_fetch_granularity(int(os.environ.get("CODA_B200_L2_FETCH", "64"))), "l2_fetch")
        self.preds = preds
        self.dev = preds.device
        self.H, self.N, self.C = (int(s) for s in preds.shape)
        self.Hp = (self.H + 31) // 32 * 32
        self.W = self.Hp // 32
        self.P = 256
        self.mode = mode
        self.comm = comm or LocalComm()
        self.n_offset = int(n_offset)
        self.n_global = int(n_global if n_global is not None else self.N)
        self.lr = float(learning_rate)
        self.hyp_w = float(hyp_w)
        self.prior_strength = 1 - alpha                       # coda.py:189
        self.multiplier = float(multiplier)
        self.uniform_prior = bool(uniform_prior)
        if self.H > 1024:
            raise NotImplementedError("coda_b200: H > 1024 models is not supported yet")
        if self.C > 4096:
            raise NotImplementedError("coda_b200: C > 4096 classes is not supported yet")
        self.fx_shift = max(8, min(40, 62 - math.ceil(math.log2(self.n_global + 1))))
        self.counters = {"launches": 0}
        # side-stream refresh of the class-t tables/rows concurrently with the marginal pass.  On a full-size shard
        # both want every SM (the tensor-core CTAs take a whole SM's shared memory) and the overlap is neutral; on
        # small shards (multi-GPU) the few row tiles leave most SMs to the marginal pass and the two overlap.
        # CODA_B200_OVERLAP=0/1 forces it; default: decided after the pair structure is known (see _build_pairs).
        self.overlap_env = os.environ.get("CODA_B200_OVERLAP")
        self.overlap = self.overlap_env == "1"
        self.profile, self.profile_only = None, None
        with torch.cuda.device(self.dev):
            self._alloc_static()
            self._construct()

    # ------------------------------------------------------------------------------ utils
    def _s(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    def _call(self, name, *args, n=1):
        prof = self.profile
        if prof is not None and (self.profile_only is None or name in self.profile_only):
            st = torch.cuda.current_stream(self.dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            rc = getattr(self.lib, name)(*args)
            e1.record(st)
            prof.setdefault(name, []).append((e0, e1))
        else:
            rc = getattr(self.lib, name)(*args)
        nat.check(rc, name)
        self.counters["launches"] += n

    def start_profile(self, only=None):
        """Bracket every C-ABI launch (or just ``only``) with CUDA events on the launching stream."""
        self.profile, self.profile_only = {}, (set(only) if only else None)

    def stop_profile(self):
        """-> {entry point: (launches, total ms)}; synchronises."""
        torch.cuda.synchronize(self.dev)
        out = {k: (len(v), float(sum(a.elapsed_time(b) for a, b in v))) for k, v in (self.profile or {}).items()}
        self.profile = None
        return out

    def _z(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype, device=self.dev)

    def _e(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.dev)

    # --------------------------------------------------------------------------- buffers
    def _alloc_static(self):
        H, N, C, Hp, P = self.H, self.N, self.C, self.Hp, self.P
        self.hard = self._e((N, H), torch.int16)              # uint16 bit patterns
        self.pseudo = self._e((N,), torch.int32)
        self.disagree = self._e((N,), torch.uint8)
        self.labeled = self._z((N,), torch.uint8)
        self.conf_fx = self._z((H, C, C), torch.int64)
        self.D = self._e((H, C, C), torch.float32)
        self.U = self._e((N, C), torch.float32)
        self.pisum = self._z((C,), torch.int64)
        self.grid = torch.linspace(1e-6, 1 - 1e-6, P).to(self.dev)   # coda.py:86, built on the host (trap T1)
        self.dL = self._e((C, H, P), torch.float32)
        self.G0T = self._z((C, P, Hp), torch.float32)
        self.G1T = self._z((C, P, Hp), torch.float32)
        self.PB = self._z((C, Hp), torch.float32)
        # bf16 limb tables in tensor-core operand order (pairs_tc.cu); SIMT kernel (pairs.cu) when Hp > 256
        self.use_tc = Hp <= 256 and os.environ.get("CODA_B200_TC", "1") != "0"
        self.dLb = self._z((C, Hp // 32, 3, 256 * 32), torch.bfloat16) if self.use_tc else None
        self.Gb = self._z((C, 16, 4, Hp * 16), torch.bfloat16) if self.use_tc else None
        self.pi_hat = self._z((C,), torch.float32)
        self.m0 = self._z((Hp,), torch.float32)
        self.hb = self._z((1,), torch.float32)
        self.best_model = self._z((1,), torch.int64)
        self.eig = self._e((N,), torch.float32)
        self.nblocks = int(self.lib.coda_b200_eig_blocks(N))
        self.partials = self._z((self.nblocks, 5), torch.int64)
        # report block: one D2H copy per step.  [flags | best rec (5) | tie hdr (2) | tie idx | tie val]
        self.rep = self._z((8 + TIE_CAP + TIE_CAP // 2,), torch.int64)
        self.flags = self.rep[0:1].view(torch.int32)[0:1]
        self.bestrec = self.rep[1:6]
        self.tie_hdr = self.rep[6:8]
        self.tie_idx = self.rep[8:8 + TIE_CAP]
        self.tie_val = self.rep[8 + TIE_CAP:].view(torch.float32)[:TIE_CAP]
        self.rep_host = torch.zeros(self.rep.shape, dtype=torch.int64).pin_memory()
        self.sel = self._z((2,), torch.int64)
        self.sel_host = torch.zeros((2,), dtype=torch.int64).pin_memory()
        self.jvec = self._z((H,), torch.int32)
        self.terms = self._z((2 + 8 * H + 2,), torch.int64).view(torch.int32)[: 2 + 8 * H]   # 8-byte aligned
        # ensemble sums E[n][c] (N*C floats) feed pi_rank1's majority shortcut; CODA_B200_ENS=0 disables it
        self.ens = self._e((N, C), torch.float32) if os.environ.get("CODA_B200_ENS", "1") != "0" else None
        cls_per_batch = max(1, min(C, TABLE_BATCH_BYTES // max(1, self.lib.coda_b200_tables_scratch_bytes(H, 1))))
        self.table_batch = int(cls_per_batch)
        self.scratch = self._e((int(self.lib.coda_b200_tables_scratch_bytes(H, self.table_batch)),), torch.uint8)

    # ---------------------------------------------------------------------- construction
    def _construct(self):
        H, N, C, s = self.H, self.N, self.C, self._s()
        self._call("coda_b200_scan_slab", _ptr(self.preds), H, N, C, _ptr(self.hard), _ptr(self.pseudo),
                   _ptr(self.disagree), _ptr(self.ens), _ptr(self.flags), s)
        if C <= 128:
            order = torch.argsort(self.pseudo).to(torch.int32)      # init-time plumbing: any grouping by label will do
            self._call("coda_b200_confusion_sorted", _ptr(self.preds), _ptr(self.pseudo), _ptr(order), H, N, C,
                       self.fx_shift, _ptr(self.conf_fx), s)
            del order
        else:
            self._call("coda_b200_confusion_accum", _ptr(self.preds), _ptr(self.pseudo), H, N, C, self.fx_shift,
                       _ptr(self.conf_fx), s)
        self.comm.allreduce_sum_(self.conf_fx)
        self._call("coda_b200_init_dirichlets", _ptr(self.conf_fx), H, C, self.fx_shift, self.prior_strength,
                   self.multiplier, int(self.uniform_prior), _ptr(self.D), s)
        self.conf_fx = None                                     # H*C*C int64, only needed once
        self._refresh_marginals_full()
        self._build_pairs()
        self._build_shadow()
        self._tables(0, C)
        self._mixture()
        self.cache_valid = False     # incremental mode: P(best | hypothetical) rows of every pair are cached
        self.side = torch.cuda.Stream(device=self.dev)
        self.ev_fork, self.ev_join, self.ev_tables = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
        self.pending = None
        self.scored = False
        self.check_flags(sync=True)

    def _refresh_marginals_full(self):
        H, N, C, s = self.H, self.N, self.C, self._s()
        self._call("coda_b200_pi_full", _ptr(self.preds), _ptr(self.D), H, N, C, _ptr(self.U), s)
        self.pisum.zero_()
        self._call("coda_b200_pi_reduce", _ptr(self.U), N, C, self.fx_shift, None, _ptr(self.pisum),
                   _ptr(self.flags), s)
        self.comm.allreduce_sum_(self.pisum)

    def _build_pairs(self):
        H, N, C, W, s = self.H, self.N, self.C, self.W, self._s()
        ent_cnt = self._e((N,), torch.int32)
        cls_heavy = self._z((C,), torch.int32)
        self._call("coda_b200_pair_count", _ptr(self.hard), H, N, C, _ptr(ent_cnt), _ptr(cls_heavy), s)
        self.ent_off = self._z((N + 1,), torch.int64)
        torch.cumsum(ent_cnt, 0, out=self.ent_off[1:])          # init-time plumbing
        heavy = cls_heavy.cpu().numpy().astype(np.int64)        # host sync (construction only)
        n_ent = int(self.ent_off[-1].item())
        per_cls = 1 + H + heavy
        cls_base = np.zeros(C + 1, dtype=np.int64)
        np.cumsum(per_cls, out=cls_base[1:])
        self.npairs = int(cls_base[-1])
        self.n_heavy = int(heavy.sum())
        self.n_entries = n_ent
        if self.npairs >= 2 ** 31:
            raise NotImplementedError("coda_b200: more than 2^31 pairs in one shard")
        self.cls_base_host = cls_base
        self.cls_base = torch.from_numpy(cls_base).to(self.dev)
        # tiles of <= 32 (SIMT) or <= 128 (tcgen05) same-class pairs
        def make_tiles(width):
            nt = (per_cls + width - 1) // width
            tile_off = np.zeros(C + 1, dtype=np.int64)
            np.cumsum(nt, out=tile_off[1:])
            cls_of_tile = np.repeat(np.arange(C, dtype=np.int64), nt)
            k_in_cls = np.arange(int(tile_off[-1]), dtype=np.int64) - tile_off[cls_of_tile]
            start = cls_base[cls_of_tile] + width * k_in_cls
            cnt = np.minimum(width, per_cls[cls_of_tile] - width * k_in_cls)
            tiles = np.stack([cls_of_tile, start, cnt, np.zeros_like(cnt)], axis=1).astype(np.int32)
            return tile_off, int(nt.max()), torch.from_numpy(tiles).to(self.dev)
        width = 128 if self.use_tc else 32
        tile_off, self.max_cls_tiles, self.tiles = make_tiles(width)
        self.tile_off_host = tile_off
        self.tile_off = torch.from_numpy(tile_off).to(self.dev)
        self.ntiles = int(tile_off[-1])
        if self.overlap_env is None:    # measured: +5 % at 2 GPUs, neutral on one full-size shard
            self.overlap = self.comm.world > 1 or self.max_cls_tiles * 2 <= int(self.lib.coda_b200_sm_count())
        self.ent_pair = self._e((max(1, n_ent),), torch.int32)
        self.ent_cls = self._e((max(1, n_ent),), torch.int16)
        self.zmask = self._e((self.npairs, W), torch.int32)
        self.pair_cls = self._e((self.npairs,), torch.int16)
        self.pair_item = torch.full((self.npairs,), -1, dtype=torch.int32, device=self.dev)
        cursor = self._z((C,), torch.int32)
        self._call("coda_b200_pair_fill", _ptr(self.hard), H, N, C, _ptr(self.ent_off), _ptr(self.cls_base),
                   _ptr(cursor), _ptr(self.ent_pair), _ptr(self.ent_cls), _ptr(self.zmask), _ptr(self.pair_cls),
                   _ptr(self.pair_item), s, n=2)
        # ELL copy of the per-item lists when the longest one fits a warp (eig_points then needs no offset lookup)
        max_cnt = int(ent_cnt.max().item()) if N else 0
        self.ell, self.ell_k = None, 0
        if 0 < max_cnt <= 32:
            self.ell_k = (max_cnt + 3) // 4 * 4
            self.ell = self._e((N, self.ell_k, 2), torch.int32)
            self._call("coda_b200_ell_build", _ptr(self.ent_off), _ptr(self.ent_pair), _ptr(self.ent_cls), N,
                       self.ell_k, _ptr(self.ell), s)
        self.gain = self._z((self.npairs,), torch.float32)
        self.ph_cache = self._e((self.npairs, self.Hp), torch.float32) if self.mode == "incremental" else None

    def _build_shadow(self):
        """Class-major shadow copy of as many models as spare HBM allows (least accurate first)."""
        self.shadow, self.slot_of_model, self.n_shadow = None, None, 0
        if self.mode == "recompute_all" or os.environ.get("CODA_B200_SHADOW", "1") == "0":
            return
        H, N, C = self.H, self.N, self.C
        torch.cuda.synchronize(self.dev)
        torch.cuda.empty_cache()
        free, _total = torch.cuda.mem_get_info(self.dev)
        reserve = int(float(os.environ.get("CODA_B200_SHADOW_RESERVE_GB", "8")) * 2 ** 30)
        per_model = N * C * 4
        S = int(min(H, max(0, (free - reserve) // per_model)))
        cap = os.environ.get("CODA_B200_SHADOW_MODELS")
        if cap is not None:
            S = min(S, int(cap))
        if S <= 0:
            return
        # disagreement of every model with the ensemble pseudo-label: the models that will need gathers most often
        dis = torch.zeros(H, dtype=torch.int64, device=self.dev)
        step = max(1, (64 << 20) // max(1, H))
        for n0 in range(0, N, step):
            blk = self.hard[n0:n0 + step].to(torch.int32) & 0xFFFF
            dis += (blk != self.pseudo[n0:n0 + step, None]).sum(0)
        order = torch.argsort(dis, descending=True, stable=True)[:S].to(torch.int32)
        slot = torch.full((H,), -1, dtype=torch.int32, device=self.dev)
        slot[order.long()] = torch.arange(S, dtype=torch.int32, device=self.dev)
        self.shadow = self._e((S, C, N), torch.float32)
        self._call("coda_b200_shadow_build", _ptr(self.preds), H, N, C, _ptr(order), S, _ptr(self.shadow), self._s())
        self.slot_of_model, self.n_shadow = slot, S

    # ------------------------------------------------------------------------ step pieces
    def _tables(self, lo, hi):
        H, C, s = self.H, self.C, self._s()
        for b0 in range(lo, hi, self.table_batch):
            b1 = min(hi, b0 + self.table_batch)
            self._call("coda_b200_beta_tables", _ptr(self.D), _ptr(self.grid), H, C, self.P, self.hyp_w, b0, b1, None,
                       _ptr(self.scratch), _ptr(self.dL), _ptr(self.G0T), _ptr(self.G1T), _ptr(self.PB),
                       _ptr(self.dLb), _ptr(self.Gb), _ptr(self.flags), s, n=3)

    def _mixture(self):
        self._call("coda_b200_mixture", _ptr(self.pisum), _ptr(self.PB), self.H, self.C, _ptr(self.pi_hat),
                   _ptr(self.m0), _ptr(self.hb), _ptr(self.best_model), _ptr(self.flags), self._s())

    def _pair_rows(self, tile_lo, tile_hi, gains=True, sel=False):
        tail = (_ptr(self.PB), _ptr(self.m0) if gains else None, _ptr(self.pi_hat) if gains else None, self.H,
                _ptr(self.ph_cache), _ptr(self.gain) if gains else None, _ptr(self.sel) if sel else None,
                _ptr(self.tile_off) if sel else None, _ptr(self.flags), self._s())
        if self.use_tc:
            self._call("coda_b200_pair_rows_tc", _ptr(self.tiles), int(tile_lo), int(tile_hi), _ptr(self.zmask),
                       _ptr(self.dLb), _ptr(self.Gb), *tail)
        else:
            self._call("coda_b200_pair_rows", _ptr(self.tiles), int(tile_lo), int(tile_hi), _ptr(self.zmask),
                       _ptr(self.dL), _ptr(self.G0T), _ptr(self.G1T), *tail)

    def _pair_gain(self, filt, cls=None, on_dev=False):
        self._call("coda_b200_pair_gain", _ptr(self.ph_cache), _ptr(self.pair_cls), self.npairs, self.H,
                   _ptr(self.PB), _ptr(self.m0), _ptr(self.pi_hat), _ptr(self.gain),
                   _ptr(self.sel) if on_dev else None, _ptr(self.cls_base), int(cls or 0), filt, self._s())

    def post_label(self, idx_global: int | None = None, true_class: int | None = None, device_sel: bool = False):
        """coda.py:316-319: posterior update + marginal refresh + the tables that depend on them.
        ``device_sel``: the {local idx, class} record is already in ``self.sel`` on the device (host-free loop).

        incremental mode: the class-t tables and the cached rows of the class-t pairs only need the new D, so
        they are rebuilt on a side stream (FP32-pipe bound) while the main stream does the HBM-bound marginal
        refresh; the two join before the mixture."""
        H, N, C, s = self.H, self.N, self.C, self._s()
        if not device_sel:
            loc = idx_global - self.n_offset
            self.sel_host[0] = loc if 0 <= loc < N else -1
            self.sel_host[1] = true_class
            self.sel.copy_(self.sel_host, non_blocking=True)
        self._call("coda_b200_label_row", _ptr(self.hard), H, N, _ptr(self.sel), _ptr(self.jvec), _ptr(self.labeled), s)
        if self.comm.world > 1:
            self.comm.share_jvec_(self.jvec, self.sel)
        self._call("coda_b200_label_apply", _ptr(self.D), H, C, _ptr(self.sel), _ptr(self.jvec), self.lr, s)
        if self.mode == "recompute_all":
            self._refresh_marginals_full()
            self._tables(0, C)
        else:
            main = torch.cuda.current_stream(self.dev)
            overlap = self.overlap and self.mode == "incremental" and self.cache_valid
            if overlap:
                self.ev_fork.record(main)
                self.side.wait_event(self.ev_fork)
                ctx = torch.cuda.stream(self.side)
            else:
                ctx = contextlib.nullcontext()
            with ctx:
                if device_sel:
                    self._call("coda_b200_beta_tables", _ptr(self.D), _ptr(self.grid), H, C, self.P, self.hyp_w, 0, 1,
                               _ptr(self.sel), _ptr(self.scratch), _ptr(self.dL), _ptr(self.G0T), _ptr(self.G1T),
                               _ptr(self.PB), _ptr(self.dLb), _ptr(self.Gb), _ptr(self.flags), self._s(), n=3)
                else:
                    self._tables(true_class, true_class + 1)
                if self.mode == "incremental" and self.cache_valid:
                    # refresh the cached rows of the class-t pairs (no gains: m0 / pi_hat are not final yet)
                    if overlap:
                        self.ev_tables.record(self.side)
                    if device_sel:
                        self._pair_rows(0, self.max_cls_tiles, gains=False, sel=True)
                    else:
                        self._pair_rows(self.tile_off_host[true_class], self.tile_off_host[true_class + 1], gains=False)
            self._call("coda_b200_pi_rank1", _ptr(self.preds), _ptr(self.ens), _ptr(self.shadow),
                       _ptr(self.slot_of_model), H, N, C, _ptr(self.sel), _ptr(self.jvec), self.lr, self.fx_shift,
                       _ptr(self.terms), _ptr(self.U), _ptr(self.pisum), _ptr(self.flags),
                       4 if overlap else 8, s, n=3)
            self.comm.allreduce_sum_(self.pisum)
            if overlap:
                self.ev_join.record(self.side)
                main.wait_event(self.ev_tables)     # the mixture needs PB[t]; the rows are awaited in score()
                self.pending = (None if device_sel else int(true_class), bool(device_sel))
        self._mixture()
        self.scored = False

    def score(self, ties=True):
        """coda.py:235-281 + 306-309: EIG of every item, candidate arg-max, isclose tie scan (enqueue only).
        ``ties=False`` (host-free loop): stop after the merged arg-max record."""
        if self.scored:
            return
        N, C, s = self.N, self.C, self._s()
        if self.mode == "incremental":
            if not self.cache_valid:
                self._pair_rows(0, self.ntiles, gains=False)    # fill the row cache once
                self.cache_valid = True
            if self.pending is None:
                self._pair_gain(0)
            else:   # gains of every other class while the side stream still rebuilds the class-t rows
                cls, on_dev = self.pending
                self._pair_gain(1, cls, on_dev)
                torch.cuda.current_stream(self.dev).wait_event(self.ev_join)
                self._pair_gain(2, cls, on_dev)
                self.pending = None
        else:
            self._pair_rows(0, self.ntiles)
        self._call("coda_b200_eig_points", _ptr(self.U), N, C, _ptr(self.ent_off), _ptr(self.ent_pair),
                   _ptr(self.ent_cls), _ptr(self.gain), _ptr(self.cls_base), _ptr(self.labeled), _ptr(self.disagree),
                   self.n_offset, _ptr(self.ell), self.ell_k, _ptr(self.eig), _ptr(self.partials), _ptr(self.flags), s)
        self._call("coda_b200_select_merge", _ptr(self.partials), self.nblocks, _ptr(self.bestrec), s)
        if self.comm.world > 1:
            recs = self.comm.allgather(self.bestrec)            # (world, 5)
            self._call("coda_b200_select_merge", _ptr(recs), self.comm.world, _ptr(self.bestrec), s)
        if not ties:
            return
        self._call("coda_b200_ties", _ptr(self.eig), N, _ptr(self.labeled), _ptr(self.disagree), self.n_offset,
                   _ptr(self.bestrec), TIE_CAP, _ptr(self.tie_hdr), _ptr(self.tie_idx), _ptr(self.tie_val), s, n=2)
        if self.comm.world > 1:
            self.rep_all = self.comm.allgather(self.rep)        # every rank's tie list, judged against the global best
        self.scored = True

    def device_step(self, labels_dev: torch.Tensor, step: int, hist_idx=None, hist_q=None):
        """One acquisition step with no host round trip (bench ``value`` loop): score, pick the lowest tied
        index, look the label up on the device (coda/oracle.py:23-24), update the posterior.  The pick is the merged
        arg-max record (first index wins on equal values, coda.py:309); the isclose tie rule needs the host RNG and
        is part of the API path only."""
        self.score(ties=False)
        self._call("coda_b200_device_pick", None, _ptr(self.bestrec), _ptr(labels_dev), self.n_offset, self.N,
                   _ptr(self.eig), _ptr(self.sel), _ptr(hist_idx), _ptr(hist_q), int(step), self._s())
        self.post_label(device_sel=True)

    def fetch(self):
        """One D2H copy of the report block + a stream sync.  Returns a dict of host values."""
        if self.comm.world > 1:
            return self._fetch_sharded()
        self.rep_host.copy_(self.rep, non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()
        r = self.rep_host.numpy()
        flags = int(r[0:1].view(np.int32)[0])
        use_a = int(r[3]) > 0
        bits = int(r[1] if use_a else r[4])
        best_val = float(np.array([bits & 0xFFFFFFFF], dtype=np.uint32).view(np.float32)[0])
        best_idx = int(r[2] if use_a else r[5])
        n_ties = int(r[6])
        k = min(n_ties, TIE_CAP)
        tie_idx = r[8:8 + k].copy()
        tie_val = r[8 + TIE_CAP:].view(np.float32)[:k].copy()
        return dict(flags=flags, use_a=use_a, n_cand=int(r[3]), best_val=best_val, best_idx=best_idx,
                    n_ties=n_ties, tie_min=int(r[7]), tie_idx=tie_idx, tie_val=tie_val)

    def _fetch_sharded(self):
        if getattr(self, "rep_all_host", None) is None or self.rep_all_host.shape != self.rep_all.shape:
            self.rep_all_host = torch.zeros(self.rep_all.shape, dtype=torch.int64).pin_memory()
        self.rep_all_host.copy_(self.rep_all, non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()
        allr = self.rep_all_host.numpy()                        # (world, len(rep))
        r0 = allr[0]
        flags = 0
        for r in allr:
            flags |= int(r[0:1].view(np.int32)[0])
        use_a = int(r0[3]) > 0                                  # bestrec is the merged (global) record on every rank
        bits = int(r0[1] if use_a else r0[4])
        best_val = float(np.array([bits & 0xFFFFFFFF], dtype=np.uint32).view(np.float32)[0])
        n_ties = int(sum(int(r[6]) for r in allr))
        idxs, vals = [], []
        for r in allr:
            k = min(int(r[6]), TIE_CAP)
            idxs.append(r[8:8 + k])
            vals.append(r[8 + TIE_CAP:].view(np.float32)[:k])
        return dict(flags=flags, use_a=use_a, n_cand=int(r0[3]), best_val=best_val,
                    best_idx=int(r0[2] if use_a else r0[5]), n_ties=n_ties,
                    tie_min=int(min(int(r[7]) for r in allr)), tie_idx=np.concatenate(idxs),
                    tie_val=np.concatenate(vals))

    def check_flags(self, sync=False, flags=None):
        if flags is None:
            flags = int(self.flags.item()) if sync else 0
        if not flags:
            return
        if flags & nat.FLAG_RANGE_INPUT and not flags & nat.FLAG_NONFINITE_INPUT:
            raise ValueError("coda_b200: dataset.preds must hold post-softmax scores in [0, 1] (coda/datasets.py:6)")
        names = [v for k, v in nat.FLAG_NAMES.items() if flags & k]
        raise RuntimeError(f"[NUMERIC ERROR] {', '.join(names)} has bad values (NaN/Inf)")   # util.py:20-25

    def mark_labeled(self, idx_global: int):
        loc = idx_global - self.n_offset
        if 0 <= loc < self.N:
            self.labeled[loc] = 1
        self.scored = False

    # ------------------------------------------------------------------------- read-outs
    def pbest(self) -> torch.Tensor:
        return self.m0[: self.H].clone().view(1, self.H)        # coda.py:329 -> (1, H)

    def pi_hat_xi(self) -> torch.Tensor:
        xi = torch.empty_like(self.U)
        scratch = torch.zeros_like(self.pisum)
        self._call("coda_b200_pi_reduce", _ptr(self.U), self.N, self.C, self.fx_shift, _ptr(xi), _ptr(scratch),
                   _ptr(self.flags), self._s())
        return xi
