"""``CODA`` -- host-side mirror of the reference selector (coda/coda.py:171-346) over the sm_100a kernels.

Same constructor, same three ``ModelSelector`` calls, same attributes callers read
(``stochastic``, ``unlabeled_idxs``, ``pi_hat``, ``pi_hat_xi``, ``dirichlets``, ``labeled_idxs``,
``labels``, ``q_vals``, ``step``, ``H/N/C``, ``device``), same error types.  The arithmetic is in
``libcoda_b200.so``; there is no CPU route -- a CPU ``dataset.preds`` raises.
"""
from __future__ import annotations

import random

import numpy as np
import torch

from . import _native as nat
from .base import ModelSelector
from .dist import choose_among_ties, default_comm
from .engine import TIE_CAP, Engine


class _Unlabeled:
    """List-like view of the unlabeled item indices (coda.py:200, 323; demo/app.py:188 calls ``remove``).
    The reference keeps a Python list and pays O(N) per removal; this keeps a removed-set and a device mask."""

    def __init__(self, n_lo: int, n_hi: int, on_remove):
        self._lo, self._hi = n_lo, n_hi
        self._removed = set()
        self._on_remove = on_remove

    def remove(self, idx):
        idx = int(idx)
        if not (self._lo <= idx < self._hi) or idx in self._removed:
            raise ValueError("list.remove(x): x not in list")             # coda.py:323 behaviour
        self._removed.add(idx)
        self._on_remove(idx)

    def __contains__(self, idx):
        return self._lo <= int(idx) < self._hi and int(idx) not in self._removed

    def __len__(self):
        return self._hi - self._lo - len(self._removed)

    def __iter__(self):
        rem = self._removed
        return (i for i in range(self._lo, self._hi) if i not in rem)

    def __getitem__(self, k):
        return list(self)[k]


class CODA(ModelSelector):
    def __init__(self, dataset, prefilter_n=0, alpha=0.9, learning_rate=0.01, multiplier=2.0,
                 disable_diag_prior=False, q="eig", *, mode="incremental", comm=None):
        self.dataset = dataset
        self.device = dataset.preds.device
        self.prefilter_n = prefilter_n
        self.disable_diag_prior = disable_diag_prior
        self.q = q
        self.prior_strength = 1 - alpha                     # coda.py:189
        self.update_strength = learning_rate                # coda.py:190
        comm = comm or default_comm()
        n_offset = int(getattr(dataset, "n_offset", 0))
        n_global = int(getattr(dataset, "n_global", dataset.preds.shape[1]))
        self.engine = Engine(dataset.preds, alpha=alpha, learning_rate=learning_rate, multiplier=multiplier,
                             uniform_prior=bool(disable_diag_prior), mode=mode, n_offset=n_offset,
                             n_global=n_global, comm=comm)
        self.H, self.C = self.engine.H, self.engine.C
        self.N = n_global                                   # callers see the whole task (coda.py:183)
        self.labeled_idxs, self.labels = [], []
        self.unlabeled_idxs = _Unlabeled(0, n_global, self.engine.mark_labeled)
        self.q_vals = []
        self.stochastic = False
        self.step = 0
        self.last_report = None

    @classmethod
    def from_args(cls, dataset, args):
        """coda.py:205-213"""
        return cls(dataset, prefilter_n=args.prefilter_n, alpha=args.alpha, learning_rate=args.learning_rate,
                   multiplier=args.multiplier, disable_diag_prior=args.no_diag_prior, q=args.q)

    # -- attributes the reference exposes as tensors ---------------------------------------
    @property
    def dirichlets(self):
        return self.engine.D

    @property
    def pi_hat(self):
        return self.engine.pi_hat

    @property
    def pi_hat_xi(self):
        return self.engine.pi_hat_xi()

    # -- acquisition -------------------------------------------------------------------------
    def get_next_item_to_label(self):
        """coda.py:283-313.  Returns (global item index: int, q: float)."""
        if self.q in ("iid", "uncertainty"):
            return self._select_ablation()                  # coda.py:287-295
        if self.q != "eig":
            raise NotImplementedError(self.q)               # coda.py:297
        eng = self.engine
        eng.score()
        if self.prefilter_n:
            return self._select_prefiltered()
        rep = eng.fetch()
        self.last_report = rep
        eng.check_flags(flags=rep["flags"])
        if rep["n_ties"] == 0:
            raise RuntimeError("no unlabeled items left to select from")
        if rep["n_ties"] > 1:                               # coda.py:308-311
            self.stochastic = True
            if rep["n_ties"] <= len(rep["tie_idx"]):
                ties = rep["tie_idx"]
                idx = choose_among_ties(ties, random)
                q = float(rep["tie_val"][int(np.nonzero(ties == idx)[0][0])])
            else:
                idx, q = self._select_many_ties(rep)
            return idx, q
        return int(rep["tie_idx"][0]), float(rep["tie_val"][0])   # == arg-max, first index wins (coda.py:309)

    def _candidate_mask(self):
        eng = self.engine
        m = (eng.labeled == 0) & (eng.disagree != 0)
        if not bool(m.any()):
            m = eng.labeled == 0                            # coda.py:239 `or self.unlabeled_idxs`
        return m

    def _select_ablation(self):
        """coda.py:287-295: the two ablation acquisitions of the paper (random / ensemble-entropy sampling) followed by
        the same tie rule (coda.py:306-313).  Cold path: a few torch ops on vectors the kernels already produced
        (``ens`` = sum_h preds from the slab scan), nothing slab-sized."""
        eng = self.engine
        if eng.comm.world > 1:
            raise NotImplementedError(f"q={self.q!r} with a sharded slab")
        if self.prefilter_n:
            raise NotImplementedError(f"q={self.q!r} together with prefilter_n")
        mask = self._candidate_mask()
        n = int(mask.sum())
        if n == 0:
            raise RuntimeError("no unlabeled items left to select from")
        if self.q == "iid":
            qv = torch.full((eng.N,), np.float32(1.0 / n).item(), dtype=torch.float32, device=eng.dev)
        else:
            if getattr(self, "_ens_entropy", None) is None:      # non-adaptive: computed once (uncertainty.py:6-11)
                if eng.ens is None:
                    raise RuntimeError("q='uncertainty' needs the ensemble sums (CODA_B200_ENS=0 disables them)")
                mean = eng.ens / float(eng.H)
                self._ens_entropy = -(mean * torch.log(mean + 1e-8)).sum(-1)
            qv = self._ens_entropy
        best = qv[mask].max()
        ties = torch.isclose(qv, best, rtol=1e-8) & mask        # coda.py:307
        nt = int(ties.sum())
        if nt > 1:                                              # coda.py:308-311
            idx = random.choice(torch.nonzero(ties, as_tuple=True)[0].tolist())
            self.stochastic = True
        else:
            idx = int(torch.nonzero(ties, as_tuple=True)[0][0])
        return idx, float(qv[idx])

    def _select_many_ties(self, rep):
        """More than TIE_CAP isclose-ties: evaluate the tie rule on the full vector (cold path)."""
        eng = self.engine
        if eng.comm.world > 1:
            raise NotImplementedError("more than %d tied candidates across shards" % TIE_CAP)
        eig = eng.eig.cpu().numpy()
        cand = np.nonzero(self._candidate_mask().cpu().numpy())[0]
        qv = eig[cand]
        best = np.float32(rep["best_val"])
        tol = np.float32(1e-8) + np.abs(np.float32(1e-8) * best)
        ties = cand[(qv == best) | (np.abs(qv - best) <= tol)]
        idx = choose_among_ties(ties, random)
        return int(idx), float(eig[idx])

    def _select_prefiltered(self):
        """coda.py:221-223: random subsample of the candidates (``--prefilter-n``), then coda.py:306-313
        on the subsample in sample order.  Cold ablation path; uses the EIG vector the kernels produced."""
        eng = self.engine
        if eng.comm.world > 1:
            raise NotImplementedError("prefilter_n with a sharded slab")
        torch.cuda.current_stream(eng.dev).synchronize()
        eng.check_flags(sync=True)
        m = (eng.labeled == 0) & (eng.disagree != 0)
        ids = torch.nonzero(m, as_tuple=True)[0].tolist()
        if self.prefilter_n and len(ids) > self.prefilter_n:
            ids = random.sample(ids, self.prefilter_n)
            self.stochastic = True
        if not ids:
            ids = torch.nonzero(eng.labeled == 0, as_tuple=True)[0].tolist()
        qv = eng.eig[torch.tensor(ids, device=eng.dev)]
        best = qv.max()
        ties = torch.isclose(qv, best, rtol=1e-8)
        if int(ties.sum()) > 1:
            loc = random.choice(torch.nonzero(ties, as_tuple=True)[0].tolist())
            self.stochastic = True
        else:
            loc = int(torch.argmax(qv))
        return ids[loc], float(qv[loc])

    # -- posterior update --------------------------------------------------------------------
    def add_label(self, idx, true_class, selection_prob):
        """coda.py:315-323"""
        idx, true_class = int(idx), int(true_class)
        if not (0 <= true_class < self.C):
            raise IndexError(f"index {true_class} is out of bounds for dimension 1 with size {self.C}")
        if idx not in self.unlabeled_idxs:
            raise ValueError("list.remove(x): x not in list")
        self.engine.post_label(idx, true_class)
        self.labeled_idxs.append(idx)
        self.labels.append(true_class)
        self.q_vals.append(selection_prob)
        self.unlabeled_idxs._removed.add(idx)               # the label kernels already set the device mask

    def get_pbest(self):
        """coda.py:325-332 -> (1, H) float32 tensor on the device."""
        return self.engine.pbest()

    def get_best_model_prediction(self):
        """coda.py:334-346 -> 0-d LongTensor (trap T10); bumps ``step``."""
        self.step += 1
        return self.engine.best_model[0].clone()                # a fresh 0-d tensor like torch.argmax (coda.py:346)
