"""CPU model of the PRODUCT's restated algorithm, checked against the oracle / goldens.

The CUDA path does not evaluate the reference's dense (B, C, H, P) tensor program; it evaluates an algebraic
restatement: per-class Beta tables, hypothetical-update pairs with template reuse, EIG as sum_c xi*gain, the
rank-1 marginal refresh with the majority shortcut, fixed-point marginal sums.  This file re-implements exactly
that restatement in NumPy (fp64) -- no GPU, no product code -- and shows it reproduces the reference numbers,
so the CPU-only test tier also guards the mathematics the kernels implement."""
import random

import numpy as np
import pytest
import torch

from helpers import coda_oracle, golden_slab, load_golden
from test_limb_precision_model import class_tables, gain_of


class ProductModel:
    def __init__(self, preds, D, lr=0.01, fx_shift=40):
        self.p = preds.astype(np.float64)                     # (H, N, C)
        self.H, self.N, self.C = preds.shape
        self.D = D.astype(np.float64).copy()
        self.lr, self.fx = lr, float(2 ** fx_shift)
        self.hard = preds.argmax(-1).T                        # (N, H)
        self.E = self.p.sum(0)                                # (N, C) ensemble sums
        self.U = np.einsum("hcs,hns->nc", self.D, self.p)     # un-normalised pi_hat_xi
        self.labeled = np.zeros(self.N, bool)
        self.disagree = (self.hard != self.hard[:, :1]).any(1)
        self.tables = [None] * self.C
        for c in range(self.C):
            self._table(c)
        self._mixture()

    def _table(self, c):
        a = self.D[:, c, c].astype(np.float32).astype(np.float64)
        b = (self.D[:, c, :].sum(1).astype(np.float32) - a.astype(np.float32)).astype(np.float64)
        self.tables[c] = class_tables(a, b)                   # dL, G0, G1, PB

    def _mixture(self):
        xi = self.U / np.maximum(self.U.sum(1, keepdims=True), 1e-12)
        fx = np.rint(xi.astype(np.float32).astype(np.float64) * self.fx).astype(np.int64).sum(0)   # int64 fixed point
        self.xi = xi
        self.pi_hat = fx / fx.sum()
        self.PB = np.stack([t[3] for t in self.tables])       # (C, H)
        self.m0 = self.pi_hat @ self.PB                       # P(best)

    def _gain(self, c, z):
        dL, G0, G1, PB = self.tables[c]
        Dv = np.exp(z @ dL)
        prob = np.where(z > 0, Dv @ G1, Dv @ G0)
        PH = prob / max(prob.sum(), 1e-30)
        return gain_of(PH[None], PB, self.m0, self.pi_hat[c])[0]

    def eig(self):
        templ = {}
        g0 = np.array([self._gain(c, np.zeros(self.H)) for c in range(self.C)])
        out = self.xi @ g0
        for n in range(self.N):
            for c in np.unique(self.hard[n]):
                z = (self.hard[n] == c).astype(np.float64)
                if z.sum() == 1:                              # singleton template: depends on (c, h') only
                    key = (int(c), int(z.argmax()))
                    if key not in templ:
                        templ[key] = self._gain(c, z)
                    g = templ[key]
                else:
                    g = self._gain(c, z)
                out[n] += self.xi[n, c] * (g - g0[c])
        return out

    def add_label(self, idx, t):
        j = self.hard[idx]                                    # p_h(idx)
        self.D[np.arange(self.H), t, j] += np.float32(self.lr)
        tp = np.bincount(j, minlength=self.C).argmax()        # majority shortcut of the rank-1 refresh
        mis = np.nonzero(j != tp)[0]
        inc = self.E[:, tp] + (self.p[mis, :, j[mis]] - self.p[mis, :, tp]).sum(0)
        direct = self.p[np.arange(self.H), :, j].sum(0)
        assert np.allclose(inc, direct, rtol=1e-12, atol=1e-12)
        self.U[:, t] += np.float32(self.lr) * inc
        self.labeled[idx] = True
        self._table(t)
        self._mixture()


@pytest.mark.parametrize("name", ["traj_tiny_h8_n300_c5", "traj_nodiag_h10_n400_c6"])
def test_restated_algorithm_reproduces_the_reference(name):
    g = load_golden(name)
    preds, labels = golden_slab(g)
    m = ProductModel(preds.numpy(), g["init_dirichlets"], lr=g["ctor"].get("learning_rate", 0.01))
    np.testing.assert_allclose(m.pi_hat, g["init_pi_hat"], rtol=2e-6)
    np.testing.assert_allclose(m.m0, g["init_pbest"][0], atol=2e-6)
    for k in range(min(3, int(g["steps"]))):
        e = m.eig()
        ref = g["eig"][k]
        cand = ~np.isnan(ref)
        assert np.array_equal(cand, (~m.labeled) & m.disagree) or not ((~m.labeled) & m.disagree).any()
        np.testing.assert_allclose(e[cand], ref[cand], atol=3e-6)            # reference fp32 noise floor ~1e-6
        pick = int(np.flatnonzero(cand)[np.argmax(e[cand])])
        assert pick == int(g["idx"][k])
        gi = int(g["idx"][k])
        m.add_label(gi, int(labels[gi]))
        np.testing.assert_allclose(m.pi_hat, g["pi_hat"][k], rtol=2e-6)
        np.testing.assert_allclose(m.m0, g["pbest"][k], atol=2e-6)
        np.testing.assert_allclose(m.D[:, int(labels[gi])], g["dir_row"][k], rtol=2e-7)
