"""Ensemble + numeric guards (reference coda/util.py:7-39); the plotting helper is out of scope."""
import torch


class Ensemble:
    def __init__(self, preds, **kwargs):
        self.preds = preds
        self.device = preds.device

    def get_preds(self, **kwargs):
        return self.preds.mean(dim=0)


def _check(t, name, *, raise_err=True):
    bad = ~torch.isfinite(t)
    if bad.any():
        msg = f"[NUMERIC ERROR] {name} has {int(bad.sum())} bad values (NaN/Inf) out of {t.numel()}"
        if raise_err:
            raise RuntimeError(msg)
        print(msg)
