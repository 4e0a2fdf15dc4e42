"""Host-side logic of coda_b200.CODA with the device engine replaced by a scripted fake (CPU tier):
tie handling and Python-RNG consumption (coda.py:306-313), bookkeeping of add_label (coda.py:320-323),
error behaviour, flag -> exception mapping (util.py:20-25)."""
import random

import numpy as np
import pytest
import torch

from coda_b200 import _native as nat
from coda_b200 import selector as selmod


class FakeComm:
    world = 1


class FakeEngine:
    """Mimics the attributes / calls CODA uses; `script` is a list of reports fetch() returns in order."""
    instances = []

    stream = None

    def __init__(self, preds, **kw):
        self.H, self.N, self.C = (int(s) for s in preds.shape)
        self.comm = FakeComm()
        self.dev = torch.device("cpu")
        self.script, self.posted, self.marked = [], [], []
        self.best_model = torch.zeros(1, dtype=torch.int64)
        self.D = torch.zeros(self.H, self.C, self.C)
        self.pi_hat = torch.full((self.C,), 1.0 / self.C)
        FakeEngine.instances.append(self)

    def report(self):
        pass

    def sync(self):
        pass

    def _on(self):
        import contextlib
        return contextlib.nullcontext()

    def fetch(self):
        return self.script.pop(0)

    def check_flags(self, sync=False, flags=None):
        if flags:
            names = [v for k, v in nat.FLAG_NAMES.items() if flags & k]
            raise RuntimeError(f"[NUMERIC ERROR] {', '.join(names)} has bad values (NaN/Inf)")

    def label_stage(self, idx, cls):
        self.posted.append((idx, cls))

    def api_graph_wanted(self):
        return False

    def label_run(self, eager_report=True):
        pass

    def mark_labeled(self, idx):
        self.marked.append(idx)

    def pbest(self):
        return torch.full((1, self.H), 1.0 / self.H)


class DS:
    def __init__(self, H=3, N=50, C=4):
        self.preds = torch.rand(H, N, C).softmax(-1)
        self.labels = torch.zeros(N, dtype=torch.int64)
        self.device = self.preds.device


def report(ties, vals, flags=0, n_ties=None):
    ties = np.asarray(ties, dtype=np.int64)
    return dict(flags=flags, use_a=True, n_cand=10, best_val=float(max(vals)), best_idx=int(ties[0]),
                n_ties=len(ties) if n_ties is None else n_ties, tie_min=int(ties.min()), tie_idx=ties,
                tie_val=np.asarray(vals, dtype=np.float32))


def fake_build(shards, group, **kw):
    kw.pop("n_global", None)
    return [FakeEngine(p, **kw) for p, _off in shards]


@pytest.fixture()
def sel(monkeypatch):
    monkeypatch.setattr(selmod, "build_engines", fake_build)
    s = selmod.CODA(DS())
    return s, s.engine


def test_single_maximum_is_deterministic(sel):
    s, eng = sel
    eng.script.append(report([17], [0.25]))
    st = random.getstate()
    assert s.get_next_item_to_label() == (17, 0.25)
    assert random.getstate() == st and not s.stochastic            # no RNG consumed (coda.py:309)


def test_ties_use_random_choice_over_ascending_candidates(sel):
    s, eng = sel
    ties, vals = [41, 7, 19], [0.5, 0.5, 0.5]                      # device order is arbitrary
    for seed in range(5):
        eng.script.append(report(ties, vals))
        random.seed(seed)
        idx, q = s.get_next_item_to_label()
        after = random.getstate()
        random.seed(seed)
        want = random.choice(sorted(ties))                          # coda.py:308 on candidates in ascending order
        assert idx == want and q == 0.5 and random.getstate() == after
    assert s.stochastic                                             # coda.py:310-311


def test_add_label_bookkeeping_and_errors(sel):
    s, eng = sel
    s.add_label(5, 2, 0.125)
    assert eng.posted == [(5, 2)] and s.labeled_idxs == [5] and s.labels == [2] and s.q_vals == [0.125]
    assert 5 not in s.unlabeled_idxs and len(s.unlabeled_idxs) == 49
    with pytest.raises(ValueError):                                 # list.remove semantics (coda.py:323)
        s.add_label(5, 2, 0.0)
    with pytest.raises(IndexError):
        s.add_label(6, 99, 0.0)
    s.unlabeled_idxs.remove(9)                                      # demo/app.py:188
    assert eng.marked == [9] and len(s.unlabeled_idxs) == 48
    b = s.get_best_model_prediction()
    assert s.step == 1 and b.dim() == 0 and b.dtype == torch.int64  # trap T10


def test_flags_become_the_reference_runtime_error(sel):
    s, eng = sel
    eng.script.append(report([3], [0.1], flags=nat.FLAG_NONFINITE_EIG))
    with pytest.raises(RuntimeError, match=r"\[NUMERIC ERROR\]"):
        s.get_next_item_to_label()


def test_no_candidates_and_unknown_acquisition(sel, monkeypatch):
    s, eng = sel
    eng.script.append(report([0], [0.0], n_ties=0))
    with pytest.raises(RuntimeError, match="no unlabeled"):
        s.get_next_item_to_label()
    s.q = "bogus"
    with pytest.raises(NotImplementedError):                        # coda.py:297
        s.get_next_item_to_label()


def test_from_args_maps_the_cli_namespace(monkeypatch):
    monkeypatch.setattr(selmod, "build_engines", fake_build)

    class A:
        prefilter_n = 7; alpha = 0.8; learning_rate = 0.05; multiplier = 1.5; no_diag_prior = True; q = "eig"
    s = selmod.CODA.from_args(DS(), A)                              # coda.py:205-213
    assert (s.prefilter_n, s.disable_diag_prior, s.q) == (7, True, "eig")
    assert abs(s.prior_strength - 0.2) < 1e-12 and s.update_strength == 0.05
