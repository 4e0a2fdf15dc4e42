"""Compact slab (top-K scores + uniform remainder; BASELINE.json configs[4] cannot exist as a dense tensor): the
compact kernels must produce what the dense path -- and the reference's algorithm -- produce on the densified slab."""
import random

import numpy as np
import pytest
import torch

from helpers import coda_oracle

pytestmark = pytest.mark.gpu

EIG_ATOL = 5e-6


def _case(H, N, C, K, seed):
    from coda_b200 import CompactSlab
    from coda_b200.synth import synth_compact
    ids, probs, labels = synth_compact(H, N, C, K, seed=seed)
    slab = CompactSlab(ids, probs, C)
    return slab, slab.densify(), labels


@pytest.mark.parametrize("shape", [(24, 1200, 30, 4, 3), (40, 700, 150, 3, 5), (9, 500, 8, 2, 7)])
def test_compact_slab_follows_the_oracle_on_the_densified_slab(shape):
    from coda_b200 import CODA, CompactDataset, TensorDataset
    H, N, C, K, seed = shape
    slab, dense, labels = _case(H, N, C, K, seed)
    assert torch.allclose(dense.sum(-1), torch.ones(H, N), atol=1e-5) and float(dense.min()) >= 0
    assert torch.equal(dense.argmax(-1), slab.ids[..., 0].long())
    dev = torch.device("cuda:0")
    random.seed(0)
    ora = coda_oracle.OracleSelector(dense)
    random.seed(0)
    sel = CODA(CompactDataset(slab.to(dev), labels.to(dev)))
    random.seed(0)
    twin = CODA(TensorDataset(dense.to(dev), labels.to(dev)))              # the dense kernels on the densified slab
    # integer work: hard predictions, unanimity, and the fixed-point confusion sums -> identical posterior bits
    assert torch.equal(sel.engine.hard, twin.engine.hard) and torch.equal(sel.engine.disagree, twin.engine.disagree)
    assert torch.equal(sel.dirichlets, twin.dirichlets)
    np.testing.assert_allclose(sel.dirichlets.cpu().numpy(), ora.dirichlets.numpy(), rtol=3e-6, atol=1e-7)
    np.testing.assert_allclose(sel.pi_hat.cpu().numpy(), ora.pi_hat.numpy(), rtol=5e-6)
    np.testing.assert_allclose(sel.pi_hat_xi.cpu().numpy(), ora.pi_hat_xi.numpy(), rtol=1e-5, atol=1e-9)
    for _ in range(4):
        i_ref, q_ref = ora.get_next_item_to_label()
        i, q = sel.get_next_item_to_label()
        got = sel.engine.eig.cpu().numpy()[np.asarray(ora.last_cand)]
        np.testing.assert_allclose(got, ora.last_q.numpy(), atol=EIG_ATOL)
        assert float(ora.last_q[ora.last_cand.index(i)]) >= float(ora.last_q.max()) - EIG_ATOL
        t = int(labels[i_ref])
        ora.add_label(i_ref, t, q_ref)
        sel.add_label(i_ref, t, q)
        assert int(ora.get_best_model_prediction()) == int(sel.get_best_model_prediction())
        np.testing.assert_allclose(sel.get_pbest().cpu().numpy(), ora.get_pbest().numpy(), atol=1e-5)
        np.testing.assert_allclose(sel.pi_hat.cpu().numpy(), ora.pi_hat.numpy(), rtol=5e-6)
        assert torch.equal(sel.dirichlets[:, t], torch.from_numpy(ora.dirichlets[:, t].numpy()).to(dev)) or \
            np.allclose(sel.dirichlets[:, t].cpu().numpy(), ora.dirichlets[:, t].numpy(), rtol=3e-6)


def test_compact_slab_shards_and_device_loop():
    from coda_b200 import CODA, CompactDataset
    slab, dense, labels = _case(24, 1500, 30, 4, 11)
    dev = torch.device("cuda:0")
    one = CODA(CompactDataset(slab.to(dev), labels.to(dev)))
    many = CODA(CompactDataset(slab.to(dev), labels.to(dev)), shards=3)
    assert torch.equal(one.dirichlets, many.dirichlets) and torch.equal(one.pi_hat, many.pi_hat)
    one.run_steps(6, labels)
    many.run_steps(6, labels)
    assert one.history()[0].tolist() == many.history()[0].tolist()
    assert torch.equal(one.dirichlets, many.dirichlets) and torch.equal(one.pi_hat, many.pi_hat)
    assert torch.equal(one.get_pbest(), many.get_pbest())
    # and the device loop picks what the oracle picks on the densified slab
    random.seed(0)
    ora = coda_oracle.OracleSelector(dense)
    picks = []
    for _ in range(6):
        i, q = ora.get_next_item_to_label()
        picks.append(i)
        ora.add_label(i, int(labels[i]), q)
    if not ora.stochastic:                                   # no isclose tie on the way: arg-max is the reference's rule
        assert one.history()[0].tolist() == picks
        np.testing.assert_allclose(one.get_pbest().cpu().numpy(), ora.get_pbest().numpy(), atol=1e-5)


@pytest.mark.parametrize("shape", [(24, 1500, 30, 4, 11), (17, 3000, 300, 3, 2)])
def test_inverted_index_refresh_equals_the_slab_scan(shape, monkeypatch):
    """The rank-1 marginal refresh from the inverted index (H short lists, int64 fixed-point scatter) against the kernel that
    scans the whole compact slab: same picks, marginals within fp32 rounding, and the index path itself bit-identical for
    1 and 3 shards; the scatter target is back to zero after every step."""
    from coda_b200 import CODA, CompactDataset
    H, N, C, K, seed = shape
    slab, _dense, labels = _case(H, N, C, K, seed)
    dev = torch.device("cuda:0")
    monkeypatch.setenv("CODA_B200_COMPACT_INDEX", "0")
    scan = CODA(CompactDataset(slab.to(dev), labels.to(dev)))
    assert scan.engine.cidx is None
    monkeypatch.setenv("CODA_B200_COMPACT_INDEX", "1")
    idx = CODA(CompactDataset(slab.to(dev), labels.to(dev)))
    many = CODA(CompactDataset(slab.to(dev), labels.to(dev)), shards=3)
    ix = idx.engine.cidx
    assert ix is not None and int(ix["off"][-1]) == int((slab.ids < C).sum())
    for step in range(5):
        random.seed(step)
        i, q = idx.get_next_item_to_label()
        random.seed(step)
        k, _ = many.get_next_item_to_label()
        assert i == k, step                                   # bit-identical state: same scores, same tie draw
        scan.get_next_item_to_label()
        assert float(scan._cat("eig")[i]) >= scan.last_report["best_val"] - 1e-6   # the scan path scores the same item (near-)best
        for s in (idx, scan, many):
            s.add_label(i, int(labels[i]), q)
        torch.cuda.synchronize()
        np.testing.assert_allclose(idx.engine.U.cpu().numpy(), scan.engine.U.cpu().numpy(), rtol=2e-6)
        np.testing.assert_allclose(idx.pi_hat.cpu().numpy(), scan.pi_hat.cpu().numpy(), rtol=1e-6)
        assert torch.equal(idx.pi_hat, many.pi_hat) and torch.equal(idx.dirichlets, many.dirichlets)
        assert int(ix["delta"].abs().max()) == 0
    idx.run_steps(4, labels)
    many.run_steps(4, labels)
    assert idx.history()[0].tolist() == many.history()[0].tolist()
    assert torch.equal(idx.pi_hat, many.pi_hat)
