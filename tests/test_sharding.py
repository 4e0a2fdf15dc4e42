"""N-axis sharding with every shard driven by THIS process (coda_b200.dist.InProcessGroup): the same fused step kernels
and peer-memory exchange as the one-process-per-GPU path, but runnable on a single GPU (each shard on its own stream),
so the driver's 1-GPU test lease covers the exchange logic.  Plus: the host-free CUDA-graph loop (run_steps), shards
that are strided views of the caller's slab, and checkpoint / resume.

Shard-count invariance is exact by construction: selection merges (value, lowest index) records, the marginal sums are
int64 fixed point, and every per-item result is computed by the same kernel from the same replicated tables."""
import random

import numpy as np
import pytest
import torch

from helpers import coda_oracle, golden_names, golden_slab, load_golden

pytestmark = pytest.mark.gpu

EIG_ATOL = 5e-6


def _mk(preds, labels=None, **kw):
    from coda_b200 import CODA, TensorDataset
    dev = torch.device("cuda:0")
    return CODA(TensorDataset(preds.to(dev), None if labels is None else labels.to(dev)), **kw)


@pytest.mark.parametrize("shards", [2, 3])
@pytest.mark.parametrize("mode", ["incremental", "recompute"])
def test_in_process_shards_equal_one_shard_api_path(shards, mode):
    """API path (get_next / add_label / get_best) on `shards` shards of one GPU vs one shard: identical picks and
    RNG use, bit-identical pi_hat / dirichlets / EIG, and both equal to the reference golden."""
    g = load_golden("traj_small_h32_n3000_c10")
    preds, labels = golden_slab(g)
    random.seed(0)
    one = _mk(preds, labels, mode=mode)
    random.seed(0)
    many = _mk(preds, labels, mode=mode, shards=shards)
    assert len(many.engines) == shards and many.group.world == shards
    assert many.engines[0].preds.data_ptr() == many.dataset.preds.data_ptr()      # shard 0 is a view, not a copy
    np.testing.assert_array_equal(many.dirichlets.cpu().numpy(), one.dirichlets.cpu().numpy())
    assert torch.equal(many.pi_hat, one.pi_hat)
    for k in range(int(g["steps"])):
        st = random.getstate()
        i1, q1 = one.get_next_item_to_label()
        after = random.getstate()
        random.setstate(st)
        i2, q2 = many.get_next_item_to_label()
        assert (i1, q1) == (i2, q2) and random.getstate() == after
        assert i1 == int(g["idx"][k])
        assert torch.equal(many.eig, one.eig)
        t = int(labels[i1])
        one.add_label(i1, t, q1)
        many.add_label(i2, t, q2)
        assert int(one.get_best_model_prediction()) == int(many.get_best_model_prediction()) == int(g["best_model"][k])
        assert torch.equal(many.pi_hat, one.pi_hat) and torch.equal(many.dirichlets, one.dirichlets)
        assert torch.equal(many.get_pbest(), one.get_pbest())
        np.testing.assert_allclose(many.get_pbest().cpu().numpy()[0], g["pbest"][k], atol=1e-5)
    assert torch.equal(many.pi_hat_xi, one.pi_hat_xi)
    for e in many.engines[1:]:                                   # the replicas of the small state never diverge
        assert torch.equal(e.D.cpu(), many.engine.D.cpu()) and torch.equal(e.m0.cpu(), many.engine.m0.cpu())


@pytest.mark.parametrize("shards", [1, 2, 4])
def test_graph_loop_equals_eager_loop_and_the_reference(shards):
    """run_steps (one captured CUDA graph per step, exchanges inside the kernels) vs the reference's free-running
    trajectory and vs eager device steps: same picks, same posterior bits."""
    g = load_golden("traj_small_h32_n3000_c10")
    preds, labels = golden_slab(g)
    K = int(g["steps"])
    sel = _mk(preds, labels, shards=shards)
    sel.run_steps(K, labels)
    idx, q, tie = sel.history()
    assert idx.tolist() == [int(i) for i in g["idx"]] and not tie.any()
    np.testing.assert_allclose(q, g["q"], atol=EIG_ATOL)
    np.testing.assert_allclose(sel.get_pbest().cpu().numpy()[0], g["pbest"][-1], atol=1e-5)
    np.testing.assert_allclose(sel.dirichlets.cpu().numpy(), g["final_dirichlets"], rtol=2e-6, atol=1e-7)
    assert sel.labeled_idxs == idx.tolist() and len(sel.unlabeled_idxs) == int(g["N"]) - K
    # eager twin, one step at a time
    twin = _mk(preds, labels)
    lab = labels.cuda()
    for k in range(K):
        twin.engine.device_step(lab)
    assert twin.history()[0].tolist() == idx.tolist()
    assert torch.equal(twin.dirichlets, sel.dirichlets) and torch.equal(twin.pi_hat, sel.pi_hat)
    assert torch.equal(twin.get_pbest(), sel.get_pbest())
    # the API path continues from a device loop: next pick equals the oracle's continuation
    random.seed(0)
    ora = coda_oracle.OracleSelector(preds)
    for i in idx.tolist():
        ora.add_label(i, int(labels[i]), 0.0)
    i_ref, q_ref = ora.get_next_item_to_label()
    i, qq = sel.get_next_item_to_label()
    assert i == i_ref and abs(qq - q_ref) < EIG_ATOL


def test_device_loop_flags_isclose_ties_and_exhaustion():
    """Two items with identical predictions tie exactly: the host-free loop takes the lower index (torch.argmax) and
    records that the reference would have drawn from random.choice (coda.py:306-311).  Running past the last
    unlabeled item is reported, not a fault."""
    from coda_b200.synth import synth
    preds, labels = synth(10, 400, 6, seed=8)
    random.seed(1)
    first, _ = coda_oracle.OracleSelector(preds).get_next_item_to_label()
    twin = (first + 137) % 400
    preds[:, twin] = preds[:, first]
    labels[twin] = labels[first]
    sel = _mk(preds, labels)
    sel.run_steps(2, labels)
    idx, q, tie = sel.history()
    assert idx[0] == min(first, twin) and tie[0] == 1
    tiny_p, tiny_l = synth(6, 5, 4, seed=3)
    s2 = _mk(tiny_p, tiny_l)
    s2.run_steps(5, tiny_l)
    assert sorted(s2.history()[0].tolist()) == [0, 1, 2, 3, 4]
    s2.run_steps(1, tiny_l)
    with pytest.raises(RuntimeError, match="no unlabeled"):
        s2.history()


def test_state_dict_resume_is_bit_exact():
    """Checkpoint after 3 labels, resume in a fresh selector (different shard count), continue: same picks, same bits."""
    g = load_golden("traj_small_h32_n3000_c10")
    preds, labels = golden_slab(g)
    random.seed(0)
    a = _mk(preds, labels)
    for k in range(3):
        i, q = a.get_next_item_to_label()
        a.add_label(i, int(labels[i]), q)
        a.get_best_model_prediction()
    sd = a.state_dict()
    b = _mk(preds, labels, shards=2)
    b.load_state_dict(sd)
    assert b.labeled_idxs == a.labeled_idxs and b.step == a.step and len(b.unlabeled_idxs) == len(a.unlabeled_idxs)
    assert torch.equal(b.dirichlets, a.dirichlets) and torch.equal(b.pi_hat, a.pi_hat) and torch.equal(b.get_pbest(), a.get_pbest())
    for k in range(3, int(g["steps"])):
        ia, qa = a.get_next_item_to_label()
        ib, qb = b.get_next_item_to_label()
        assert (ia, qa) == (ib, qb) and ia == int(g["idx"][k])
        assert torch.equal(a.eig, b.eig)
        a.add_label(ia, int(labels[ia]), qa)
        b.add_label(ib, int(labels[ib]), qb)
        assert int(a.get_best_model_prediction()) == int(b.get_best_model_prediction())
    assert torch.equal(b.dirichlets, a.dirichlets) and torch.equal(b.pi_hat_xi, a.pi_hat_xi)


def test_cold_paths_work_across_in_process_shards():
    """prefilter_n, q='uncertainty' and a tie list longer than the device buffer on a sharded slab (they gather the
    per-item vectors over this process's shards)."""
    from coda_b200.synth import synth
    preds, labels = synth(10, 600, 6, seed=8)
    random.seed(5)
    ora = coda_oracle.OracleSelector(preds, prefilter_n=50)
    i_ref, q_ref = ora.get_next_item_to_label()
    state_ref = random.getstate()
    random.seed(5)
    sel = _mk(preds, labels, prefilter_n=50, shards=2)
    i, q = sel.get_next_item_to_label()
    assert i == i_ref and abs(q - q_ref) < EIG_ATOL and random.getstate() == state_ref and sel.stochastic
    random.seed(4)
    ora = coda_oracle.OracleSelector(preds, q="uncertainty")
    random.seed(4)
    sel = _mk(preds, labels, q="uncertainty", shards=3)
    for _ in range(3):
        i_ref, q_ref = ora.get_next_item_to_label()
        i, qq = sel.get_next_item_to_label()
        assert i == i_ref and abs(qq - q_ref) < 1e-6
        ora.add_label(i, int(labels[i]), q_ref)
        sel.add_label(i, int(labels[i]), qq)
        assert int(ora.get_best_model_prediction()) == int(sel.get_best_model_prediction())
    p1, l1 = synth(1, 400, 2, seed=21)                            # H = 1: all 400 candidates tie
    random.seed(9)
    ora = coda_oracle.OracleSelector(p1)
    i_ref, _ = ora.get_next_item_to_label()
    st_ref = random.getstate()
    random.seed(9)
    sel = _mk(p1, l1, shards=2)
    i, _ = sel.get_next_item_to_label()
    assert sel.last_report["n_ties"] == 400
    if int(torch.isclose(ora.last_q, ora.last_q.max(), rtol=1e-8).sum()) == 400:
        assert i == i_ref and random.getstate() == st_ref


def test_single_process_multi_gpu_front_end():
    """`gpus=2`: one process, shard 1 copied to the second GPU over NVLink, exchanges through peer access."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    g = load_golden("traj_small_h32_n3000_c10")
    preds, labels = golden_slab(g)
    random.seed(0)
    sel = _mk(preds, labels, gpus=2)
    assert {e.dev.index for e in sel.engines} == {0, 1}
    for k in range(int(g["steps"])):
        i, q = sel.get_next_item_to_label()
        assert i == int(g["idx"][k])
        sel.add_label(i, int(labels[i]), q)
        assert int(sel.get_best_model_prediction()) == int(g["best_model"][k])
        np.testing.assert_allclose(sel.get_pbest().cpu().numpy()[0], g["pbest"][k], atol=1e-5)
    loop = _mk(preds, labels, gpus=2)
    loop.run_steps(int(g["steps"]), labels)
    assert loop.history()[0].tolist() == [int(i) for i in g["idx"]]


def test_sharded_file_dataset_feeds_the_selector(tmp_path):
    """an (H, N, C) `.pt` slab on disk read through mmap, one N-range per shard (no shard ever
    materialises the whole tensor), straight into the selector: the same run as from the in-memory tensor."""
    from coda_b200 import CODA, ShardedFileDataset, TensorDataset
    g = load_golden("traj_small_h32_n3000_c10")
    preds, labels = golden_slab(g)
    f = str(tmp_path / "task.pt")
    torch.save(preds, f)
    torch.save(labels, f.replace(".pt", "_labels.pt"))
    whole = ShardedFileDataset(f, "cuda:0")
    assert whole.preds.shape == preds.shape and whole.n_global == 3000
    sel = CODA(whole)
    sel.run_steps(int(g["steps"]), whole.labels)
    assert sel.history()[0].tolist() == [int(i) for i in g["idx"]]
    # one rank's range of a 3-way split: the loader hands over exactly that N-range, offsets in global item indices
    part = ShardedFileDataset(f, "cuda:0", rank=1, world=3)
    assert (part.n_offset, part.preds.shape[1], part.n_global) == (1000, 1000, 3000)
    assert torch.equal(part.preds.cpu(), preds[:, 1000:2000])
