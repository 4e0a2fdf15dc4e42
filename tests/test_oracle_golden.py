"""Pin the CPU oracle (oracle/coda_oracle.py) against outputs of the reference itself
(tests/golden/*.npz, produced by tests/golden/make_golden.py in the build container)."""
import random

import numpy as np
import pytest
import torch

from helpers import coda_oracle, golden_names, golden_slab, load_golden, GOLDEN


def test_quadrature_known_answers():
    z = np.load(f"{GOLDEN}/quadrature_kat.npz")
    got = coda_oracle.pbest_rows(torch.from_numpy(z["alpha"]), torch.from_numpy(z["beta"]))
    np.testing.assert_allclose(got.numpy(), z["pbest"], rtol=2e-6, atol=1e-9)
    assert np.array_equal(coda_oracle.quad_grid().numpy(), z["grid"])   # trap T1: same fp32 grid bits
    np.testing.assert_allclose(got.sum(-1).numpy(), 1.0, atol=1e-5)


@pytest.mark.parametrize("name", golden_names())
def test_trajectory_matches_reference(name):
    g = load_golden(name)
    if int(g["N"]) > 5000 or int(g["H"]) * int(g["N"]) * int(g["C"]) > 2e7:
        pytest.skip("large golden is for the GPU parity test; oracle replay would take minutes")
    preds, labels = golden_slab(g)
    random.seed(0)
    sel = coda_oracle.OracleSelector(preds, **g["ctor"])
    np.testing.assert_allclose(sel.dirichlets.numpy(), g["init_dirichlets"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(sel.pi_hat.numpy(), g["init_pi_hat"], rtol=1e-6)
    np.testing.assert_allclose(sel.pi_hat_xi.numpy(), g["init_pi_hat_xi"], rtol=1e-5, atol=1e-8)
    assert int(sel.get_best_model_prediction()) == int(g["init_best_model"])
    np.testing.assert_allclose(sel.get_pbest().numpy(), g["init_pbest"], rtol=1e-5, atol=1e-8)
    for k in range(int(g["steps"])):
        idx, q = sel.get_next_item_to_label()
        ref_eig = g["eig"][k]
        cand = np.asarray(sel.last_cand)
        assert len(cand) == int(g["n_cand"][k])
        assert np.all(np.isfinite(ref_eig[cand])) and np.isnan(np.delete(ref_eig, cand)).all()
        # same arithmetic order as the reference => agreement to fp32 rounding noise
        np.testing.assert_allclose(sel.last_q.numpy(), ref_eig[cand], atol=2e-6)
        assert idx == int(g["idx"][k]), (k, idx, int(g["idx"][k]))
        assert abs(q - float(g["q"][k])) < 2e-6
        t = int(labels[idx])
        sel.add_label(idx, t, q)
        assert int(sel.get_best_model_prediction()) == int(g["best_model"][k])
        np.testing.assert_allclose(sel.get_pbest().numpy()[0], g["pbest"][k], rtol=1e-5, atol=1e-8)
        np.testing.assert_allclose(sel.pi_hat.numpy(), g["pi_hat"][k], rtol=1e-6)
        np.testing.assert_array_equal(sel.dirichlets[:, t].numpy(), g["dir_row"][k])  # update is one fp32 add
    np.testing.assert_allclose(sel.dirichlets.numpy(), g["final_dirichlets"], rtol=1e-6, atol=1e-7)
    assert sel.step == int(g["steps"]) + 1
    assert int(sel.stochastic) == int(g["stochastic"])


def test_error_behaviour_matches_reference():
    """coda.py:297 NotImplementedError(q); coda.py:323 list.remove ValueError; util.py:20-25 RuntimeError."""
    g = load_golden("traj_tiny_h8_n300_c5")
    preds, labels = golden_slab(g)
    sel = coda_oracle.OracleSelector(preds, q="bogus")
    with pytest.raises(NotImplementedError):
        sel.get_next_item_to_label()
    sel = coda_oracle.OracleSelector(preds)
    sel.add_label(3, 1, 0.0)
    with pytest.raises(ValueError):
        sel.add_label(3, 1, 0.0)
    bad = torch.tensor([[float("nan"), 1.0]])
    with pytest.raises(RuntimeError, match="NUMERIC ERROR"):
        coda_oracle.pbest_rows(bad, torch.ones(1, 2))


@pytest.mark.parametrize("q", ["iid", "uncertainty"])
def test_oracle_ablation_acquisitions_vs_live_reference(q):
    """No golden for the ablation acquisitions: compare with the reference itself where it is mounted
    (build container only; skipped on the GPU box)."""
    import os
    import sys
    import types
    ref = os.environ.get("CODA_REFERENCE_PATH", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "coda")):
        pytest.skip("reference checkout not available")
    from coda_b200.synth import synth
    saved = {k: v for k, v in sys.modules.items() if k == "coda" or k.startswith("coda.")}
    for k in saved:
        del sys.modules[k]
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, ref)
    try:
        import coda.coda as ref_coda
        assert ref_coda.__file__.startswith(ref)
        preds, labels = synth(12, 500, 6, seed=17)

        class DS:
            pass
        ds = DS()
        ds.preds, ds.labels, ds.device = preds, labels, preds.device
        random.seed(4)
        r = ref_coda.CODA(ds, q=q)
        random.seed(4)
        o = coda_oracle.OracleSelector(preds, q=q)
        for _ in range(4):
            st = random.getstate()
            ir, qr = r.get_next_item_to_label()
            after = random.getstate()
            random.setstate(st)
            io, qo = o.get_next_item_to_label()
            assert (io, random.getstate()) == (ir, after) and abs(qo - qr) < 1e-7
            r.add_label(ir, int(labels[ir]), qr)
            o.add_label(io, int(labels[io]), qo)
            assert int(r.get_best_model_prediction()) == int(o.get_best_model_prediction())
    finally:
        sys.path.remove(ref)
        for k in [k for k in sys.modules if k == "coda" or k.startswith("coda.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_oracle_prefilter_subsample_vs_live_reference():
    """coda.py:221-223 (--prefilter-n): random.sample over the candidate list, then the tie rule on the subsample --
    checked against the reference itself where it is mounted (same RNG consumption, same pick)."""
    import os
    import sys
    import types
    ref = os.environ.get("CODA_REFERENCE_PATH", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "coda")):
        pytest.skip("reference checkout not available")
    from coda_b200.synth import synth
    saved = {k: v for k, v in sys.modules.items() if k == "coda" or k.startswith("coda.")}
    for k in saved:
        del sys.modules[k]
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, ref)
    try:
        import coda.coda as ref_coda
        ref_coda.tqdm = lambda it, *a, **k: it
        preds, labels = synth(10, 600, 6, seed=8)

        class DS:
            pass
        ds = DS()
        ds.preds, ds.labels, ds.device = preds, labels, preds.device
        random.seed(5)
        r = ref_coda.CODA(ds, prefilter_n=50)
        ir, qr = r.get_next_item_to_label()
        after = random.getstate()
        random.seed(5)
        o = coda_oracle.OracleSelector(preds, prefilter_n=50)
        io, qo = o.get_next_item_to_label()
        assert (io, random.getstate()) == (ir, after) and abs(qo - qr) < 2e-6 and o.stochastic and r.stochastic
    finally:
        sys.path.remove(ref)
        for k in [k for k in sys.modules if k == "coda" or k.startswith("coda.")]:
            del sys.modules[k]
        sys.modules.update(saved)
