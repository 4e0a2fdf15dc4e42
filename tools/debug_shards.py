"""Debug helper: in-process sharded API path vs one shard, step by step."""
import os, random, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import golden_slab, load_golden
from coda_b200 import CODA, TensorDataset

g = load_golden("traj_small_h32_n3000_c10")
preds, labels = golden_slab(g)
dev = torch.device("cuda:0")
mk = lambda **kw: CODA(TensorDataset(preds.to(dev), labels.to(dev)), **kw)
random.seed(0); one = mk()
random.seed(0); many = mk(shards=int(sys.argv[1]) if len(sys.argv) > 1 else 2)
print("init D eq", torch.equal(one.dirichlets, many.dirichlets), "pi eq", torch.equal(one.pi_hat, many.pi_hat))
for k in range(6):
    i1, q1 = one.get_next_item_to_label(); i2, q2 = many.get_next_item_to_label()
    print(k, "pick", i1, i2, q1, q2, "eig eq", torch.equal(one.eig, many.eig))
    t = int(labels[i1])
    one.add_label(i1, t, q1); many.add_label(i1, t, q1)
    one._sync(); many._sync(); torch.cuda.synchronize()
    U1 = one.engine.U; Um = torch.cat([e.U for e in many.engines], 0)
    print("   t", t, "D eq", torch.equal(one.dirichlets, many.dirichlets), "U eq", torch.equal(U1, Um),
          "bad rows", int((U1 != Um).any(1).sum()), "pi eq", torch.equal(one.pi_hat, many.pi_hat))
    print("   jvec eq", [torch.equal(e.jvec, one.engine.jvec) for e in many.engines], "hdr", one.engine.terms[:2].tolist(),
          [e.terms[:2].tolist() for e in many.engines], "sel", one.engine.sel.tolist(), [e.sel.tolist() for e in many.engines])
    ps = sum(e.pisum.cpu() for e in many.engines)
    print("   pisum sum eq", torch.equal(ps, one.engine.pisum.cpu()), "epochs", [e._mailbox.epoch[:4].tolist() for e in many.engines])
    if not torch.equal(U1, Um):
        bad = (U1 != Um).any(1).nonzero().flatten()
        print("   first bad rows", bad[:10].tolist(), "cols", (U1 != Um).any(0).nonzero().flatten().tolist())
