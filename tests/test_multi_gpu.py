"""N-axis sharding over >= 2 GPUs of one box (NCCL): identical picks, bit-identical pi_hat / dirichlets."""
import json
import os
import subprocess
import sys

import pytest
import torch

from helpers import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_equals_single(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29400 + world), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("MGPU_RESULT ")][-1]
    res = json.loads(line[len("MGPU_RESULT "):])
    for mode, o in res.items():
        assert o["same_on_all_ranks"], mode
        assert o["picks"] == o["picks_single"], (mode, o["picks"], o["picks_single"])
        assert o["pi_hat_equal"] and o["D_equal"], mode        # int64 fixed-point statistics: shard-count invariant bits
        assert o["pbest"] < 1e-6 and o["eig"] < 1e-7, (mode, o)
        assert o["loop_same_on_all_ranks"] and o["loop_picks"] == o["loop_picks_single"], (mode, o)
        assert o["loop_D_equal"] and o["loop_pi_hat_equal"], mode
