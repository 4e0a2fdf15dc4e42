"""BASELINE.json configs[0]: run the reference's REAL driver (`/root/reference/main.py`, unmodified, with the
reference's own `coda` package) on the cifar10_5592 stand-in on CPU and keep what it logged.

    python tests/golden/make_cfg1_golden.py [iters]     # needs /root/reference; ~70 s per iteration on 8 cores

The paper's tensors are not in the reference checkout (README.md:31, a 3.25 GB download), so the task file is the
synthetic stand-in SURVEY.md 8(d) names: synth(80, 10000, 10, seed 0) saved as cifar10_5592.pt / _labels.pt.
MLflow is not installed: `tests/stubs/mlflow` records every call (parameters, nested runs, the per-step `regret`
and `cumulative regret` metrics of main.py:102-103) plus the loop's chosen_idx / best_model_idx_pred.
Output: tests/golden/cfg1_main_py.json.
"""
import json
import os
import subprocess
import sys
import tempfile

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("CODA_REFERENCE_PATH", "/root/reference")
TASK = dict(H=80, N=10000, C=10, seed=0)


def write_task(d):
    from coda_b200.synth import synth
    preds, labels = synth(TASK["H"], TASK["N"], TASK["C"], TASK["seed"])
    torch.save(preds, os.path.join(d, "cifar10_5592.pt"))
    torch.save(labels, os.path.join(d, "cifar10_5592_labels.pt"))


def run_main(main_py, data_dir, iters, log, pythonpath, extra_env=None, safe_path=False):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join(pythonpath), MLFLOW_STUB_LOG=log)
    if safe_path:
        env["PYTHONSAFEPATH"] = "1"
    env.update(extra_env or {})
    cmd = [sys.executable, main_py, "--task", "cifar10_5592", "--data-dir", data_dir, "--method", "coda", "--seeds", "1",
           "--iters", str(iters)]
    return subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=data_dir)


def parse_log(path):
    recs = [json.loads(l) for l in open(path)]
    out = {"regret": [], "cumulative_regret": [], "chosen_idx": [], "true_class": [], "best_model": [], "runs": [],
           "params": None, "seed_params": []}
    for r in recs:
        if r["kind"] == "log_metric" and r["key"] == "regret":
            out["regret"].append(r["value"]); out["chosen_idx"].append(r.get("chosen_idx"))
            out["true_class"].append(r.get("true_class")); out["best_model"].append(r.get("best_model_idx_pred"))
            assert r["step"] == len(out["regret"])
        elif r["kind"] == "log_metric" and r["key"] == "cumulative regret":
            out["cumulative_regret"].append(r["value"])
        elif r["kind"] == "start_run":
            out["runs"].append([r["run_name"], r["nested"]])
        elif r["kind"] == "log_params":
            out["params"] = r["params"]
        elif r["kind"] == "log_param":
            out["seed_params"].append([r["key"], r["value"]])
    return out


if __name__ == "__main__":
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 25
    with tempfile.TemporaryDirectory() as d:
        write_task(d)
        log = os.path.join(d, "mlflow.jsonl")
        # sys.path[0] is the script's directory (the reference checkout): `coda` is the REFERENCE's package here
        r = run_main(os.path.join(REF, "main.py"), d, iters, log, [os.path.join(ROOT, "tests", "stubs")])
        if r.returncode != 0:
            sys.exit(r.stdout[-3000:] + r.stderr[-3000:])
        out = parse_log(log)
    out["task"] = TASK
    out["iters"] = iters
    out["stdout_head"] = r.stdout.splitlines()[:6]
    json.dump(out, open(os.path.join(HERE, "cfg1_main_py.json"), "w"), indent=1)
    print("cfg1 golden:", out["chosen_idx"], out["regret"])
