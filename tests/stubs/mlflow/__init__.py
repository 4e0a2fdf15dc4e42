"""Recording stand-in for the part of the MLflow API the reference driver uses (main.py:15-17, 102-103, 132-164).

mlflow is not installed in the build / GPU images.  Every call is appended as one JSON line to the file named by
MLFLOW_STUB_LOG, so a test can check that `main.py` logged the parameters, the nested runs and the per-step
`regret` / `cumulative regret` metrics.  `log_metric` also records the driver loop's own local variables
(chosen_idx, true_class, best_model_idx_pred of main.py:91-94) by looking at the calling frame -- the reference
does not log them, and this keeps the driver itself unmodified.  Test infrastructure only.
"""
import contextlib
import json
import os
import sys

_LOG = os.environ.get("MLFLOW_STUB_LOG")


def _rec(kind, **kw):
    if not _LOG:
        return
    with open(_LOG, "a") as f:
        f.write(json.dumps(dict(kind=kind, **kw), default=str) + "\n")


def set_tracking_uri(uri):
    _rec("set_tracking_uri", uri=uri)


def set_experiment(name):
    _rec("set_experiment", name=name)


class _NoRuns:
    """search_runs() result for a fresh store: len() == 0 (main.py:139-146 only indexes it when non-empty)."""
    columns = ()

    def __len__(self):
        return 0


def search_runs(*a, **k):
    return _NoRuns()


@contextlib.contextmanager
def start_run(run_id=None, run_name=None, nested=False, **k):
    _rec("start_run", run_name=run_name, nested=bool(nested))
    yield None
    _rec("end_run", run_name=run_name)


def log_params(params):
    _rec("log_params", params={k: (v if isinstance(v, (int, float, str, bool, type(None))) else str(v)) for k, v in dict(params).items()})


def log_param(key, value):
    _rec("log_param", key=key, value=value if isinstance(value, (int, float, str, bool, type(None))) else str(value))


def log_metric(key, value, step=None):
    loc = sys._getframe(1).f_locals
    extra = {}
    for name in ("chosen_idx", "true_class", "best_model_idx_pred"):
        if name in loc:
            try:
                extra[name] = int(loc[name])
            except Exception:
                pass
    _rec("log_metric", key=key, value=float(value), step=step, **extra)
