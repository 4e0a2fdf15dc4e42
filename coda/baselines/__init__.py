"""The competing selectors (IID, Uncertainty, ActiveTesting, VMA, ModelPicker) are outside the scope of
this package.  ``main.py`` imports their names unconditionally
(main.py:10), so they resolve here: to the reference's own classes when a reference checkout is
reachable through ``CODA_REFERENCE_PATH``, else to placeholders that raise on construction."""
import importlib.util
import os
import sys

_NAMES = {"IID": "iid", "ActiveTesting": "activetesting", "VMA": "vma", "ModelPicker": "modelpicker",
          "Uncertainty": "uncertainty"}


def _placeholder(name):
    class _Missing:
        def __init__(self, *a, **k):
            raise NotImplementedError(
                f"coda.baselines.{name} is not part of coda_b200; set CODA_REFERENCE_PATH to a checkout of "
                "justinkay/coda to use the reference implementation")
    _Missing.__name__ = name
    return _Missing


def _load_reference(path):
    out = {}
    base = os.path.join(path, "coda", "baselines")
    for cls, mod in _NAMES.items():
        f = os.path.join(base, mod + ".py")
        if not os.path.exists(f):
            return None
        spec = importlib.util.spec_from_file_location(f"coda.baselines.{mod}", f)
        m = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = m
        spec.loader.exec_module(m)
        out[cls] = getattr(m, cls)
    return out


_ref = os.environ.get("CODA_REFERENCE_PATH")
_loaded = None
if _ref and os.path.isdir(_ref):
    try:
        _loaded = _load_reference(_ref)
    except Exception:  # pragma: no cover - the reference needs matplotlib etc.
        _loaded = None
for _cls in _NAMES:
    globals()[_cls] = _loaded[_cls] if _loaded else _placeholder(_cls)
