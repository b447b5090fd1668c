"""Import the byte-compiled reference modules from ``baseline/_ref/`` (TEST / BASELINE INFRASTRUCTURE, see build_ref.py).

``load()`` executes the reference's own ``quant``, ``conformer``, ``losses`` (and on demand ``metrics`` / ``train``) under
private module names, so the pure reference and the reference-with-our-layer can live side by side in one process:

    ref = load()                                   # the unmodified reference, fp32 PyTorch layer
    swapped = load(quant_module=onebit_b200.quant) # the unmodified conformer.py, its flat ``from quant import QuantizedLinear``
                                                   # (conformer.py:12) resolved to the B200 layer - the drop-in seam of SURVEY 8b

Nothing here is imported by the product package.
"""
import importlib.machinery
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
_cache = {}


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, *p)) for p in (("MANIFEST.json",), ("onebit_asr", "quant.bc"), ("onebit_asr", "conformer.bc")))


def _exec(name: str, as_name: str, visible: dict):
    """Execute ``baseline/_ref/onebit_asr/<name>.bc`` as module ``as_name`` with ``visible`` temporarily in sys.modules."""
    path = os.path.join(REF_DIR, "onebit_asr", name + ".bc")
    loader = importlib.machinery.SourcelessFileLoader(as_name, path)
    spec = importlib.util.spec_from_loader(as_name, loader)
    mod = importlib.util.module_from_spec(spec)
    saved = {k: sys.modules.get(k) for k in visible}
    sys.modules.update(visible)
    try:
        loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def load(quant_module=None, with_train: bool = False):
    """Namespace with ``quant``, ``conformer``, ``losses`` (+ ``metrics``, ``train`` when ``with_train``).

    ``quant_module``: module that serves ``from quant import QuantizedLinear`` for conformer.py (None = the reference's own)."""
    if not available():
        raise FileNotFoundError("baseline/_ref is not built: run `python oracle/build_ref.py` where /root/reference exists")
    key = (id(quant_module), with_train)
    if key in _cache:
        return _cache[key]
    tag = "_obref%d" % len(_cache)
    ns = types.SimpleNamespace()
    ns.quant = _exec("quant", tag + ".quant", {}) if quant_module is None else quant_module
    ns.conformer = _exec("conformer", tag + ".conformer", {"quant": ns.quant})
    ns.losses = _exec("losses", tag + ".losses", {})
    if with_train:
        ns.metrics = _exec("metrics", tag + ".metrics", {})
        pkg = types.ModuleType("onebit_asr")
        pkg.__path__ = []
        visible = {"onebit_asr": pkg, "onebit_asr.conformer": ns.conformer, "onebit_asr.losses": ns.losses,
                   "onebit_asr.metrics": ns.metrics}
        for optional in ("wandb", "sentencepiece"):          # imported at the top of train.py, unused by run_epoch
            try:
                __import__(optional)
            except Exception:  # noqa: BLE001
                visible[optional] = types.ModuleType(optional)
        os.environ.setdefault("TQDM_DISABLE", "1")
        ns.train = _exec("train", tag + ".train", visible)
    _cache[key] = ns
    return ns
