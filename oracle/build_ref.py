"""Recipe for ``baseline/_ref/``: the reference's own Python modules of the hot path, byte-compiled from the sources where
they lie (TEST / BASELINE INFRASTRUCTURE - never imported by the product package).

    python oracle/build_ref.py            # build container only (needs /root/reference); also run by __graft_entry__.build()

The reference is pure Python, so "compiling" it means ``py_compile``: every module below is compiled straight from
``/root/reference/onebit_asr/<name>.py`` into ``baseline/_ref/onebit_asr/<name>.bc`` (CPython bytecode, imported sourceless; not ``.pyc``: gpurun snapshots skip that suffix).  No reference
source text is copied into the repository; ``baseline/_ref/`` is git-ignored (it stays out of the history) but not
gpurun-ignored, so - like the built ``libonebit.so`` - it travels to the GPU box, where ``/root/reference`` does not exist.
``MANIFEST.json`` records the sha256 of each source file, the interpreter and the torch version it was built with.

Used by: tests (the unmodified ``conformer.ConformerASR`` on the CUDA layer), ``bench.py``'s ``cpu_baseline``
leg and ``--impl reference`` arm (the reference's own ``run_epoch`` on the host cores, ``kind: "reference"``).
"""
import hashlib
import json
import os
import py_compile
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "baseline", "_ref")
MODULES = ("quant", "conformer", "losses", "metrics", "train")     # onebit_asr/*.py on or around the hot path


def build_ref(reference: str = "/root/reference", out: str = OUT) -> bool:
    """Returns True when ``out`` holds a complete build (built now, or already there and the reference tree is absent)."""
    src_dir = os.path.join(reference, "onebit_asr")
    manifest_path = os.path.join(out, "MANIFEST.json")
    if not os.path.isdir(src_dir):
        return os.path.exists(manifest_path)
    pkg = os.path.join(out, "onebit_asr")
    os.makedirs(pkg, exist_ok=True)
    manifest = {"python": sys.version.split()[0], "modules": {}, "reference": reference}
    try:
        import torch
        manifest["torch"] = torch.__version__
    except Exception:  # noqa: BLE001
        pass
    for name in MODULES:
        src = os.path.join(src_dir, name + ".py")
        py_compile.compile(src, cfile=os.path.join(pkg, name + ".bc"), dfile=f"<reference>/onebit_asr/{name}.py", doraise=True)
        with open(src, "rb") as f:
            manifest["modules"][name] = hashlib.sha256(f.read()).hexdigest()
    with open(manifest_path, "w") as f:
        json.dump(manifest, f, indent=1)
    return True


if __name__ == "__main__":
    ok = build_ref(*(sys.argv[1:2]))
    print("baseline/_ref:", "ready" if ok else "NOT built (no reference tree and no earlier build)")
