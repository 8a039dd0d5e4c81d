"""Load the compiled reference (oracle/_ref) next to this library.  TEST INFRASTRUCTURE ONLY.

The reference's python package does ``from splat_cuda import ...`` and ``from splat_py.x import ...``
with absolute names, so it is imported once with those names pointing where we want and then
re-registered under an alias:

  load_reference()            -> (splat_cuda_ref, splat_py_ref)   reference python on reference CUDA
  load_reference_on_b200()    -> splat_py_on_b200                 reference python on THIS library's
                                                                   `splat_cuda` (the drop-in proof)
"""
from __future__ import annotations

import importlib
import importlib.util
import sys
import sysconfig
from pathlib import Path

REF_DIR = Path(__file__).resolve().parent / "_ref"
_SUBMODULES = ("structs", "utils", "cuda_autograd_functions", "tile_culling", "rasterize", "depth")
_cache = {}


def ref_ext_path() -> Path:
    return REF_DIR / f"splat_cuda_ref{sysconfig.get_config_var('EXT_SUFFIX')}"


def reference_available() -> bool:
    return ref_ext_path().exists() and (REF_DIR / "splat_py" / "rasterize.py").exists()


def _load_ref_ext():
    if "ext" not in _cache:
        import torch  # noqa: F401

        spec = importlib.util.spec_from_file_location("splat_cuda_ref", str(ref_ext_path()))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        sys.modules["splat_cuda_ref"] = mod
        _cache["ext"] = mod
    return _cache["ext"]


def _import_reference_package(cuda_module, alias: str, submodules=_SUBMODULES, extra_path=None):
    """Import oracle/_ref/splat_py with `splat_cuda` bound to `cuda_module`; register it as `alias`."""
    saved = {k: v for k, v in sys.modules.items() if k == "splat_cuda" or k == "splat_py" or k.startswith("splat_py.")}
    for k in saved:
        del sys.modules[k]
    sys.modules["splat_cuda"] = cuda_module
    sys.path.insert(0, str(REF_DIR))
    if extra_path:
        sys.path.insert(0, str(extra_path))
    try:
        pkg = importlib.import_module("splat_py")
        for sub in submodules:
            importlib.import_module(f"splat_py.{sub}")
        loaded = {k: v for k, v in sys.modules.items() if k == "splat_py" or k.startswith("splat_py.")}
    finally:
        sys.path.remove(str(REF_DIR))
        if extra_path:
            sys.path.remove(str(extra_path))
        for k in [k for k in sys.modules if k == "splat_cuda" or k == "splat_py" or k.startswith("splat_py.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    for k, v in loaded.items():
        sys.modules[alias + k[len("splat_py"):]] = v
    return pkg


def load_reference():
    if "ref" not in _cache:
        if not reference_available():
            raise ImportError("oracle/_ref is not built (run `python oracle/build_ref.py` where /root/reference exists)")
        ext = _load_ref_ext()
        _cache["ref"] = (ext, _import_reference_package(ext, "splat_py_ref"))
    return _cache["ref"]


def load_reference_on_b200():
    if "on_b200" not in _cache:
        if not (REF_DIR / "splat_py" / "rasterize.py").exists():
            raise ImportError("oracle/_ref/splat_py is not installed")
        import gaussian_splatting_b200 as g

        _cache["on_b200"] = _import_reference_package(g.native(), "splat_py_on_b200")
    return _cache["on_b200"]


def load_reference_trainer():
    """The reference's trainer / optimizer-manager / config modules (splat_py/trainer.py etc.), imported with
    THIS library's `splat_cuda` underneath (the trainer's densification code does not touch CUDA kernels) and the
    stand-ins for the two packages the image lacks (tools/e2e/shims: torchmetrics SSIM, plotext).  Registered as
    `splat_py_trainer_ref.*`.  Used by tests/test_densify_gpu.py as the checker of the flat-buffer densification."""
    if "trainer" not in _cache:
        if not (REF_DIR / "splat_py" / "trainer.py").exists():
            raise ImportError("oracle/_ref/splat_py is not installed")
        import gaussian_splatting_b200 as g

        shims = Path(__file__).resolve().parent.parent / "tools" / "e2e" / "shims"
        _cache["trainer"] = _import_reference_package(
            g.native(), "splat_py_trainer_ref", submodules=_SUBMODULES + ("config", "optimizer_manager", "trainer"),
            extra_path=shims)
    return _cache["trainer"]
