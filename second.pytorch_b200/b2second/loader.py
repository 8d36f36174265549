"""Locate / import the two ``spconv`` implementations that live in this repository.

* product  : ``second.pytorch_b200/spconv``  (CUDA, libb2second.so)  -> ``import spconv``
* oracle   : ``oracle/spconv_cpu/spconv``    (CPU restatement, TEST INFRASTRUCTURE ONLY)

Both are packages called ``spconv`` (that is the reference's plugin name), so a process that needs
both (GPU parity tests, bench.py's cpu_baseline) loads the oracle under an alias.
"""
import importlib
import importlib.util
import os
import sys

REPO_ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
PRODUCT_DIR = os.path.join(REPO_ROOT, "second.pytorch_b200")
ORACLE_DIR = os.path.join(REPO_ROOT, "oracle", "spconv_cpu")


def product_spconv():
    """import the CUDA drop-in as ``spconv`` (raises if another ``spconv`` is already imported)."""
    if PRODUCT_DIR not in sys.path:
        sys.path.insert(0, PRODUCT_DIR)
    mod = importlib.import_module("spconv")
    if getattr(mod, "__oracle__", False):
        raise RuntimeError("the oracle spconv is imported as `spconv` in this process")
    return mod


def oracle_spconv(alias="b2s_oracle_spconv"):
    """load the CPU oracle package under ``alias`` (tests / bench cpu_baseline only)."""
    if alias in sys.modules:
        return sys.modules[alias]
    pkg_dir = os.path.join(ORACLE_DIR, "spconv")
    spec = importlib.util.spec_from_file_location(alias, os.path.join(pkg_dir, "__init__.py"),
                                                  submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[alias] = mod
    spec.loader.exec_module(mod)
    return mod
