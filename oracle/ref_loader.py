"""
TEST INFRASTRUCTURE ONLY -- imports the compiled, unmodified reference from oracle/_ref.

`load()` returns a namespace with the reference's hot-path callables.  pysam / pyBigWig / matplotlib
are replaced by the shims in oracle/shims/ when (and only when) the real packages are not importable.
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
SHIMS = os.path.join(HERE, "shims")
_loaded = None


def available():
    from . import build_ref
    return build_ref.is_built() or os.path.isdir("/root/reference/tobias")


def _install_shims():
    for name in ("pysam", "pyBigWig", "matplotlib"):
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
        except ImportError:
            if SHIMS not in sys.path:
                sys.path.insert(0, SHIMS)
            importlib.import_module(name)
    for sub in ("matplotlib.pyplot", "matplotlib.ticker", "matplotlib.backends.backend_pdf"):
        importlib.import_module(sub)


def load():
    """Build (if the reference tree is present and _ref is missing) and import the compiled reference."""
    global _loaded
    if _loaded is not None:
        return _loaded
    from . import build_ref
    build_ref.build()          # no-op when oracle/_ref is complete (restores .pyc files a file sync may have dropped)
    _install_shims()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    ns = types.SimpleNamespace()
    ns.signals = importlib.import_module("tobias.utils.signals")
    ns.sequences = importlib.import_module("tobias.utils.sequences")
    ns.ngs = importlib.import_module("tobias.utils.ngs")
    ns.regions = importlib.import_module("tobias.utils.regions")
    ns.utilities = importlib.import_module("tobias.utils.utilities")
    ns.logger = importlib.import_module("tobias.utils.logger")
    ns.atacorrect_functions = importlib.import_module("tobias.tools.atacorrect_functions")
    ns.score_bigwig = importlib.import_module("tobias.tools.score_bigwig")
    ns.parsers = importlib.import_module("tobias.parsers")
    _loaded = ns
    return ns


def load_driver():
    """The reference's ATACorrect driver (tobias/tools/atacorrect.py:53), compiled."""
    load()
    return importlib.import_module("tobias.tools.atacorrect")
