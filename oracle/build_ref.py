#!/usr/bin/env python
"""
TEST INFRASTRUCTURE ONLY -- builds `oracle/_ref/`: the UNMODIFIED reference (loosolab/TOBIAS 0.17.3),
compiled from the sources where they lie under /root/reference into binary extension modules.

What is compiled: the three .pyx files go Cython -> C -> .so exactly as the reference's own
setup.py:25-28 does; the pure-python modules of the hot path are byte-compiled to sourceless
.pyc (Cython rejects utilities.py for mixed tabs/spaces), so that only binaries -- never
reference sources -- land in the repo tree:

    tobias/utils/signals.pyx      -> fast_rolling_math / tobias_footprint_array / FOS_score
    tobias/utils/sequences.pyx    -> SequenceMatrix / NucleotideMatrix / DiNucleotideMatrix / GenomicSequence
    tobias/utils/ngs.pyx          -> OneRead / ReadList
    tobias/utils/{regions,utilities,logger}.py, tobias/parsers.py
    tobias/tools/{atacorrect_functions,score_bigwig,atacorrect}.py

The generated C files live in a temporary directory outside the repo and are deleted.  The
`__init__.py` files written into oracle/_ref/tobias are generated here (they only carry the
version string parsed from the reference).  oracle/_ref/ is git-ignored but NOT gpurun-ignored,
so the binaries travel to the GPU box; nothing on the GPU box reads /root/reference.

pysam / pyBigWig / matplotlib are absent from this image: `oracle/ref_loader.py` installs the
shim modules from `oracle/shims/` into sys.modules before importing the compiled reference.

Usage:  python oracle/build_ref.py [--reference /root/reference] [--force]
"""
import argparse
import glob
import os
import re
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

MODULES = [
    "tobias/utils/signals.pyx",
    "tobias/utils/sequences.pyx",
    "tobias/utils/ngs.pyx",
    "tobias/utils/regions.py",
    "tobias/utils/utilities.py",
    "tobias/utils/logger.py",
    "tobias/parsers.py",
    "tobias/tools/atacorrect_functions.py",
    "tobias/tools/score_bigwig.py",
    "tobias/tools/atacorrect.py",
]


def is_built():
    for m in MODULES:
        stem, ext = os.path.splitext(m)
        if ext == ".pyx" and not glob.glob(os.path.join(OUT, stem + ".*.so")):
            return False
        if ext == ".py" and not (os.path.exists(os.path.join(OUT, stem + ".pyc")) or os.path.exists(os.path.join(OUT, stem + ".pycbin"))):
            return False
    return os.path.exists(os.path.join(OUT, "tobias", "__init__.py"))


def restore_pyc():
    """Sourceless imports need the .pyc suffix; restore them from the .pycbin twins if a copy step dropped '*.pyc'."""
    for m in MODULES:
        stem, ext = os.path.splitext(m)
        pyc = os.path.join(OUT, stem + ".pyc")
        if ext == ".py" and not os.path.exists(pyc) and os.path.exists(pyc + "bin"):
            shutil.copyfile(pyc + "bin", pyc)


def build(reference="/root/reference", force=False, quiet=True):
    if is_built() and not force:
        restore_pyc()
        return OUT
    if not os.path.isdir(os.path.join(reference, "tobias")):
        raise FileNotFoundError("reference tree not found at %s (oracle/_ref must be prebuilt)" % reference)

    import numpy
    from setuptools import Extension
    from setuptools.dist import Distribution
    from Cython.Build import cythonize

    tmp = tempfile.mkdtemp(prefix="tobias_ref_build_")
    try:
        # transient copy of the sources to compile (outside the repo; deleted below)
        for m in MODULES:
            dst = os.path.join(tmp, m)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(os.path.join(reference, m), dst)
        exts = []
        for m in [m for m in MODULES if m.endswith(".pyx")]:
            name = os.path.splitext(m)[0].replace("/", ".")
            exts.append(Extension(name, [os.path.join(tmp, m)], include_dirs=[numpy.get_include()],
                                  extra_compile_args=["-O2", "-w"],
                                  define_macros=[("NPY_NO_DEPRECATED_API", "NPY_1_7_API_VERSION")]))
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            ext_modules = cythonize(exts, quiet=quiet, language_level=3,
                                    compiler_directives={"binding": True})
            dist = Distribution({"name": "tobias_ref", "ext_modules": ext_modules})
            cmd = dist.get_command_obj("build_ext")
            cmd.build_lib = os.path.join(tmp, "_lib")
            cmd.build_temp = os.path.join(tmp, "_tmp")
            cmd.parallel = os.cpu_count() or 1
            cmd.ensure_finalized()
            if quiet:
                devnull = os.open(os.devnull, os.O_WRONLY)
                saved = os.dup(1), os.dup(2)
                os.dup2(devnull, 1)
                os.dup2(devnull, 2)
            try:
                cmd.run()
            finally:
                if quiet:
                    os.dup2(saved[0], 1)
                    os.dup2(saved[1], 2)
                    os.close(devnull)
        finally:
            os.chdir(cwd)

        if os.path.isdir(OUT):
            shutil.rmtree(OUT)
        for so in glob.glob(os.path.join(tmp, "_lib", "**", "*.so"), recursive=True):
            rel = os.path.relpath(so, os.path.join(tmp, "_lib"))
            dst = os.path.join(OUT, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(so, dst)
        import py_compile
        for m in [m for m in MODULES if m.endswith(".py")]:
            dst = os.path.join(OUT, os.path.splitext(m)[0] + ".pyc")
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            py_compile.compile(os.path.join(tmp, m), cfile=dst, dfile=m, doraise=True, optimize=0)
            shutil.copyfile(dst, dst + "bin")      # same bytes under a name file-sync tools do not treat as a cache file
        # generated package markers (carry only the version string)
        version = "unknown"
        with open(os.path.join(reference, "tobias", "__init__.py")) as fh:
            mt = re.search(r'__version__\s*=\s*"([^"]+)"', fh.read())
            if mt:
                version = mt.group(1)
        for pkg in ("tobias", "tobias/utils", "tobias/tools"):
            with open(os.path.join(OUT, pkg, "__init__.py"), "w") as fh:
                fh.write("# generated by oracle/build_ref.py (compiled reference; binaries only)\n")
                if pkg == "tobias":
                    fh.write('__version__ = "%s"\n' % version)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return OUT


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    out = build(a.reference, a.force, quiet=not a.verbose)
    print("oracle/_ref built at", out)
