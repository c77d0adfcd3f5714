"""The C-ABI library exports exactly the entry points include/tobias_b200.h declares; it loads without a GPU; and the
Python package refuses to run without the CUDA build (no CPU fallback)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tobias_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_documented_entry_points():
    names = declared_functions()
    for must in ("tb_init", "tb_destroy", "tb_last_error", "tb_genome_create", "tb_genome_upload", "tb_reads_upload", "tb_count_reads",
                 "tb_pileup", "tb_bias_counts", "tb_allreduce_i64", "tb_set_bias_model", "tb_bias_model_info", "tb_score_sequence", "tb_correct_run",
                 "tb_correct_download", "tb_score_regions", "tb_score_gather", "tb_rolling"):
        assert must in names


def test_cuda_library_exports_every_declared_symbol():
    from tobias_b200 import build, lib
    path = build.build()                       # nvcc cross-compiles without a GPU
    handle = ctypes.CDLL(path)                 # loads without a device (libcudart binds the driver lazily)
    for name in declared_functions():
        assert hasattr(handle, name), name
    assert set(lib.exported_symbols()) <= set(declared_functions())
    assert handle.tb_abi_version() == 1
    sass = subprocess.run(["cuobjdump", "-lelf", path], capture_output=True, text=True).stdout
    assert "sm_100a" in sass


def test_emulator_exports_the_same_abi(emu_lib):
    handle = ctypes.CDLL(emu_lib)
    for name in declared_functions():
        assert hasattr(handle, name), name


def test_no_cpu_fallback_without_a_device():
    """On a machine without a GPU the product path fails loudly at tb_init (run in a subprocess: the test process itself
    may have the emulator loaded)."""
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from tobias_b200 import lib\n"
            "import ctypes\n"
            "try:\n"
            "    lib.Context(0)\n"
            "    print('CONTEXT-OK')\n"
            "except lib.TobiasB200Error as e:\n"
            "    print('REFUSED', e)\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True).stdout
    import shutil
    has_gpu = shutil.which("nvidia-smi") is not None and subprocess.run(["nvidia-smi", "-L"], capture_output=True).returncode == 0
    if has_gpu:
        assert "CONTEXT-OK" in out
    else:
        assert "REFUSED" in out and "no CPU fallback" in out


@pytest.mark.gpu
def test_gpu_context_is_the_cuda_build():
    from tobias_b200 import lib
    with lib.Context(0) as c:
        info = c.device_info()
        assert not info["is_emulator"] and info["sm_count"] >= 100
        assert os.path.abspath(lib.loaded_path()) == os.path.abspath(lib.DEFAULT_LIB)
