import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def emu_lib():
    """GPU-less kernel emulator build of the C-ABI library (test infrastructure, see tests/emu/cuda_emu.h)."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    return build_emu.build()


@pytest.fixture(scope="session", params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def backend(request):
    return request.param


@pytest.fixture(scope="session")
def ctx(backend, request):
    """A tobias_b200 Context: the CUDA library on a GPU box (-m gpu), the kernel emulator otherwise."""
    from tobias_b200 import lib
    if backend == "emu":
        c = lib.Context(0, request.getfixturevalue("emu_lib"))
        assert c.device_info()["is_emulator"]
    else:
        c = lib.Context(0)               # default path: tobias_b200/csrc/libtobias_b200.so, fails loudly if missing
        assert not c.device_info()["is_emulator"]
        assert os.path.abspath(lib.loaded_path()) == os.path.abspath(lib.DEFAULT_LIB)
    yield c
    c.close()


class GoldenData:
    def __init__(self):
        from tobias_b200.reads import ReadColumns
        z = np.load(os.path.join(GOLDEN, "golden_dataset.npz"))
        self.names = [str(x) for x in z["contig_names"]]
        self.lengths = [int(x) for x in z["contig_lengths"]]
        self.genome = [z["genome_" + n] for n in self.names]
        cols = {f: z["reads_" + f] for f in ("chrom_id", "ref_start", "ref_end", "left_soft", "right_soft", "query_length",
                                             "flag", "first_op", "last_op", "tlen")}
        self.reads = ReadColumns(self.names, self.lengths, **cols)
        self.peaks = z["peaks"]
        self.atac = np.load(os.path.join(GOLDEN, "golden_atacorrect.npz"))
        self.signals = np.load(os.path.join(GOLDEN, "golden_signals.npz"))
        self.score = np.load(os.path.join(GOLDEN, "golden_scorebigwig.npz"))


@pytest.fixture(scope="session")
def golden():
    return GoldenData()


@pytest.fixture(scope="session")
def loaded_ctx(ctx, golden):
    """Context with the golden genome and reads resident."""
    ctx.genome_load(golden.genome)
    ctx.reads_upload_columns(golden.reads)
    return ctx
