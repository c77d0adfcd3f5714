"""N > 1 path on CPU: two `gloo` ranks, contigs sharded over the ranks, count all-reduce, rank 0 gathers and writes.
The bigWigs must be identical to the single-process run (the count tensors are integers: exact in any order)."""
import os
import socket
import subprocess
import sys

import numpy as np

from tobias_b200 import synth
from tobias_b200.dist import shard_contigs
from tobias_b200.io.bigwig import BigWigReader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_contigs_balanced_and_deterministic():
    lengths = dict(synth.HG38_LENGTHS)
    for world in (1, 2, 4, 8):
        owner = shard_contigs(lengths, world)
        assert owner == shard_contigs(lengths, world)
        loads = [sum(lengths[c] for c in lengths if owner[c] == r) for r in range(world)]
        assert sum(loads) == sum(lengths.values())
        assert max(loads) <= 1.15 * sum(loads) / world + max(lengths.values()) * (world == 1)


def test_two_rank_gloo_equals_single_process(tmp_path, emu_lib):
    ds = synth.SynthDataset(31, {"chr1": 50000, "chr2": 42000, "chr3": 36000, "chrM": 5000}, n_peaks=12, n_records=36000)
    paths = ds.write(str(tmp_path), bam=True)
    worker = os.path.join(ROOT, "tests", "dist_worker.py")
    common = ["--bam", paths["bam"], "--genome", paths["fasta"], "--peaks", paths["peaks"], "--verbosity", "0", "--prefix", "x"]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    single, multi = str(tmp_path / "single"), str(tmp_path / "multi")
    subprocess.run([sys.executable, worker] + common + ["--outdir", single], check=True, env=env, timeout=900)
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                    "--master-port", str(_free_port()), worker] + common + ["--outdir", multi], check=True, env=env, timeout=900)
    for track in ("uncorrected", "bias", "expected", "corrected"):
        a, b = BigWigReader(os.path.join(single, "x_%s.bw" % track)), BigWigReader(os.path.join(multi, "x_%s.bw" % track))
        assert a.chroms == b.chroms
        for c, n in a.chroms.items():
            assert np.array_equal(a.values(c, 0, n), b.values(c, 0, n), equal_nan=True), (track, c)


def test_more_ranks_than_contigs(tmp_path, emu_lib):
    """A rank that owns no contig (3 ranks, 2 kept contigs) joins the all-reduces with zeros and the gathers with nothing: same bigWigs."""
    ds = synth.SynthDataset(32, {"chr1": 40000, "chr2": 30000, "chrM": 5000}, n_peaks=6, n_records=20000)
    paths = ds.write(str(tmp_path), bam=True)
    worker = os.path.join(ROOT, "tests", "dist_worker.py")
    common = ["--bam", paths["bam"], "--genome", paths["fasta"], "--peaks", paths["peaks"], "--verbosity", "0", "--prefix", "x"]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    single, multi = str(tmp_path / "single"), str(tmp_path / "multi")
    subprocess.run([sys.executable, worker] + common + ["--outdir", single], check=True, env=env, timeout=900)
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "3", "--master-addr", "127.0.0.1",
                    "--master-port", str(_free_port()), worker] + common + ["--outdir", multi], check=True, env=env, timeout=900)
    for track in ("uncorrected", "corrected"):
        a, b = BigWigReader(os.path.join(single, "x_%s.bw" % track)), BigWigReader(os.path.join(multi, "x_%s.bw" % track))
        for c, n in a.chroms.items():
            assert np.array_equal(a.values(c, 0, n), b.values(c, 0, n), equal_nan=True), (track, c)
