"""N > 1 on hardware: two NCCL ranks over the same workload produce the same results as one rank.  Every contig is seeded by its index in
the whole workload (synth_torch.TorchDataset(universe=...)), so the 2-rank job processes exactly the data of the 1-rank job; the bench
line's order-independent checksums over the count tensors (after the NCCL all-reduce) and over the five output tracks must agree.
Needs two GPUs (skipped otherwise; run with `gpurun --gpus 2`)."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _bench(n):
    common = [os.path.join(ROOT, "bench.py"), "--gpus", str(n), "--workload", "small", "--steps", "1", "--warmup", "3", "--no-cpu-baseline"]
    if n == 1:
        cmd = [sys.executable] + common
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
               "--master-port", str(_free_port())] + common
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])


@pytest.mark.gpu
def test_two_nccl_ranks_equal_one_rank():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    one, two = _bench(1), _bench(2)
    assert one["n_gpus"] == 1 and two["n_gpus"] == 2
    assert one["config"]["out_bp_total"] == two["config"]["out_bp_total"]
    assert one["checksum"]["count_tensors"] == two["checksum"]["count_tensors"] and one["checksum"]["no_reads"] == two["checksum"]["no_reads"]
    assert one["checksum"]["tracks"] == two["checksum"]["tracks"], (one["checksum"], two["checksum"])
    assert one["correction_factor"] == two["correction_factor"]
