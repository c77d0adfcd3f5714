"""
One process per GPU.  The hot path shards by contig (regions never span contigs, halos are <= 162 bp and are read from
the rank's own contig data), so the only exchange steps are:
  * one all-reduce (sum) of the packed bias-training count tensors + normalisation counters   -- AtacBias.join,
    tobias/tools/atacorrect_functions.py:54-60; driver loop tobias/tools/atacorrect.py:285-300,324-326
  * one all-reduce (sum) of the verification PWM counts                                       -- atacorrect.py:416-424
  * gathering the per-contig tracks on rank 0 for the single bigWig writer per track           -- utilities.py:129-232
torch.distributed is the plumbing: backend "nccl" on GPUs (NVLink / NVSwitch), "gloo" in the CPU tests.
"""
import os

import numpy as np


class Comm:
    """Thin wrapper over an (optional) initialised torch.distributed process group."""

    def __init__(self):
        self.rank, self.world, self.local_rank = 0, 1, int(os.environ.get("LOCAL_RANK", "0"))
        self._dist = None
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self._dist = dist
                self.rank, self.world = dist.get_rank(), dist.get_world_size()
        except ImportError:
            pass

    @staticmethod
    def init_from_env(backend=None):
        """Initialise the default process group from torchrun's environment (RANK / WORLD_SIZE / MASTER_*)."""
        import torch
        import torch.distributed as dist
        if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not dist.is_initialized():
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            kw = {}
            if backend == "nccl":
                torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
                kw["device_id"] = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
            dist.init_process_group(backend, **kw)
        return Comm()

    def _device(self):
        import torch
        if self._dist.get_backend() == "nccl":
            return torch.device("cuda", self.local_rank)
        return torch.device("cpu")

    def allreduce_sum(self, buf):
        """In-place sum of an int64 / float64 numpy array over all ranks (exact for the integer count tensors)."""
        if self._dist is None or self.world == 1:
            return buf
        import torch
        t = torch.from_numpy(np.ascontiguousarray(buf)).to(self._device())
        self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM)
        buf[...] = t.cpu().numpy().reshape(buf.shape)
        return buf

    def barrier(self):
        if self._dist is not None and self.world > 1:
            self._dist.barrier()

    def gather_arrays(self, arr, dst=0):
        """Rank `dst` receives the list of every rank's 1-D array (variable lengths); other ranks get None."""
        if self._dist is None or self.world == 1:
            return [arr]
        import torch
        dev = self._device()
        n = torch.tensor([arr.shape[0]], dtype=torch.int64, device=dev)
        sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(self.world)]
        self._dist.all_gather(sizes, n)
        sizes = [int(s.item()) for s in sizes]
        if self.rank == dst:
            out = []
            for r in range(self.world):
                if r == dst:
                    out.append(arr)
                else:
                    buf = torch.empty(sizes[r], dtype=torch.from_numpy(arr[:0]).dtype, device=dev)
                    if sizes[r]:
                        self._dist.recv(buf, src=r)
                    out.append(buf.cpu().numpy())
            return out
        if arr.shape[0]:
            self._dist.send(torch.from_numpy(np.ascontiguousarray(arr)).to(dev), dst=dst)
        return None


def bind_to_device_numa(device_index):
    """Restricts this process to the CPUs next to its GPU (NVML's CPU affinity of the device), so that the pinned buffers it allocates
    afterwards and the threads that fill them sit on the GPU's NUMA node: with one process per GPU the host-side copies of the eight
    ranks then do not cross the socket interconnect.  Returns the CPU set, or None when NVML / the topology gives nothing to do."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(device_index)
        bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        near = {i for i in range(ncpu) if (int(mask[i // 64]) >> (i % 64)) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = near & allowed
        if not cpus or cpus == allowed:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def shard_contigs(lengths, world, weights=None):
    """Greedy longest-first bin packing of contigs onto ranks: dict name -> rank.  `weights` (name -> work estimate,
    e.g. reads + alpha * bp) default to the contig lengths."""
    w = weights or lengths
    loads = [0.0] * world
    owner = {}
    for name in sorted(lengths, key=lambda n: (-w[n], n)):
        r = min(range(world), key=lambda i: (loads[i], i))
        owner[name] = r
        loads[r] += w[name]
    return owner
