"""
Builds tobias_b200/csrc/libtobias_b200.so in-tree with nvcc for sm_100a (B200).  nvcc cross-compiles without a
GPU, so this runs in the CPU-only build container; the .so is git-ignored but travels to the GPU box.

    python -m tobias_b200.build [--force] [--verbose]
"""
import glob
import os
import subprocess
import sys

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.path.join(CSRC, "libtobias_b200.so")
SOURCES = ["tb_api.cu", "tb_pileup.cu", "tb_counts.cu", "tb_score.cu", "tb_correct.cu", "tb_footprint.cu", "tb_nccl.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-fmad=false",                  # parity: the reference's C / numpy arithmetic has no fused multiply-add
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def is_current():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    deps = glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        [os.path.join(CSRC, "..", "..", "include", "tobias_b200.h")]
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False, verbose=False, ptxas_info=False, extra_flags=(), out=None):
    """extra_flags / out: build a tuning variant next to the product library (used by kernel experiments only)."""
    global LIB
    if out is None and is_current() and not force:
        return LIB
    lib_path = out or LIB
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        if out is not None:
            obj = obj[:-2] + "." + os.path.basename(out) + ".o"
        cmd = [_nvcc()] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if ptxas_info else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        log, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed on %s:\n%s\n" % (src, log))
        elif (verbose or ptxas_info) and log.strip():
            print(log)
    if failed:
        raise RuntimeError("nvcc compilation failed")
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib_path] + objs + ["-lcudart", "-ldl"]
    subprocess.check_call(cmd)
    if out is not None:
        for o in objs:
            os.remove(o)
    return lib_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, ptxas_info="--ptxas" in sys.argv))
