"""
TEST INFRASTRUCTURE ONLY -- compiles the CUDA sources of tobias_b200/csrc with g++ against the kernel emulator
(tests/emu/cuda_emu.h, -DTB_EMU) into tests/emu/_build/libtobias_b200_emu.so so that the kernels' logic can be
exercised through the same C-ABI by the CPU test-suite (`pytest -m "not gpu"`) in a container without a GPU.
The product never loads this library.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "tobias_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libtobias_b200_emu.so")
SOURCES = ["tb_api.cu", "tb_pileup.cu", "tb_counts.cu", "tb_score.cu", "tb_correct.cu", "tb_footprint.cu", "tb_nccl.cu"]


def is_current():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    deps = glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        [os.path.join(HERE, "cuda_emu.h"), os.path.join(ROOT, "include", "tobias_b200.h")]
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False):
    if is_current() and not force:
        return LIB
    os.makedirs(OUT, exist_ok=True)
    procs, objs = [], []
    for src in SOURCES:
        obj = os.path.join(OUT, src.replace(".cu", ".o"))
        cmd = ["g++", "-std=c++17", "-O2", "-g", "-fPIC", "-DTB_EMU", "-DTB_FIRST_GEN_FITS", "-mfma", "-ffp-contract=off", "-I", HERE, "-x", "c++",
               "-Wno-unused-value", "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write("g++ (emulator build) failed on %s:\n%s\n" % (src, out))
    if failed:
        raise RuntimeError("emulator build failed")
    subprocess.check_call(["g++", "-shared", "-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
