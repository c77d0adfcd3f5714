"""Helper of tests/test_dist.py: one rank of a `gloo` world running the sharded ATACorrect driver on the kernel emulator."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))


def main():
    import build_emu
    from tobias_b200 import lib
    lib.load(build_emu.build())                      # CPU test: kernels run in the emulator
    import torch.distributed as dist
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        dist.init_process_group("gloo")
    from tobias_b200 import atacorrect, parsers
    args = parsers.add_atacorrect_arguments(argparse.ArgumentParser()).parse_args(sys.argv[1:])
    atacorrect.run_atacorrect(args)
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
