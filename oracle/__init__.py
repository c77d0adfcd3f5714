"""
oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU checkers for the TOBIAS per-base hot path:
  * oracle/_ref/      the UNMODIFIED reference compiled to binaries by oracle/build_ref.py
                      (git-ignored; travels to the GPU box)
  * oracle/restate.py + oracle/restate_c.c
                      an independent CPU restatement of the reference's algorithm (numpy + C),
                      pinned against oracle/_ref and the committed fixtures in tests/golden/

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import anything from here.  The product (tobias_b200/) never does.
"""
