"""`python -m tobias_b200 <Tool> ...` -- the `TOBIAS <Tool>` command line for the two tools on the hot path
(reference dispatcher: tobias/TOBIAS.py:49-163)."""
import sys

from . import REFERENCE_VERSION, __version__

TOOLS = {"ATACorrect": ("tobias_b200.atacorrect", "main"), "ScoreBigwig": ("tobias_b200.score_bigwig", "main"),
         "FootprintScores": ("tobias_b200.score_bigwig", "main")}


def main():
    argv = sys.argv[1:]
    if not argv or argv[0] in ("-h", "--help"):
        print("TOBIAS (tobias_b200 %s, mirrors TOBIAS %s) -- B200 implementation of the per-base hot path\n"
              "usage: TOBIAS <ATACorrect|ScoreBigwig> [arguments]\n"
              "Other TOBIAS tools are outside this package: use the reference for them." % (__version__, REFERENCE_VERSION))
        return 0
    if argv[0] == "--version":
        print("TOBIAS %s (tobias_b200 %s)" % (REFERENCE_VERSION, __version__))
        return 0
    tool = argv[0]
    if tool not in TOOLS:
        sys.exit("TOBIAS: tool '%s' is not part of the B200 hot path (available: %s)" % (tool, ", ".join(sorted(TOOLS))))
    import importlib
    mod, fn = TOOLS[tool]
    sys.argv = ["TOBIAS " + tool] + argv[1:]
    getattr(importlib.import_module(mod), fn)(argv[1:] if len(argv) > 1 else None)
    return 0


if __name__ == "__main__":
    sys.exit(main())
