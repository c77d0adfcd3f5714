"""
Argument parsers of the two tools on the hot path, flag for flag as in the reference (tobias/parsers.py:17-105;
`--x-y` / `--x_y` aliasing from tobias/utils/utilities.py:300-321), plus `--gpus` / `--device` / `--lm-exact` additions.
`--cores` and `--split` are accepted for command-line compatibility; the GPU path batches all regions and ignores them.
"""
import argparse
import re
import textwrap

from .logger import add_logger_args


def format_help_description(name, description, width=90):
    header = "TOBIAS ~ {0}".format(name)
    ws = int((width - len(header)) / 2.0)
    out = "_" * width + "\n" * 2 + "{0}{1}{0}\n".format(" " * ws, header) + "_" * width + "\n" * 2
    for segment in description.split("\n"):
        out += "\n".join(textwrap.wrap(segment, width)) + "\n"
    if description != "":
        out += "\n" + "-" * width + "\n"
    return out


def add_underscore_options(parser):
    """Every --some-flag is also accepted as --some_flag (hidden), like the reference."""
    for group in list(parser._action_groups):
        for option in list(group._group_actions):
            opt_string = option.option_strings[-1]
            stripped = re.sub(r'^\-*', "", opt_string)
            if "-" in stripped:
                alias = "--" + stripped.replace("-", "_")
                keep = ["nargs", "const", "default", "type", "choices", "required", "metavar"]
                kw = {k: option.__dict__[k] for k in keep}
                kw["nargs"] = "?" if kw["nargs"] == 0 else kw["nargs"]
                if isinstance(option, (argparse._StoreTrueAction,)):
                    parser.add_argument(alias, dest=option.dest, action="store_true", help=argparse.SUPPRESS)
                else:
                    parser.add_argument(alias, dest=option.dest, help=argparse.SUPPRESS, **kw)
    return parser


def _add_gpu_args(parser):
    g = parser.add_argument_group('B200 arguments (additions to the reference command line)')
    g.add_argument('--device', metavar="<int>", type=int, default=0, help="CUDA device of this process (default: 0, or LOCAL_RANK under torchrun)")
    g.add_argument('--lm-exact', action="store_true", help="Window fits in MINPACK's own arithmetic (bit-faithful to scipy's curve_fit on identical inputs; ~3x slower fits). Default: on for --score_mat PWM (bit-identical tracks), off for DWM (one-pass outer iterations, same tolerance as the DWM bias input)")
    return parser


def add_atacorrect_arguments(parser):
    parser.formatter_class = lambda prog: argparse.RawDescriptionHelpFormatter(prog, max_help_position=35, width=90)
    description = "ATACorrect corrects the cutsite-signal from ATAC-seq with regard to the underlying sequence preference of Tn5 transposase.\n\n"
    description += "Usage:\nTOBIAS ATACorrect --bam <reads.bam> --genome <genome.fa> --peaks <peaks.bed>\n\n"
    description += "Output files:\n"
    description += "\n".join(["- <outdir>/<prefix>_{0}.bw".format(track) for track in ["uncorrected", "bias", "expected", "corrected"]]) + "\n"
    description += "- <outdir>/<prefix>_atacorrect.pdf"
    parser.description = format_help_description("ATACorrect", description)
    parser._action_groups.pop()
    reqargs = parser.add_argument_group('Required arguments')
    reqargs.add_argument('-b', '--bam', metavar="<bam>", help="A .bam-file containing reads to be corrected")
    reqargs.add_argument('-g', '--genome', metavar="<fasta>", help="A .fasta-file containing whole genomic sequence")
    reqargs.add_argument('-p', '--peaks', metavar="<bed>", help="A .bed-file containing ATAC peak regions")
    optargs = parser.add_argument_group('Optional arguments')
    optargs.add_argument('--regions-in', metavar="<bed>", help="Input regions for estimating bias (default: regions not in peaks.bed)")
    optargs.add_argument('--regions-out', metavar="<bed>", help="Output regions (default: peaks.bed)")
    optargs.add_argument('--blacklist', metavar="<bed>", help="Blacklisted regions in .bed-format (default: None)")
    optargs.add_argument('--extend', metavar="<int>", type=int, help="Extend output regions with basepairs upstream/downstream (default: 100)", default=100)
    optargs.add_argument('--split-strands', help="Write out tracks per strand", action="store_true")
    optargs.add_argument('--norm-off', help="Switches off normalization based on number of reads", action='store_true')
    optargs.add_argument('--track-off', metavar="<track>", help="Switch off writing of individual .bigwig-tracks (uncorrected/bias/expected/corrected)", nargs="*", choices=["uncorrected", "bias", "expected", "corrected"], default=[])
    optargs.add_argument('--drop-chroms', metavar="<chrom>", help="Drop any chromosomes in the list from the correction. The default is to drop the mitochrondrial chromosome. Default: ['chrM', 'chrMT', 'M', 'MT', 'Mito']", nargs="*", default=['chrM', 'chrMT', 'M', "MT", "Mito"])
    optargs = parser.add_argument_group('Advanced ATACorrect arguments (no need to touch)')
    optargs.add_argument('--k_flank', metavar="<int>", help="Flank +/- of cutsite to estimate bias from (default: 12)", type=int, default=12)
    optargs.add_argument('--read_shift', metavar="<int>", help="Read shift for forward and reverse reads (default: 4 -5)", nargs=2, type=int, default=[4, -5])
    optargs.add_argument('--bg_shift', metavar="<int>", type=int, help="Read shift for estimation of background frequencies (default: 100)", default=100)
    optargs.add_argument('--window', metavar="<int>", help="Window size for calculating expected signal (default: 100)", type=int, default=100)
    optargs.add_argument('--score_mat', metavar="<mat>", help="Type of matrix to use for bias estimation (PWM/DWM) (default: DWM)", choices=["PWM", "DWM"], default="DWM")
    optargs.add_argument('--bias-pkl', metavar="<obj>", help="Path to a pre-calculated AtacBias.pkl-object, as output from a previous ATACorrect run (default: None). Can be used to bypass the internal bias estimation.", default=None)
    runargs = parser.add_argument_group('Run arguments')
    runargs.add_argument('--prefix', metavar="<prefix>", help="Prefix for output files (default: same as .bam file)")
    runargs.add_argument('--outdir', metavar="<directory>", help="Output directory for files (default: current working directory)", default="")
    runargs.add_argument('--cores', metavar="<int>", type=int, help="Number of cores to use for computation (default: 1)", default=1)
    runargs.add_argument('--split', metavar="<int>", type=int, help="Split of multiprocessing jobs (default: 100)", default=100)
    add_logger_args(runargs)
    _add_gpu_args(parser)
    return parser


def add_scorebigwig_arguments(parser):
    parser.formatter_class = lambda prog: argparse.RawDescriptionHelpFormatter(prog, max_help_position=40, width=90)
    description = "ScoreBigwig calculates scores (such as footprint-scores) from bigwig files (such as ATAC-seq cutsites calculated using the ATACorrect tool). "
    description += "NOTE: ScoreBigwig replaces the previous FootprintScores command.\n\n"
    description += "Usage: ScoreBigwig --signal <cutsites.bw> --regions <regions.bed> --output <output.bw>\n\n"
    description += "Output:\n- <output.bw>"
    parser.description = format_help_description("ScoreBigwig", description)
    parser._action_groups.pop()
    required = parser.add_argument_group('Required arguments')
    required.add_argument('-s', '--signal', metavar="<bigwig>", help="A .bw file of ATAC-seq cutsite signal")
    required.add_argument('-o', '--output', metavar="<bigwig>", help="Full path to output bigwig")
    required.add_argument('-r', '--regions', metavar="<bed>", help="Genomic regions to run footprinting within")
    optargs = parser.add_argument_group('Optional arguments')
    optargs.add_argument('--score', metavar="<score>", choices=["footprint", "sum", "mean", "none"], help="Type of scoring to perform on cutsites (footprint/sum/mean/none) (default: footprint)", default="footprint")
    optargs.add_argument('--absolute', action='store_true', help="Convert bigwig signal to absolute values before calculating score")
    optargs.add_argument('--extend', metavar="<int>", type=int, help="Extend input regions with bp (default: 100)", default=100)
    optargs.add_argument('--smooth', metavar="<int>", type=int, help="Smooth output signal by mean in <bp> windows (default: no smoothing)", default=1)
    optargs.add_argument('--min-limit', metavar="<float>", type=float, help="Limit input bigwig value range (default: no lower limit)")
    optargs.add_argument('--max-limit', metavar="<float>", type=float, help="Limit input bigwig value range (default: no upper limit)")
    footprintargs = parser.add_argument_group('Parameters for score == footprint')
    footprintargs.add_argument('--fp-min', metavar="<int>", type=int, help="Minimum footprint width (default: 20)", default=20)
    footprintargs.add_argument('--fp-max', metavar="<int>", type=int, help="Maximum footprint width (default: 50)", default=50)
    footprintargs.add_argument('--flank-min', metavar="<int>", type=int, help="Minimum range of flanking regions (default: 10)", default=10)
    footprintargs.add_argument('--flank-max', metavar="<int>", type=int, help="Maximum range of flanking regions (default: 30)", default=30)
    sumargs = parser.add_argument_group('Parameters for score == sum')
    sumargs.add_argument('--window', metavar="<int>", type=int, help="The window for calculation of sum (default: 100)", default=100)
    runargs = parser.add_argument_group('Run arguments')
    runargs.add_argument('--cores', metavar="<int>", type=int, help="Number of cores to use for computation (default: 1)", default=1)
    runargs.add_argument('--split', metavar="<int>", type=int, help="Split of multiprocessing jobs (default: 100)", default=100)
    add_logger_args(runargs)
    _add_gpu_args(parser)
    return parser
