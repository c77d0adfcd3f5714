"""
`TOBIAS ScoreBigwig` on the B200 (reference driver: tobias/tools/score_bigwig.py:102-219; worker `calculate_scores`
:33-98): same arguments and output bigWig; all regions are scored in one batched launch sequence (K4).
"""
import argparse
import sys

import numpy as np

from . import lib
from .atacorrect import check_files, write_track
from .io.bigwig import BigWigReader
from .logger import TobiasLogger
from .parsers import add_scorebigwig_arguments, add_underscore_options
from .regions import OneRegion, RegionList


def region_flank(args):
    """score_bigwig.py:153-159."""
    if args.score == "sum":
        return int(args.window / 2.0)
    if args.score in ("footprint", "FOS"):
        return int(args.flank_max)
    return 0


def calculate_scores(regions, args, ctx=None):
    """Worker mirror (score_bigwig.py:33-98): scores `regions` (a RegionList) of the bigWig args.signal and puts
    ("scores", (chrom, start, end), float64 array) on args.writer_qs["scores"] per region, in input order.  Returns 1."""
    own = ctx is None
    if own:
        ctx = lib.Context(getattr(args, "device", 0))
    try:
        flank = args.region_flank
        bw = BigWigReader(args.signal)
        ext, lens, keys, spans = [], [], [], []
        for region in regions:
            region.extend_reg(flank)
            keys.append((region.chrom, region.start + flank, region.end - flank))
            spans.append((region.chrom, region.start, region.end))
            lens.append(region.end - region.start)
        # runs of regions on one chromosome with ascending starts are read in one sweep over the bigWig (values_many); float32, NaN where empty
        i = 0
        while i < len(spans):
            j = i + 1
            while j < len(spans) and spans[j][0] == spans[i][0] and spans[j][1] >= spans[j - 1][1]:
                j += 1
            flat, _off = bw.values_many(spans[i][0], [s[1] for s in spans[i:j]], [s[2] for s in spans[i:j]])
            ext.append(flat)
            i = j
        bw.close()
        if not ext:
            return 1
        p = ctx.score_params(args.score, flank=flank, absolute=args.absolute, min_limit=args.min_limit, max_limit=args.max_limit,
                             window=args.window, smooth=args.smooth, flank_min=args.flank_min, flank_max=args.flank_max,
                             fp_min=args.fp_min, fp_max=args.fp_max, as_f64=True)
        scores = ctx.score_regions(np.concatenate(ext), lens, p)
        off = 0
        for key in keys:
            n = key[2] - key[1]
            args.writer_qs["scores"].put(("scores", key, scores[off:off + n]))
            off += n
        return 1
    finally:
        if own:
            ctx.close()


class _ListQueue:
    def __init__(self):
        self.items = []

    def put(self, x):
        self.items.append(x)


def run_scorebigwig(args):
    for arg in ("signal", "output", "regions"):
        if getattr(args, arg) is None:
            sys.exit("ERROR: Missing argument --{0}".format(arg))
    check_files([args.signal, args.regions], "r")
    check_files([args.output], "w")
    logger = TobiasLogger("ScoreBigwig", args.verbosity)
    logger.begin()
    parser = add_scorebigwig_arguments(argparse.ArgumentParser())
    logger.arguments_overview(parser, args)
    logger.output_files([args.output])

    import time
    t_mark = [time.perf_counter()]
    phases = {}

    def phase(name):
        now = time.perf_counter()
        phases[name] = round(now - t_mark[0], 3)
        t_mark[0] = now
    logger.info("Processing input files")
    logger.info("- Opening input cutsite bigwig")
    bw = BigWigReader(args.signal)
    chrom_info = {c: int(n) for c, n in bw.chroms.items()}
    bw.close()
    logger.info("- Getting output regions ready")
    regions = RegionList().from_bed(args.regions)
    not_in_bigwig = list(set(regions.get_chroms()) - set(chrom_info))
    if not_in_bigwig:
        logger.warning("Contigs {0} were found in input --regions, but were not found in input --signal. These regions cannot be scored and will therefore be excluded from output.".format(not_in_bigwig))
        regions = regions.remove_chroms(not_in_bigwig)
    regions.apply_method(OneRegion.extend_reg, args.extend)
    regions.merge()
    regions = regions.apply_method(OneRegion.check_boundary, chrom_info, "cut")
    args.region_flank = region_flank(args)
    kept = RegionList()
    for region in regions:                                   # score_bigwig.py:162-165
        region.extend_reg(args.region_flank)
        region = region.check_boundary(chrom_info, "cut")
        if region is None:
            continue
        region.extend_reg(-args.region_flank)
        if region.get_length() > 0:
            kept.append(region)
    regions = kept
    reference_chroms = sorted(chrom_info)
    header = [(c, chrom_info[c]) for c in reference_chroms]
    regions.loc_sort(reference_chroms)

    phase("headers_and_regions")
    logger.info("Calculating footprints in regions...")
    q = _ListQueue()
    args.writer_qs = {"scores": q}
    from copy import deepcopy
    calculate_scores(deepcopy(regions), args)
    vals = np.concatenate([x[2] for x in q.items]).astype(np.float32) if q.items else np.zeros(0, np.float32)
    offs = np.concatenate([[0], np.cumsum([x[2].shape[0] for x in q.items])])[:-1].tolist() if q.items else []
    phase("read_signal_and_score")
    logger.info("Closing {0} (this might take some time)".format(args.output))
    write_track(args.output, header, list(regions), vals, offs)
    phase("write_bigwig")
    logger.end()
    return {"phases_s": phases}


def main(argv=None):
    parser = add_scorebigwig_arguments(argparse.ArgumentParser())
    add_underscore_options(parser)
    args = parser.parse_args(argv)
    if argv is None and len(sys.argv[1:]) == 0:
        parser.print_help()
        sys.exit()
    run_scorebigwig(args)


if __name__ == '__main__':
    main()
