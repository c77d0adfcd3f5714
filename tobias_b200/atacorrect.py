"""
`TOBIAS ATACorrect` on the B200: same arguments, same output files, same STATS log lines as the reference driver
(tobias/tools/atacorrect.py:53-487), with the multiprocessing fan-out over region chunks replaced by batched GPU stages
(tobias_b200.pipeline.Engine).  Under torchrun (one process per GPU) contigs are sharded over the ranks; the only
collectives are the all-reduces of the count tensors / normalisation counters and of the verification PWMs, and rank 0
gathers the tracks for the bigWig writers (tobias_b200.dist).
"""
import argparse
import itertools
import os
import sys
from collections import OrderedDict

import numpy as np

from . import lib
from .dist import Comm, shard_contigs
from .io.bam import BamReader
from .io.bigwig import BigWigWriter
from .io.fasta import FastaFile
from .logger import TobiasLogger
from .parsers import add_atacorrect_arguments
from .pipeline import AtacBias, Engine, STRANDS, prepare_regions, prepost_to_matrices
from .regions import RegionList


def check_files(paths, action="r"):
    for p in paths:
        if p is None:
            continue
        if action == "r" and not os.path.exists(p):
            sys.exit("ERROR: {0} does not exist".format(p))
        if action == "w":
            d = os.path.dirname(os.path.abspath(p))
            if os.path.exists(p) and not os.access(p, os.W_OK):
                sys.exit("ERROR: {0} could not be opened for writing".format(p))
            if not os.path.isdir(d):
                sys.exit("ERROR: directory {0} does not exist".format(d))


def write_track_arrays(path, header, cid, start, end, values, offsets):
    """bigwig_writer semantics (tobias/utils/utilities.py:129-232): regions in header order, non-zero positions only, float32 values,
    span 1.  Region i = (header contig cid[i], start[i], end[i]) holds values[offsets[i] : offsets[i] + end[i] - start[i]].  Array
    operations throughout: the non-zero positions of the whole track are found at once and handed to the writer contig by contig
    (sections are compressed in parallel, zoom levels follow)."""
    cid, start, end, offsets = (np.asarray(a, dtype=np.int64) for a in (cid, start, end, offsets))
    values = np.asarray(values)
    w = BigWigWriter(path)
    w.add_header(header)
    if cid.size:
        order = np.lexsort((start, cid))
        cid, start, end, offsets = cid[order], start[order], end[order], offsets[order]
        lens = end - start
        # concatenated in header order: out position of every value, then the non-zero ones
        src = np.repeat(offsets - np.concatenate([[0], np.cumsum(lens)[:-1]]), lens) + np.arange(int(lens.sum()), dtype=np.int64)
        vals = values[src]
        pos = np.repeat(start - np.concatenate([[0], np.cumsum(lens)[:-1]]), lens) + np.arange(int(lens.sum()), dtype=np.int64)
        reg_cid = np.repeat(cid, lens)
        nz = np.flatnonzero(vals)
        pos, vals, reg_cid = pos[nz], vals[nz], reg_cid[nz]
        cuts = np.flatnonzero(np.diff(reg_cid)) + 1
        for a, b in zip(np.concatenate([[0], cuts]).tolist(), np.concatenate([cuts, [reg_cid.size]]).tolist()):
            if b > a:
                w.add_entries(header[int(reg_cid[a])][0], pos[a:b], vals[a:b])
    w.close()


def write_track(path, header, regions, values, region_offsets):
    """write_track_arrays for a list of regions (objects with chrom / start / end)."""
    order = {c: i for i, (c, _) in enumerate(header)}
    write_track_arrays(path, header, [order[r.chrom] for r in regions], [r.start for r in regions], [r.end for r in regions], values,
                       region_offsets)


def run_atacorrect(args):
    if args.bam is None:
        sys.exit("Error: No .bam-file given")
    if args.genome is None:
        sys.exit("Error: No .fasta-file given")
    if args.peaks is None:
        sys.exit("Error: No .peaks-file given")
    import time
    phases, t_last = OrderedDict(), [time.perf_counter()]

    def phase(name):                    # wall clock per phase of the run (bench.py --cli reports it)
        now = time.perf_counter()
        phases[name] = phases.get(name, 0.0) + (now - t_last[0])
        t_last[0] = now
    comm = Comm.init_from_env()
    root = comm.rank == 0
    args.prefix = os.path.splitext(os.path.basename(args.bam))[0] if args.prefix is None else args.prefix
    args.outdir = os.path.abspath(args.outdir) if args.outdir else os.path.abspath(os.getcwd())
    tracks = [t for t in ("uncorrected", "bias", "expected", "corrected") if t not in args.track_off]
    strands_out = ["forward", "reverse"] if args.split_strands else ["both"]
    output_bws = {}
    for track in tracks:
        output_bws[track] = {}
        for strand in strands_out:
            elements = [args.prefix, track] if strand == "both" else [args.prefix, track, strand]
            output_bws[track][strand] = os.path.join(args.outdir, "{0}.bw".format("_".join(elements)))
    bigwigs = [output_bws[t][s] for (t, s) in itertools.product(tracks, strands_out)]
    figures_f = os.path.join(args.outdir, "{0}_atacorrect.pdf".format(args.prefix))
    output_files = list(OrderedDict.fromkeys(bigwigs + [figures_f]))

    logger = TobiasLogger("ATACorrect", args.verbosity if root else 0)
    logger.begin()
    parser = add_atacorrect_arguments(argparse.ArgumentParser())
    logger.arguments_overview(parser, args)
    logger.output_files(output_files)

    logger.info("----- Processing input data -----")
    check_files([args.bam, args.genome, args.peaks, args.regions_in, args.regions_out, args.blacklist, args.bias_pkl], "r")
    if root:
        os.makedirs(args.outdir, exist_ok=True)
    comm.barrier()
    check_files(bigwigs, "w")

    logger.info("Reading info from .bam file")
    bam = BamReader(args.bam)                   # header + BGZF block table only; the records follow once this rank knows its contigs
    bam_references = list(bam.references)
    bam_chrom_info = dict(zip(bam.references, bam.lengths))
    logger.info("Reading info from .fasta file")
    fasta = FastaFile(args.genome)
    fasta_chrom_info = dict(zip(fasta.references, fasta.lengths))
    chrom_in_common = set(bam_chrom_info).intersection(fasta_chrom_info)
    for chrom in chrom_in_common:
        if bam_chrom_info[chrom] != fasta_chrom_info[chrom]:
            logger.warning("(Fastafile)\t{0} has length {1}".format(chrom, fasta_chrom_info[chrom]))
            logger.warning("(Bamfile)\t{0} has length {1}".format(chrom, bam_chrom_info[chrom]))
            sys.exit("Error: .bam and .fasta have different chromosome lengths. Please make sure the genome file is similar to the one used in mapping.")
    not_in_fasta = set(bam_references) - set(fasta_chrom_info)
    if len(not_in_fasta) > 1:
        logger.warning("The following contigs in --bam did not have sequences in --fasta: {0}. NOTE: These contigs will be skipped in calculation and output.".format(not_in_fasta))
    bam_references = [r for r in bam_references if r in fasta_chrom_info]
    chrom_in_common = [r for r in bam_references]
    kept = [c for c in chrom_in_common if c not in args.drop_chroms]
    dropped = set(chrom_in_common) - set(kept)
    bam_references = kept
    if dropped:
        logger.info("The following contigs were dropped from analysis because they were found in '--drop-chroms': {0}".format(list(dropped)))
    else:
        logger.warning("No additional chromosomes were removed. Consider using '--drop-chroms' to remove mitochondrial and/or other unwanted contigs.")
    if len(bam_references) == 0:
        logger.error("No common contigs left to run ATACorrect on. Please check that '--bam' and '--fasta' are matching.")
        sys.exit()

    logger.info("Processing input/output regions")
    chrom_info = {c: bam_chrom_info[c] for c in bam_references}
    rd = prepare_regions(RegionList().from_bed(args.peaks), chrom_info, bam_references, extend=args.extend, k_flank=args.k_flank,
                         window=args.window,
                         regions_in=RegionList().from_bed(args.regions_in) if args.regions_in else None,
                         regions_out=RegionList().from_bed(args.regions_out) if args.regions_out else None,
                         blacklist=RegionList().from_bed(args.blacklist) if args.blacklist else None)
    genome_bp = sum(r.get_length() for r in rd["genome"])
    for key in rd:
        total_bp = sum(r.get_length() for r in rd[key])
        logger.stats("{0}: {1} regions | {2} bp | {3:.2f}% coverage".format(key, len(rd[key]), total_bp, total_bp / genome_bp * 100))
    if min(len(rd[k]) for k in ("input_regions", "output_regions", "peak_regions", "nonpeak_regions")) == 0:
        logger.error("No regions found - exiting!")
        sys.exit()

    phase("headers_and_regions")
    # ---- shard by contig, make this rank's data resident on its GPU
    owner = shard_contigs(chrom_info, comm.world)
    mine = [c for c in bam_references if owner[c] == comm.rank]
    local = {k: RegionList([r for r in rd[k] if owner[r.chrom] == comm.rank]) for k in
             ("input_regions", "output_regions", "peak_regions", "nonpeak_regions")}
    device = int(os.environ.get("LOCAL_RANK", args.device))
    # every rank inflates and decodes the BGZF blocks of its own contigs only; without a .bai / .tbidx rank 0 first walks the file once
    # (no records kept) and leaves the contig offsets next to the BAM
    if comm.world > 1:
        if root and not bam.has_index():
            logger.info("Indexing {0} (one pass, contig offsets only)".format(args.bam))
            bam.build_index()
        comm.barrier()
    eng = None
    if mine:              # a rank without contigs (more ranks than kept contigs) holds no data: it adds zeros to the all-reduces, nothing to the gathers
        eng = Engine(device)
        phase("gpu_context")
        eng.load_genome_fasta(fasta, mine)
        phase("genome_pack_and_upload")
        rc_mine = bam.read(contigs=mine)
        phase("bam_decode")
        eng.load_reads(rc_mine)
        del rc_mine
        phase("reads_upload")
    bam.close()
    read_shift = tuple(args.read_shift)

    logger.comment("")
    logger.info("----- Estimating normalization factors -----")
    if not args.norm_off:
        logger.info("Counting reads in peak regions")
        logger.info("Counting reads in nonpeak regions")
        cnt = np.array([eng.count_reads(local["peak_regions"], read_shift), eng.count_reads(local["nonpeak_regions"], read_shift)] if eng else [0, 0],
                       dtype=np.int64)
        comm.allreduce_sum(cnt)
        reads_peaks, reads_nonpeaks = int(cnt[0]), int(cnt[1])
        reads_total = reads_peaks + reads_nonpeaks
        logger.stats("TOTAL_READS\t{0}".format(reads_total))
        logger.stats("PEAK_READS\t{0}".format(reads_peaks))
        logger.stats("NONPEAK_READS\t{0}".format(reads_nonpeaks))
        lib_norm = 10000000 / reads_total
        frip = reads_peaks / reads_total
        correct_factor = lib_norm * (1 / frip)
        logger.stats("LIB_NORM\t{0:.5f}".format(lib_norm))
        logger.stats("FRiP\t{0:.5f}".format(frip))
    else:
        logger.info("Normalization was switched off")
        correct_factor = 1.0
    logger.stats("CORRECTION_FACTOR:\t{0:.5f}".format(correct_factor))
    phase("count_reads")

    logger.comment("")
    if args.bias_pkl is None:
        logger.info("Started estimation of sequence bias...")
        bias_obj = (eng.estimate_bias(local["input_regions"], args.k_flank, args.bg_shift, read_shift, args.score_mat) if eng
                    else AtacBias(2 * args.k_flank + 1, args.score_mat))
        packed = bias_obj.pack_counts()
        comm.allreduce_sum(packed)
        bias_obj.unpack_counts(packed)
        logger.debug("Bias estimated\tno_reads: {0}".format(bias_obj.no_reads))
    else:
        logger.info("Loading sequence bias from '--bias-pkl' file...")
        bias_obj = AtacBias().from_pickle(args.bias_pkl)
    bias_obj.correction_factor = correct_factor
    logger.info("Finalizing bias motif for scoring")
    for strand in STRANDS:
        bias_obj.bias[strand].prepare_mat()
    if root:
        bias_obj.to_pickle(os.path.join(args.outdir, args.prefix + "_AtacBias.pickle"))
    phase("bias_estimation")

    logger.comment("")
    logger.info("----- Correcting reads from .bam within output regions -----")
    if eng:
        eng.set_bias(bias_obj)
    out_regions = local["output_regions"]
    out_regions.loc_sort(mine)
    lm_exact = bool(args.lm_exact) or args.score_mat == "PWM"
    n_out = len(out_regions)
    if n_out:
        eng.correct(out_regions, bias_obj, args.k_flank, args.window, read_shift, lm_exact)
        got = eng.tracks(split_strands=args.split_strands, as_f64=False, tracks=tracks, prepost=True)
        prepost = got["prepost"]
    else:
        got, prepost = {}, np.zeros((2, 3, 5, 2 * args.k_flank + 1))
    comm.allreduce_sum(prepost)
    phase("bias_correction")

    # ---- rank 0 gathers and writes the bigWigs (header = contigs in BAM order, like the reference)
    header = [(c, bam_chrom_info[c]) for c in bam_references]
    reg_tbl = np.array([(bam_references.index(r.chrom), r.start, r.end) for r in out_regions], dtype=np.int64).reshape(-1, 3)
    all_tbl = comm.gather_arrays(reg_tbl.ravel())
    jobs = []
    if root:
        tbl = np.concatenate([t.reshape(-1, 3) for t in all_tbl]) if all_tbl else np.zeros((0, 3), np.int64)
        offs = np.concatenate([[0], np.cumsum(tbl[:, 2] - tbl[:, 1])])[:-1]
    for track in tracks:
        for si, strand in enumerate(strands_out):
            vals = got[track][si] if (args.split_strands and n_out) else (got[track] if n_out else np.zeros(0, np.float32))
            parts = comm.gather_arrays(np.ascontiguousarray(vals))
            if root:
                jobs.append((output_bws[track][strand], np.concatenate(parts) if parts else np.zeros(0, np.float32)))
    if root:
        # the files are independent: written side by side (numpy and the native helpers release the GIL)
        from concurrent.futures import ThreadPoolExecutor
        for path, _v in jobs:
            logger.info("Closing {0} (this might take some time)".format(path))
        with ThreadPoolExecutor(max(1, min(4, len(jobs)))) as ex:
            list(ex.map(lambda job: write_track_arrays(job[0], header, tbl[:, 0], tbl[:, 1], tbl[:, 2], job[1], offs), jobs))

    phase("write_bigwigs")
    logger.comment("")
    logger.info("Verifying bias correction")
    pre_bias, post_bias = prepost_to_matrices(prepost, 2 * args.k_flank + 1)
    for strand in STRANDS:
        abssum = np.abs(np.sum(post_bias[strand].neg_counts, axis=0))                 # atacorrect.py:461-466
        post_bias[strand].neg_counts = post_bias[strand].neg_counts + abssum
        post_bias[strand].counts += post_bias[strand].neg_counts
        pre_bias[strand].prepare_mat()
        post_bias[strand].prepare_mat()
        pre_var = np.mean(np.var(pre_bias[strand].bias_pwm, axis=1)[:4])
        post_var = np.mean(np.var(post_bias[strand].bias_pwm, axis=1)[:4])
        logger.stats("BIAS\tpre-bias variance {0}:\t{1:.7f}".format(strand, pre_var))
        logger.stats("BIAS\tpost-bias variance {0}:\t{1:.7f}".format(strand, post_var))
    if root:
        from .plotting import write_report
        write_report(figures_f, bias_obj, pre_bias, post_bias)          # bias motifs + pre / post nucleotide frequencies (atacorrect.py:347-348,476-478)
    if eng:
        eng.close()
    fasta.close()
    phase("verification_and_report")
    logger.end()
    return {"correction_factor": correct_factor, "bias": bias_obj, "pre_bias": pre_bias, "post_bias": post_bias,
            "phases_s": {k: round(v, 3) for k, v in phases.items()}}


def main(argv=None):
    parser = add_atacorrect_arguments(argparse.ArgumentParser())
    from .parsers import add_underscore_options
    add_underscore_options(parser)
    args = parser.parse_args(argv)
    if argv is None and len(sys.argv[1:]) == 0:
        parser.print_help()
        sys.exit()
    run_atacorrect(args)


if __name__ == '__main__':
    main()
