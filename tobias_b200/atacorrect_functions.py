"""
Worker-function mirrors of tobias/tools/atacorrect_functions.py with the reference's names, arguments and return values:

    count_reads(regions_list, params) -> int                                  (:80-105)
    bias_estimation(regions_list, params) -> AtacBias                         (:109-181)
    bias_correction(regions_list, params, bias_obj) -> [pre_bias, post_bias]  (:191-410)
        + side effect  params.qs["<track>:<strand>"].put((key, (chrom, start, end), float64 array))

`params` carries what the reference's workers read: bam, genome, k_flank, bg_shift, read_shift, score_mat, window,
split_strands, qs.  Each call batches ALL regions of the list into a few GPU launches.  The decoded BAM / FASTA and the
GPU-resident copies are cached per (bam, genome) pair inside the process, so calling the three functions one after the
other (as the reference's driver does) uploads the inputs once.
"""
import numpy as np

from .io.bam import read_bam
from .io.fasta import FastaFile
from .pipeline import AtacBias, Engine, STRANDS, prepost_to_matrices  # noqa: F401  (AtacBias re-exported like the reference)
from .reads import ReadColumns

_engines = {}


def get_engine(params):
    key = (params.bam, params.genome, getattr(params, "device", 0))
    if key not in _engines:
        rc = ReadColumns.load(params.bam) if params.bam.endswith(".npz") else read_bam(params.bam)
        fasta = FastaFile(params.genome)
        names = [n for n in rc.references if n in set(fasta.references)]
        eng = Engine(getattr(params, "device", 0))
        eng.load_genome_fasta(fasta, names)
        fasta.close()
        eng.load_reads(rc)
        _engines.clear()
        _engines[key] = eng
    return _engines[key]


def release_engines():
    for e in _engines.values():
        e.close()
    _engines.clear()


def count_reads(regions_list, params):
    return get_engine(params).count_reads(regions_list, tuple(params.read_shift))


def bias_estimation(regions_list, params):
    eng = get_engine(params)
    return eng.estimate_bias(regions_list, params.k_flank, params.bg_shift, tuple(params.read_shift), params.score_mat)


def bias_correction(regions_list, params, bias_obj):
    eng = get_engine(params)
    L = 2 * params.k_flank + 1
    if len(regions_list) == 0:
        return list(prepost_to_matrices(np.zeros((2, 3, 5, L)), L))
    eng.set_bias(bias_obj)
    order = sorted(range(len(regions_list)), key=lambda i: (eng.index[regions_list[i].chrom], regions_list[i].start))
    regs = [regions_list[i] for i in order]
    lm_exact = getattr(params, "lm_exact", params.score_mat == "PWM")
    eng.correct(regs, bias_obj, params.k_flank, params.window, tuple(params.read_shift), lm_exact)
    got = eng.tracks(split_strands=True, as_f64=True)
    offs = np.concatenate([[0], np.cumsum([r.end - r.start for r in regs])])
    pos = {order[k]: k for k in range(len(regs))}
    strands_to_write = ["forward", "reverse"] if params.split_strands else ["both"]
    for i, region in enumerate(regions_list):                  # queue order = input order, like the reference's loop
        k = pos[i]
        a, b = int(offs[k]), int(offs[k + 1])
        reg_key = (region.chrom, region.start, region.end)
        for track in ("uncorrected", "bias", "expected", "corrected"):
            for strand in strands_to_write:
                key = "{0}:{1}".format(track, strand)
                if key in params.qs:
                    if strand == "both":
                        sig = got[track][0][a:b] + got[track][1][a:b]
                    else:
                        sig = got[track][STRANDS.index(strand)][a:b].copy()
                    params.qs[key].put((key, reg_key, sig))
    pre, post = prepost_to_matrices(got["prepost"], L)
    return [pre, post]
