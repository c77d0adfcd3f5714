"""
The hot path as one engine: GPU-resident genome + alignment records, and the four stages of
ATACorrect (normalisation counts -> bias estimation -> bias model -> correction) plus ScoreBigwig scoring, each
batched over ALL regions of a shard in a few kernel launches instead of the reference's per-region Python loops
(tobias/tools/atacorrect.py:285-424 fan-out over tobias/tools/atacorrect_functions.py:80-410).

Used by the worker-function mirrors (tobias_b200/atacorrect_functions.py, score_bigwig.py), by the CLI drivers, by
bench.py and by the multi-GPU layer (tobias_b200/dist.py).
"""
import io
import pickle
from copy import deepcopy

import numpy as np

from . import lib
from .regions import OneRegion, RegionList, to_arrays
from .sequences import DiNucleotideMatrix, NucleotideMatrix, SequenceMatrix

STRANDS = ("forward", "reverse")

# ---- <prefix>_AtacBias.pickle compatibility (tobias/tools/atacorrect_functions.py:62-76, atacorrect.py:316-354) ----
# The reference pickles plain instances of tobias.tools.atacorrect_functions.AtacBias holding
# tobias.utils.sequences.{DiNucleotideMatrix,NucleotideMatrix}; the attribute names are the same here, so compatibility is a
# matter of class paths: pickles are WRITTEN with the reference's paths (loadable by `TOBIAS ATACorrect --bias-pkl` of the
# reference) and READ with either set of paths, without TOBIAS being installed and without touching sys.modules.
_REF_ATACBIAS = ("tobias.tools.atacorrect_functions", "AtacBias")
_REF_SEQ = "tobias.utils.sequences"


class _RefPathPickler(pickle._Pickler):
    """Pure-Python pickler that writes this package's bias classes under the reference's module paths."""

    def save_global(self, obj, name=None):
        path = _ref_path_of(obj)
        if path is None:
            return super().save_global(obj, name)
        self.write(pickle.GLOBAL + path[0].encode() + b"\n" + path[1].encode() + b"\n")
        self.memoize(obj)

    dispatch = dict(pickle._Pickler.dispatch)
    dispatch[type] = save_global


class _RefPathUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if (module, name) == _REF_ATACBIAS:
            return AtacBias
        if module == _REF_SEQ and name in ("DiNucleotideMatrix", "NucleotideMatrix", "SequenceMatrix"):
            return {"DiNucleotideMatrix": DiNucleotideMatrix, "NucleotideMatrix": NucleotideMatrix, "SequenceMatrix": SequenceMatrix}[name]
        return super().find_class(module, name)


def _ref_path_of(obj):
    if obj is AtacBias:
        return _REF_ATACBIAS
    for cls in (DiNucleotideMatrix, NucleotideMatrix, SequenceMatrix):
        if obj is cls:
            return (_REF_SEQ, cls.__name__)
    return None


class AtacBias:
    """Estimated bias of one run (tobias/tools/atacorrect_functions.py:43-76)."""

    def __init__(self, L=10, stype="DWM"):
        self.stype = stype
        self.bias = {"forward": SequenceMatrix.create(L, stype), "reverse": SequenceMatrix.create(L, stype),
                     "both": SequenceMatrix.create(L, stype)}
        self.no_reads = 0
        self.correction_factor = 1.0

    def join(self, obj):
        for s in ("forward", "reverse", "both"):
            self.bias[s].add_counts(obj.bias[s])
        self.no_reads += obj.no_reads

    def to_pickle(self, f):
        """Writes a pickle the reference's `--bias-pkl` loads (same class paths and attribute names)."""
        buf = io.BytesIO()
        _RefPathPickler(buf, protocol=2).dump(self)
        with open(f, "wb") as handle:
            handle.write(buf.getvalue())
        return self

    def from_pickle(self, f):
        """Reads a pickle written by the reference's ATACorrect or by this package."""
        with open(f, "rb") as handle:
            obj = _RefPathUnpickler(handle).load()
        if not isinstance(obj, AtacBias):
            raise ValueError("%s does not hold an AtacBias object" % f)
        if not hasattr(obj, "correction_factor"):
            obj.correction_factor = 1.0
        return obj

    # ---- flat int64 view of everything that is summed across shards (exact: all entries are integers) ----
    def pack_counts(self):
        parts = [np.asarray(self.bias[s].counts).ravel() for s in STRANDS] + \
                [np.asarray(self.bias[s].bg_counts).ravel() for s in STRANDS]
        scal = [self.no_reads] + [self.bias[s].no_bias for s in STRANDS] + [self.bias[s].no_bg for s in STRANDS]
        return np.concatenate(parts + [np.array(scal, dtype=np.float64)]).astype(np.int64)

    def unpack_counts(self, buf):
        off = 0
        for key in ("counts", "bg_counts"):
            for s in STRANDS:
                arr = getattr(self.bias[s], key)
                n = arr.size
                setattr(self.bias[s], key, buf[off:off + n].astype(np.float64).reshape(arr.shape))
                off += n
        self.no_reads = int(buf[off])
        for i, s in enumerate(STRANDS):
            self.bias[s].no_bias = int(buf[off + 1 + i])
            self.bias[s].no_bg = int(buf[off + 3 + i])


def prepare_regions(peaks, chrom_info, chrom_order, extend=100, k_flank=12, window=100, regions_in=None, regions_out=None,
                    blacklist=None, split_len=50000):
    """Region set algebra of the ATACorrect driver (tobias/tools/atacorrect.py:189-263).
    peaks / regions_in / regions_out / blacklist: RegionList (or None).  chrom_info: name -> length for the contigs
    taking part; chrom_order: contig names in BAM header order.  Returns dict of RegionLists."""
    chroms = [c for c in chrom_order if c in chrom_info]
    genome_regions = RegionList([OneRegion([c, 0, chrom_info[c]]) for c in chroms])
    peak_regions = deepcopy(peaks)
    peak_regions.merge()
    kept = RegionList()
    for r in peak_regions:
        r2 = r.check_boundary(chrom_info, "cut")
        if r2 is not None:
            kept.append(r2)
    peak_regions = kept
    nonpeak_regions = deepcopy(genome_regions).subtract(peak_regions)
    if regions_in is not None:
        input_regions = deepcopy(regions_in)
        input_regions.merge()
        input_regions = input_regions.apply_method(OneRegion.check_boundary, chrom_info, "cut")
    else:
        input_regions = deepcopy(nonpeak_regions)
    output_regions = deepcopy(regions_out) if regions_out is not None else deepcopy(peak_regions)
    flank_extend = k_flank + int(window / 2.0)
    output_regions.apply_method(OneRegion.extend_reg, extend + flank_extend)
    output_regions.merge()
    output_regions = output_regions.apply_method(OneRegion.check_boundary, chrom_info, "cut")
    output_regions.apply_method(OneRegion.extend_reg, -flank_extend)
    blacklist_regions = deepcopy(blacklist) if blacklist is not None else RegionList()
    out = {"genome": genome_regions, "input_regions": input_regions, "output_regions": output_regions,
           "peak_regions": peak_regions, "nonpeak_regions": nonpeak_regions, "blacklist_regions": blacklist_regions}
    for sub in ("input_regions", "output_regions", "peak_regions", "nonpeak_regions"):
        regs = out[sub]
        regs.subtract(blacklist_regions)
        regs = regs.apply_method(OneRegion.split_region, split_len)
        regs.keep_chroms(chroms)
        out[sub] = regs
    out["output_regions"].loc_sort(chroms)
    return out


class Engine:
    """One GPU (one tb_ctx): resident genome + reads, batched stages."""

    def __init__(self, device=0, lib_path=None):
        self.ctx = lib.Context(device, lib_path)
        self.names = []
        self.index = {}
        self.lengths = []

    def close(self):
        self.ctx.close()

    # ------------------------------------------------------------------ residency
    def load_genome(self, names, contigs):
        """names: contig names in id order; contigs: matching list of uint8 ASCII arrays."""
        self.names = list(names)
        self.index = {n: i for i, n in enumerate(self.names)}
        self.lengths = [int(len(c)) for c in contigs]
        self.ctx.genome_load(contigs)

    def load_genome_fasta(self, fasta, names=None, packed=True):
        names = list(names) if names is not None else list(fasta.references)
        self.names = names
        self.index = {n: i for i, n in enumerate(names)}
        self.lengths = [fasta.get_reference_length(n) for n in names]
        self.ctx.genome_create(self.lengths)
        if packed:            # 2 bit + N mask, packed on the host cores once and cached next to the FASTA (<fasta>.tb2bit)
            from .io.fasta import PackedGenome
            pg = PackedGenome.from_fasta(fasta, names)
            for i in range(len(names)):
                self.ctx.genome_upload_packed(i, pg.seq2[i], pg.nmask[i], self.lengths[i])
            return
        chunk = 1 << 26
        for i, n in enumerate(names):
            for off in range(0, self.lengths[i], chunk):
                self.ctx.genome_upload(i, fasta.fetch_array(n, off, min(off + chunk, self.lengths[i])), off)

    def load_reads(self, rc):
        """rc: ReadColumns whose contig ids refer to rc.references; remapped to this engine's contig ids."""
        if list(rc.references) != self.names:
            remap = np.array([self.index.get(n, -1) for n in rc.references], dtype=np.int32)
            cid = remap[rc.chrom_id]
        else:
            cid = rc.chrom_id
        flags = (rc.is_reverse.astype(np.uint8) * lib.TB_READ_REVERSE
                 | rc.five_prime_is_match().astype(np.uint8) * lib.TB_READ_5P_MATCH
                 | (~rc.usable()).astype(np.uint8) * lib.TB_READ_SKIP)
        self.ctx.reads_upload(cid, rc.ref_start, rc.ref_end, rc.clip5(), flags)

    def region_arrays(self, regions):
        return to_arrays(regions, self.index)

    # ------------------------------------------------------------------ stages
    def count_reads(self, regions, read_shift=(4, -5)):
        """count_reads (atacorrect_functions.py:80-105) over a whole RegionList."""
        if len(regions) == 0:
            return 0
        order = sorted(range(len(regions)), key=lambda i: (self.index[regions[i].chrom], regions[i].start, regions[i].end))
        c, s, e = self.region_arrays([regions[i] for i in order])
        return self.ctx.count_reads(c, s, e, read_shift)

    def estimate_bias(self, regions, k_flank=12, bg_shift=100, read_shift=(4, -5), score_mat="DWM", cap=10):
        """bias_estimation (atacorrect_functions.py:109-181) over a whole RegionList -> AtacBias (counts, not prepared)."""
        L = 2 * k_flank + 1
        bias = AtacBias(L, score_mat)
        if len(regions) == 0:
            return bias
        regs = sorted(regions, key=lambda r: (self.index[r.chrom], r.start, r.end))
        c, s, e = self.region_arrays(regs)
        counts, scalars = self.ctx.bias_counts(c, s, e, k_flank, bg_shift, read_shift, cap, score_mat == "DWM")
        for si, st in enumerate(STRANDS):
            bias.bias[st].counts = counts[si, 0].astype(np.float64)
            bias.bias[st].bg_counts = counts[si, 1].astype(np.float64)
            bias.bias[st].no_bias = int(scalars[1 + 2 * si])
            bias.bias[st].no_bg = int(scalars[2 + 2 * si])
        bias.no_reads = int(scalars[0])
        return bias

    def set_bias(self, bias_obj):
        """Upload the prepared (prepare_mat) forward / reverse tables."""
        for si, st in enumerate(STRANDS):
            bias_obj.bias[st].upload(self.ctx, si)

    def correct(self, regions, bias_obj, k_flank=12, window=100, read_shift=(4, -5), lm_exact_order=True):
        """bias_correction (atacorrect_functions.py:191-410) over a whole RegionList (sorted by contig id, start);
        results stay on the device: fetch with `tracks()`."""
        c, s, e = self.region_arrays(regions)
        self.ctx.correct_run(c, s, e, k_flank, window, bias_obj.correction_factor, read_shift, lm_exact_order)
        self._out_len = e - s
        return self

    def tracks(self, split_strands=False, as_f64=False, tracks=lib.TRACKS, prepost=True):
        return self.ctx.correct_download(split_strands, as_f64, tracks, prepost)

    def footprints_from_corrected(self, flank_min=10, flank_max=30, fp_min=20, fp_max=50, smooth=1, as_f64=False):
        """ScoreBigwig --score footprint over the corrected track just computed, chained on the device."""
        self.ctx.signal_from_corrected(flank_max, self._out_len)
        p = self.ctx.score_params("footprint", flank=flank_max, flank_min=flank_min, flank_max=flank_max, fp_min=fp_min,
                                  fp_max=fp_max, smooth=smooth, as_f64=as_f64)
        self.ctx.score_run(p)
        return self.ctx.score_download()


def prepost_to_matrices(prepost, L):
    """[2][3][5][L] verification counts -> (pre_bias, post_bias) dicts of PWM SequenceMatrix (atacorrect_functions.py:208-209)."""
    pre = {s: SequenceMatrix.create(L, "PWM") for s in STRANDS}
    post = {s: SequenceMatrix.create(L, "PWM") for s in STRANDS}
    for si, s in enumerate(STRANDS):
        pre[s].counts = prepost[si, 0].copy()
        pre[s].no_bias = float(prepost[si, 0, :, 0].sum())
        post[s].counts = prepost[si, 1].copy()
        post[s].neg_counts = prepost[si, 2].copy()
        post[s].no_bias = float(prepost[si, 1, :, 0].sum() + prepost[si, 2, :, 0].sum())
    return pre, post
