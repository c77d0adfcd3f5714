"""
Seeded synthetic ATAC-seq data of the shapes named in BASELINE.json:
genome (i.i.d. bases, GC 0.41, telomere/centromere-like N blocks + sparse short N runs), peaks
(non-overlapping 500 bp intervals plus a few book-ended / overlapping pairs), and paired-end
alignment records whose Tn5 insertion sites follow a planted sequence preference, with planted
footprints, ~2 % duplicates, ~1 % 5'-soft-clipped records, ~1 % unmapped mates and a tail of
positions carrying more than 10 reads (exercises the cap of atacorrect_functions.py:168).

There is no network in this environment, so this generator is the only data source of the tests
and of bench.py.  Everything is numpy-vectorised; the same seed always yields the same data.
"""
import numpy as np

from .reads import ReadColumns, FLAG_DUPLICATE, FLAG_UNMAPPED, NO_CIGAR

HG38_LENGTHS = {
    "chr1": 248956422, "chr2": 242193529, "chr3": 198295559, "chr4": 190214555, "chr5": 181538259,
    "chr6": 170805979, "chr7": 159345973, "chr8": 145138636, "chr9": 138394717, "chr10": 133797422,
    "chr11": 135086622, "chr12": 133275309, "chr13": 114364328, "chr14": 107043718, "chr15": 101991189,
    "chr16": 90338345, "chr17": 83257441, "chr18": 80373285, "chr19": 58617616, "chr20": 64444167,
    "chr21": 46709983, "chr22": 50818468, "chrX": 156040895, "chrY": 57227415, "chrM": 16569,
}

_ASCII = np.frombuffer(b"ATCGN", dtype=np.uint8)          # reference code order A0 T1 C2 G3 N4 (sequences.pyx:398-408)
_CODE_LUT = np.full(256, 4, dtype=np.uint8)
for _i, _c in enumerate(b"ATCG"):
    _CODE_LUT[_c] = _i
    _CODE_LUT[ord(chr(_c).lower())] = _i


def ascii_to_codes(seq):
    return _CODE_LUT[seq]


def planted_pwm(rng):
    """log2-odds Tn5-like preference over 9 positions centred on the cut (rows A,T,C,G)."""
    pwm = rng.normal(0.0, 0.55, size=(4, 9))
    pwm[:, 4] *= 1.6
    pwm -= pwm.mean(axis=0, keepdims=True)
    return pwm


def make_contig(rng, length, gc=0.41, lower_frac=0.0):
    """ASCII contig + list of (start, end) N blocks."""
    u = rng.integers(0, 10000, size=length, dtype=np.uint16)
    at = int((1.0 - gc) * 5000)
    codes = np.empty(length, dtype=np.uint8)
    codes[:] = 3                                   # G
    codes[u < 5000 + at] = 2                       # C
    codes[u < 2 * at] = 1                          # T
    codes[u < at] = 0                              # A
    del u
    seq = _ASCII[codes]
    blocks = []
    if length >= 20000:
        tel = int(min(10000, max(200, length // 50)))
        blocks.append((0, tel))
        blocks.append((length - tel, length))
        cen = int(length * 0.03) - 2 * tel
        if cen > 0:
            c0 = int(length * 0.4)
            blocks.append((c0, c0 + cen))
        n_small = max(1, length // 1000000) if length >= 200000 else 2
        starts = rng.integers(tel + 1000, max(tel + 1001, length - tel - 1000), size=n_small)
        lens = rng.integers(1, 51, size=n_small)
        for s, ln in zip(starts, lens):
            blocks.append((int(s), int(s + ln)))
    for s, e in blocks:
        seq[s:e] = ord("N")
    if lower_frac > 0 and length > 1000:           # soft-masked stretch (lower case must read as the same base)
        s = int(length * 0.6)
        e = min(length, s + int(length * lower_frac))
        seq[s:e] = np.where(seq[s:e] == ord("N"), seq[s:e], seq[s:e] | 0x20)
    blocks.sort()
    return seq, blocks


def make_genome(seed, lengths, gc=0.41, lower_frac=0.001):
    """dict name -> ASCII uint8 array, dict name -> N blocks."""
    genome, nblocks = {}, {}
    for k, (name, ln) in enumerate(lengths.items()):
        rng = np.random.default_rng([seed, 1, k])
        genome[name], nblocks[name] = make_contig(rng, int(ln), gc, lower_frac)
    return genome, nblocks


def _outside_blocks(pos, width, blocks, margin=0):
    ok = np.ones(pos.shape[0], dtype=bool)
    for s, e in blocks:
        if e - s >= 100:
            ok &= (pos + width <= s - margin) | (pos >= e + margin)
    return ok


def make_peaks(seed, lengths, nblocks, n_peaks, width=500, skip=("chrM",), universe=None):
    """Sorted list of (chrom, start, end): non-overlapping `width`-bp intervals outside the large N blocks,
    plus a few book-ended and overlapping pairs (exercise RegionList.merge, regions.py:343-373).
    universe: the whole workload's name -> length when `lengths` is a shard of it (seeds and shares per contig then do not
    depend on the shard)."""
    universe = universe if universe is not None else lengths
    unames = [n for n in universe if n not in skip]
    names = [n for n in lengths if n not in skip]
    tot = float(sum(universe[n] for n in unames))
    peaks = []
    for name in names:
        k = unames.index(name)
        rng = np.random.default_rng([seed, 2, k])
        ln = lengths[name]
        n = max(1, int(round(n_peaks * ln / tot)))
        if ln < 4 * width:
            continue
        cand = np.sort(rng.integers(width, ln - 2 * width, size=int(n * 1.6) + 8))
        cand = cand[_outside_blocks(cand, width, nblocks[name], margin=300)]
        keep = []
        last_end = -10 ** 9
        for s in cand:
            if s >= last_end + 250:               # leave room so that +/-100 bp extension rarely merges
                keep.append(int(s))
                last_end = s + width
                if len(keep) == n:
                    break
        for s in keep:
            peaks.append((name, s, s + width))
        # a book-ended pair and an overlapping pair per contig (if there is room)
        if len(keep) >= 4:
            s = keep[1]
            if keep[2] - (s + width) > width + 300:
                peaks.append((name, s + width, s + 2 * width))
            s = keep[3]
            if len(keep) > 4 and keep[4] - (s + width) > width + 300:
                peaks.append((name, s + width - 120, s + 2 * width - 120))
    order = {n: i for i, n in enumerate(lengths)}
    peaks.sort(key=lambda t: (order[t[0]], t[1], t[2]))
    return peaks


def write_bed(path, regions):
    with open(path, "w") as fh:
        for r in regions:
            fh.write("%s\t%d\t%d\n" % (r[0], r[1], r[2]))


def _insertion_weight(codes, pos, pwm):
    """Relative Tn5 preference 2^score at 0-based cut positions `pos` (N in the 9-mer -> small constant)."""
    n = codes.shape[0]
    w = np.zeros(pos.shape[0])
    bad = np.zeros(pos.shape[0], dtype=bool)
    for j in range(9):
        idx = np.clip(pos + (j - 4), 0, n - 1)
        c = codes[idx]
        bad |= c > 3
        w += pwm[np.minimum(c, 3), j]
    w = np.minimum(1.0, np.exp2(w - 0.45 * np.abs(pwm).max(axis=0).sum()))
    w[bad] = 0.02
    return w


def make_reads(seed, genome, nblocks, peaks, n_records, read_len=50, frip=0.3, skip=(),
               dup_frac=0.02, clip_frac=0.01, unmapped_frac=0.01, hot_frac=0.0005):
    """Alignment records (ReadColumns, coordinate sorted) for all contigs of `genome`."""
    names = list(genome.keys())
    lengths = [int(genome[n].shape[0]) for n in names]
    tot = float(sum(ln for n, ln in zip(names, lengths) if n not in skip))
    pwm = planted_pwm(np.random.default_rng([seed, 3]))
    peaks_by_chrom = {}
    for c, s, e in peaks:
        peaks_by_chrom.setdefault(c, []).append((s, e))
    parts = []
    for k, name in enumerate(names):
        ln = lengths[k]
        if ln < 2000:
            continue
        share = ln / tot if name not in skip else 0.02 * ln / tot + 0.0
        n_frag = int(round(n_records * share / 2.0)) if name not in skip else max(50, int(n_records * 1e-4))
        if n_frag <= 0:
            continue
        rng = np.random.default_rng([seed, 4, k])
        codes = ascii_to_codes(genome[name])
        pk = np.array(peaks_by_chrom.get(name, []), dtype=np.int64).reshape(-1, 2)
        starts_l, ends_l, rev_l, tl_l = [], [], [], []
        need = n_frag
        rounds = 0
        while need > 0 and rounds < 12:
            rounds += 1
            m = int(need * 3.2) + 64
            in_peak = (rng.random(m) < frip) & (pk.shape[0] > 0)
            pos = rng.integers(300, ln - 300, size=m)
            if pk.shape[0] > 0:
                pi = rng.integers(0, pk.shape[0], size=m)
                ppos = pk[pi, 0] + (rng.random(m) * (pk[pi, 1] - pk[pi, 0])).astype(np.int64)
                pos = np.where(in_peak, ppos, pos)
            anchor_left = rng.random(m) < 0.5
            comp = rng.random(m)
            flen = np.where(comp < 0.5, rng.normal(75, 18, m), np.where(comp < 0.85, rng.normal(200, 30, m),
                                                                       rng.normal(400, 45, m)))
            flen = np.clip(flen, read_len + 2, 800).astype(np.int64)
            w = _insertion_weight(codes, pos, pwm)
            ok = rng.random(m) < w
            ok &= _outside_blocks(pos - 450, 900, nblocks[name])
            pos, anchor_left, flen = pos[ok][:need], anchor_left[ok][:need], flen[ok][:need]
            left_cut = np.where(anchor_left, pos, pos - flen)
            right_cut = left_cut + flen
            good = (left_cut > 100) & (right_cut < ln - 100)
            left_cut, right_cut, flen = left_cut[good], right_cut[good], flen[good]
            # + mate: 0-based cut index = ref_start + 4 ; - mate: cut index = ref_end - 5   (ngs.pyx:91-106)
            fs = left_cut - 4
            re_ = right_cut + 5
            starts_l += [fs, re_ - read_len]
            ends_l += [fs + read_len, re_]
            rev_l += [np.zeros(fs.shape[0], dtype=bool), np.ones(fs.shape[0], dtype=bool)]
            tl_l += [flen, -flen]
            need -= fs.shape[0]
        if not starts_l:
            continue
        rs = np.concatenate(starts_l).astype(np.int64)
        ren = np.concatenate(ends_l).astype(np.int64)
        rv = np.concatenate(rev_l)
        tl = np.concatenate(tl_l)
        # planted footprints: deplete cuts inside 20-40 bp windows at random peak positions
        if pk.shape[0] > 0:
            nfp = max(1, pk.shape[0] * 2)
            fpi = rng.integers(0, pk.shape[0], size=nfp)
            fps = pk[fpi, 0] + rng.integers(40, 420, size=nfp)
            fpw = rng.integers(20, 41, size=nfp)
            cut = np.where(rv, ren - 5, rs + 4)
            order = np.argsort(fps)
            fps, fpw = fps[order], fpw[order]
            j = np.clip(np.searchsorted(fps, cut, side="right") - 1, 0, nfp - 1)
            inside = (cut >= fps[j]) & (cut < fps[j] + fpw[j])
            keep = ~(inside & (rng.random(cut.shape[0]) < 0.85))
            rs, ren, rv, tl = rs[keep], ren[keep], rv[keep], tl[keep]
        n = rs.shape[0]
        # hot positions: pile-ups of > 10 identical cut sites (cap path)
        n_hot = max(2, int(n * hot_frac))
        hi = rng.integers(0, n, size=n_hot)
        reps = rng.integers(11, 41, size=n_hot)
        h_idx = np.repeat(hi, reps)
        jitter = rng.integers(0, 30, size=h_idx.shape[0])
        h_rs = np.where(rv[h_idx], rs[h_idx] - jitter, rs[h_idx])
        h_re = np.where(rv[h_idx], ren[h_idx], ren[h_idx] + jitter)
        rs = np.concatenate([rs, h_rs])
        ren = np.concatenate([ren, h_re])
        rv = np.concatenate([rv, rv[h_idx]])
        tl = np.concatenate([tl, tl[h_idx]])
        n = rs.shape[0]
        flag = np.where(rv, 147, 99).astype(np.uint16)
        left_soft = np.zeros(n, dtype=np.int32)
        right_soft = np.zeros(n, dtype=np.int32)
        first_op = np.zeros(n, dtype=np.uint8)
        last_op = np.zeros(n, dtype=np.uint8)
        qlen = (ren - rs).astype(np.int32)
        # soft clips: mostly at the 5' end (moves ref_start/ref_end inwards, cut site unchanged), some 3'
        clip = rng.random(n) < clip_frac
        amount = rng.integers(1, 13, size=n).astype(np.int32)
        five = rng.random(n) < 0.7
        c5f = clip & five & ~rv
        c5r = clip & five & rv
        c3f = clip & ~five & ~rv
        c3r = clip & ~five & rv
        left_soft[c5f | c3r] = amount[c5f | c3r]
        right_soft[c5r | c3f] = amount[c5r | c3f]
        rs = rs + left_soft
        ren = ren - right_soft
        first_op[left_soft > 0] = 4
        last_op[right_soft > 0] = 4
        # a few hard-clipped records (first op H) and records without CIGAR information
        hard = rng.random(n) < clip_frac * 0.1
        first_op[hard & (left_soft == 0)] = 5
        flag[rng.random(n) < dup_frac] |= FLAG_DUPLICATE
        unm = rng.random(n) < unmapped_frac
        flag[unm] |= FLAG_UNMAPPED
        first_op[unm] = NO_CIGAR
        last_op[unm] = NO_CIGAR
        ren = np.where(unm, rs + 1, ren)
        parts.append(ReadColumns(names, lengths, chrom_id=np.full(n, k, dtype=np.int32), ref_start=rs, ref_end=ren,
                                 left_soft=np.where(unm, 0, left_soft), right_soft=np.where(unm, 0, right_soft),
                                 query_length=qlen, flag=flag, first_op=first_op, last_op=last_op, tlen=tl))
    rc = ReadColumns.concat(parts) if parts else ReadColumns(names, lengths, **{
        f: np.zeros(0) for f in ("chrom_id", "ref_start", "ref_end", "left_soft", "right_soft", "query_length",
                                  "flag", "first_op", "last_op", "tlen")})
    return rc.sort()


class SynthDataset:
    """Genome + peaks + reads of one configuration."""

    def __init__(self, seed, lengths, n_peaks, n_records, skip=("chrM",)):
        self.seed = seed
        self.lengths = dict(lengths)
        self.genome, self.nblocks = make_genome(seed, self.lengths)
        self.peaks = make_peaks(seed, self.lengths, self.nblocks, n_peaks, skip=skip)
        self.reads = make_reads(seed, self.genome, self.nblocks, self.peaks, n_records)

    def write(self, outdir, prefix="synth", bam=True):
        import os
        from .io.fasta import write_fasta
        os.makedirs(outdir, exist_ok=True)
        fa = os.path.join(outdir, prefix + ".fa")
        bed = os.path.join(outdir, prefix + "_peaks.bed")
        write_fasta(fa, self.genome.items())
        write_bed(bed, self.peaks)
        npz = os.path.join(outdir, prefix + "_reads.npz")
        self.reads.save(npz)
        out = {"fasta": fa, "peaks": bed, "reads_npz": npz}
        if bam:
            from .io.bam import write_bam
            out["bam"] = os.path.join(outdir, prefix + ".bam")
            write_bam(out["bam"], self.reads)
        return out


def small_config(seed=7, n_contigs=2, contig_len=120000, n_peaks=24, n_records=60000):
    lengths = {"chr%d" % (i + 1): contig_len - 7001 * i for i in range(n_contigs)}
    lengths["chrM"] = 6000
    return SynthDataset(seed, lengths, n_peaks, n_records)
