"""
GPU-accelerated variant of tobias_b200.synth for the benchmark-sized configurations (hg38-shaped: 3.1 Gb, 150 k peaks,
50 M alignment records).  Same recipe as synth.make_genome / make_reads -- i.i.d. bases with N blocks, insertion sites
sampled proportionally to 2^(planted 9-bp preference), nucleosomal fragment-length mixture, FRiP ~0.3, planted footprints,
duplicates / soft clips / unmapped mates / >10-read pile-ups -- but every per-base and per-read step is a torch tensor
op, so the data set is built in seconds on the GPU (torch is plumbing here: random numbers and sorting; nothing of the
hot path).  Works on CPU tensors too (slower), which is what the --impl reference arm uses for its bounded sample.
Results are returned as host numpy arrays (the hot path is fed from host buffers).
"""
import numpy as np
import torch

from . import synth
from .reads import ReadColumns, FLAG_DUPLICATE, FLAG_UNMAPPED, NO_CIGAR


def _gen(device, *key):
    g = torch.Generator(device=device)
    g.manual_seed(int(np.random.SeedSequence(list(key)).generate_state(1)[0]))
    return g


def make_contig(seed, k, length, device, gc=0.41):
    g = _gen(device, seed, 1, k)
    u = torch.randint(0, 10000, (length,), generator=g, device=device, dtype=torch.int16)
    at = int((1.0 - gc) * 5000)
    codes = torch.full((length,), 3, dtype=torch.uint8, device=device)
    codes[u < 5000 + at] = 2
    codes[u < 2 * at] = 1
    codes[u < at] = 0
    del u
    blocks = []
    if length >= 20000:
        tel = int(min(10000, max(200, length // 50)))
        blocks += [(0, tel), (length - tel, length)]
        cen = int(length * 0.03) - 2 * tel
        if cen > 0:
            c0 = int(length * 0.4)
            blocks.append((c0, c0 + cen))
        rng = np.random.default_rng([seed, 5, k])
        n_small = max(1, length // 1000000) if length >= 200000 else 2
        for s, ln in zip(rng.integers(tel + 1000, max(tel + 1001, length - tel - 1000), size=n_small), rng.integers(1, 51, size=n_small)):
            blocks.append((int(s), int(s + ln)))
    for s, e in blocks:
        codes[s:e] = 4
    blocks.sort()
    return codes, blocks


def _outside(pos, width, blocks):
    ok = torch.ones_like(pos, dtype=torch.bool)
    for s, e in blocks:
        if e - s >= 100:
            ok &= (pos + width <= s) | (pos >= e)
    return ok


def make_reads_contig(seed, k, codes, blocks, pk, n_frag, pwm, device, read_len=50, frip=0.3, dup_frac=0.02, clip_frac=0.01,
                      unmapped_frac=0.01, hot_frac=0.0005):
    """Columns (torch tensors on `device`) of one contig, sorted by ref_start.  `seed` only seeds the reads (a second sample of the same
    genome and peaks: another seed)."""
    g = _gen(device, seed, 4, k)
    ln = codes.shape[0]
    pwm_t = torch.tensor(pwm, dtype=torch.float32, device=device)
    wmax = float(np.abs(pwm).max(axis=0).sum())
    pk_t = torch.tensor(np.asarray(pk, dtype=np.int64).reshape(-1, 2), device=device)
    rnd = lambda n: torch.rand(n, generator=g, device=device)
    rint = lambda lo, hi, n: torch.randint(lo, hi, (n,), generator=g, device=device)
    rs_l, re_l, rv_l, tl_l = [], [], [], []
    need, rounds = n_frag, 0
    while need > 0 and rounds < 12:
        rounds += 1
        m = int(need * 3.2) + 64
        pos = rint(300, ln - 300, m)
        if pk_t.shape[0] > 0:
            in_peak = rnd(m) < frip
            pi = rint(0, pk_t.shape[0], m)
            ppos = pk_t[pi, 0] + (rnd(m) * (pk_t[pi, 1] - pk_t[pi, 0]).float()).long()
            pos = torch.where(in_peak, ppos, pos)
        anchor_left = rnd(m) < 0.5
        comp = rnd(m)
        nrm = torch.randn(m, generator=g, device=device)
        flen = torch.where(comp < 0.5, 75 + 18 * nrm, torch.where(comp < 0.85, 200 + 30 * nrm, 400 + 45 * nrm))
        flen = flen.clamp(read_len + 2, 800).long()
        w = torch.zeros(m, device=device)
        bad = torch.zeros(m, dtype=torch.bool, device=device)
        for j in range(9):
            c = codes[(pos + (j - 4)).clamp(0, ln - 1)].long()
            bad |= c > 3
            w += pwm_t[c.clamp(max=3), j]
        w = torch.exp2(w - 0.45 * wmax).clamp(max=1.0)
        w[bad] = 0.02
        ok = (rnd(m) < w) & _outside(pos - 450, 900, blocks)
        pos, anchor_left, flen = pos[ok][:need], anchor_left[ok][:need], flen[ok][:need]
        left = torch.where(anchor_left, pos, pos - flen)
        right = left + flen
        good = (left > 100) & (right < ln - 100)
        left, right, flen = left[good], right[good], flen[good]
        fs, re_ = left - 4, right + 5
        rs_l += [fs, re_ - read_len]
        re_l += [fs + read_len, re_]
        rv_l += [torch.zeros_like(fs, dtype=torch.bool), torch.ones_like(fs, dtype=torch.bool)]
        tl_l += [flen, -flen]
        need -= fs.shape[0]
    rs, ren, rv, tl = torch.cat(rs_l), torch.cat(re_l), torch.cat(rv_l), torch.cat(tl_l)
    if pk_t.shape[0] > 0:                                        # planted footprints
        nfp = pk_t.shape[0] * 2
        fpi = rint(0, pk_t.shape[0], nfp)
        fps = pk_t[fpi, 0] + rint(40, 420, nfp)
        fpw = rint(20, 41, nfp)
        fps, order = torch.sort(fps)
        fpw = fpw[order]
        cut = torch.where(rv, ren - 5, rs + 4)
        j = (torch.searchsorted(fps, cut, right=True) - 1).clamp(0, nfp - 1)
        inside = (cut >= fps[j]) & (cut < fps[j] + fpw[j])
        keep = ~(inside & (rnd(cut.shape[0]) < 0.85))
        rs, ren, rv, tl = rs[keep], ren[keep], rv[keep], tl[keep]
    n = rs.shape[0]
    n_hot = max(2, int(n * hot_frac))                            # > 10 reads on one cut site
    hi = rint(0, n, n_hot)
    reps = rint(11, 41, n_hot)
    h_idx = torch.repeat_interleave(hi, reps)
    jit = rint(0, 30, h_idx.shape[0])
    rs = torch.cat([rs, torch.where(rv[h_idx], rs[h_idx] - jit, rs[h_idx])])
    ren = torch.cat([ren, torch.where(rv[h_idx], ren[h_idx], ren[h_idx] + jit)])
    rv = torch.cat([rv, rv[h_idx]])
    tl = torch.cat([tl, tl[h_idx]])
    n = rs.shape[0]
    flag = torch.where(rv, 147, 99).to(torch.int32)
    qlen = (ren - rs).to(torch.int32)
    clip = rnd(n) < clip_frac
    amount = rint(1, 13, n).to(torch.int32)
    five = rnd(n) < 0.7
    zero = torch.zeros(n, dtype=torch.int32, device=device)
    left_soft = torch.where(clip & ((five & ~rv) | (~five & rv)), amount, zero)
    right_soft = torch.where(clip & ((five & rv) | (~five & ~rv)), amount, zero)
    rs = rs + left_soft
    ren = ren - right_soft
    first_op = torch.where(left_soft > 0, 4, 0).to(torch.uint8)
    last_op = torch.where(right_soft > 0, 4, 0).to(torch.uint8)
    hard = (rnd(n) < clip_frac * 0.1) & (left_soft == 0)
    first_op[hard] = 5
    flag = flag | torch.where(rnd(n) < dup_frac, FLAG_DUPLICATE, 0).to(torch.int32)
    unm = rnd(n) < unmapped_frac
    flag = flag | torch.where(unm, FLAG_UNMAPPED, 0).to(torch.int32)
    first_op[unm] = NO_CIGAR
    last_op[unm] = NO_CIGAR
    ren = torch.where(unm, rs + 1, ren)
    left_soft = torch.where(unm, zero, left_soft)
    right_soft = torch.where(unm, zero, right_soft)
    order = torch.argsort(rs, stable=True)
    cols = dict(ref_start=rs, ref_end=ren, left_soft=left_soft, right_soft=right_soft, query_length=qlen, flag=flag,
                first_op=first_op, last_op=last_op, tlen=tl)
    return {k_: v[order] for k_, v in cols.items()}


class TorchDataset:
    """names, genome (list of uint8 ASCII numpy arrays), peaks [(name, s, e)], reads (ReadColumns)."""

    def __init__(self, seed, lengths, n_peaks, n_records, device=None, skip=("chrM",), pin=False, universe=None, keep_codes=False):
        """universe: name -> length of the WHOLE workload when `lengths` is one rank's shard of it.  Every contig is then seeded by its
        index in the universe and receives its share of the universe's peaks / records, so a contig holds the same bases, peaks and
        records whichever rank builds it (n_peaks / n_records are the universe's totals in that case)."""
        device = device or ("cuda" if torch.cuda.is_available() else "cpu")
        self.names = list(lengths.keys())
        self.lengths = [int(lengths[n]) for n in self.names]
        universe = dict(universe) if universe is not None else dict(lengths)
        key = {n: i for i, n in enumerate(universe)}
        ascii_lut = torch.tensor(list(b"ATCGN"), dtype=torch.uint8, device=device)
        pwm = synth.planted_pwm(np.random.default_rng([seed, 3]))
        tot = float(sum(ln for n, ln in universe.items() if n not in skip))
        codes_dev, nblocks = {}, {}
        self.genome = []
        for name, ln in zip(self.names, self.lengths):
            k = key[name]
            codes, blocks = make_contig(seed, k, ln, device)
            nblocks[name] = blocks
            codes_dev[name] = codes
            host = torch.empty(ln, dtype=torch.uint8, pin_memory=pin and torch.cuda.is_available())
            host.copy_(ascii_lut[codes.long()] if device == "cpu" else torch.take(ascii_lut, codes.long()))
            self.genome.append(host.numpy())
        self.nblocks = nblocks
        self.peaks = synth.make_peaks(seed, dict(zip(self.names, self.lengths)), nblocks, n_peaks, skip=skip, universe=universe)
        self._state = (key, tot, codes_dev, pwm, device, skip)
        self.reads = self.reads_for(seed, n_records, drop_codes=not keep_codes)
        if not keep_codes:
            self._state = None
        if device != "cpu":
            torch.cuda.empty_cache()

    def reads_for(self, reads_seed, n_records, drop_codes=False):
        """Another sample of the same genome and peaks: `n_records` alignment records drawn with `reads_seed` (needs keep_codes=True
        for calls after construction)."""
        key, tot, codes_dev, pwm, device, skip = self._state
        by_chrom = {}
        for c, s, e in self.peaks:
            by_chrom.setdefault(c, []).append((s, e))
        parts = []
        for ki, (name, ln) in enumerate(zip(self.names, self.lengths)):
            k = key[name]
            if ln < 2000:
                continue
            n_frag = int(round(n_records * (ln / tot) / 2.0)) if name not in skip else max(50, int(n_records * 1e-4))
            if n_frag <= 0:
                continue
            cols = make_reads_contig(reads_seed, k, codes_dev[name], self.nblocks[name], by_chrom.get(name, []), n_frag, pwm, device)
            n = cols["ref_start"].shape[0]
            parts.append(ReadColumns(self.names, self.lengths, chrom_id=np.full(n, ki, dtype=np.int32),
                                     **{f: v.cpu().numpy() for f, v in cols.items()}))
            if drop_codes:
                del codes_dev[name]
        return ReadColumns.concat(parts)
