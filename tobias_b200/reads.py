"""
Columnar alignment records ("reads") -- the host-side form of what the reference keeps as a list of
`OneRead` objects (tobias/utils/ngs.pyx:55-89).  Only the attributes the hot path consumes are kept:

    reference_start / reference_end            ngs.pyx:78-79    -> ref_start / ref_end (0-based, half-open)
    query_alignment_start / _end, query_length ngs.pyx:80-84    -> left_soft / right_soft / query_length
    is_reverse, is_unmapped, is_duplicate      ngs.pyx:83,153   -> flag bits 0x10 / 0x4 / 0x400
    cigartuples[0][0], cigartuples[-1][0]      atacorrect_functions.py:156-158 -> first_op / last_op

Records are sorted by (chrom_id, ref_start) -- BAM coordinate order -- so that a region fetch is a
binary search (htslib overlap semantics: [ref_start, ref_end) intersects [start, end)).
"""
import numpy as np

FLAG_UNMAPPED = 0x4
FLAG_REVERSE = 0x10
FLAG_DUPLICATE = 0x400
NO_CIGAR = 255

_FIELDS = [("chrom_id", np.int32), ("ref_start", np.int32), ("ref_end", np.int32),
           ("left_soft", np.int32), ("right_soft", np.int32), ("query_length", np.int32),
           ("flag", np.uint16), ("first_op", np.uint8), ("last_op", np.uint8), ("tlen", np.int32)]


class ReadColumns:
    """Struct-of-arrays alignment records for a set of contigs."""

    def __init__(self, references, lengths, **cols):
        self.references = list(references)
        self.lengths = [int(x) for x in lengths]
        n = None
        for name, dt in _FIELDS:
            arr = np.ascontiguousarray(cols[name], dtype=dt)
            if n is None:
                n = arr.shape[0]
            if arr.shape != (n,):
                raise ValueError("column %s has shape %s, expected (%d,)" % (name, arr.shape, n))
            setattr(self, name, arr)
        self.n = int(n)
        self._chrom_bounds = None
        self._max_span = None

    # ------------------------------------------------------------------ basic
    def __len__(self):
        return self.n

    def sort(self):
        order = np.lexsort((self.ref_start, self.chrom_id))
        if not np.array_equal(order, np.arange(self.n)):
            for name, _ in _FIELDS:
                setattr(self, name, np.ascontiguousarray(getattr(self, name)[order]))
        self._chrom_bounds = None
        return self

    def is_sorted(self):
        if self.n < 2:
            return True
        key = self.chrom_id.astype(np.int64) << 32 | self.ref_start.astype(np.int64)
        return bool(np.all(key[1:] >= key[:-1]))

    def select(self, mask_or_index):
        cols = {name: getattr(self, name)[mask_or_index] for name, _ in _FIELDS}
        return ReadColumns(self.references, self.lengths, **cols)

    @property
    def chrom_bounds(self):
        """[n_chroms+1] offsets of each contig's records (records are sorted by chrom_id)."""
        if self._chrom_bounds is None:
            self._chrom_bounds = np.searchsorted(self.chrom_id, np.arange(len(self.references) + 1), side="left")
        return self._chrom_bounds

    @property
    def max_span(self):
        if self._max_span is None:
            self._max_span = int((self.ref_end - self.ref_start).max()) if self.n else 0
        return self._max_span

    # ------------------------------------------------------------------ derived
    @property
    def is_reverse(self):
        return (self.flag & FLAG_REVERSE) != 0

    def usable(self):
        """The filter of ReadList.from_bam (ngs.pyx:153): drop unmapped and duplicate records, nothing else."""
        return (self.flag & (FLAG_UNMAPPED | FLAG_DUPLICATE)) == 0

    def clip5(self):
        """Soft-clipped bases at the 5' end of the read (ngs.pyx:98): left clip for +, right clip for -."""
        return np.where(self.is_reverse, self.right_soft, self.left_soft).astype(np.int32)

    def five_prime_is_match(self):
        """5'-most CIGAR op is M (atacorrect_functions.py:156-158)."""
        op = np.where(self.is_reverse, self.last_op, self.first_op)
        return op == 0

    def fetch_index(self, chrom_id, start, end):
        """Indices of records overlapping [start, end) on contig `chrom_id`, in file order (pysam fetch)."""
        lo, hi = self.chrom_bounds[chrom_id], self.chrom_bounds[chrom_id + 1]
        if hi <= lo:
            return np.zeros(0, dtype=np.int64)
        rs = self.ref_start[lo:hi]
        a = np.searchsorted(rs, start - self.max_span, side="left")
        b = np.searchsorted(rs, end, side="left")
        idx = np.arange(lo + a, lo + b, dtype=np.int64)
        if idx.size:
            idx = idx[self.ref_end[idx] > start]
        return idx

    # ------------------------------------------------------------------ I/O
    def save(self, path):
        np.savez(path, references=np.array(self.references), lengths=np.array(self.lengths, dtype=np.int64),
                 **{name: getattr(self, name) for name, _ in _FIELDS})

    @staticmethod
    def load(path):
        z = np.load(path, allow_pickle=False)
        return ReadColumns([str(x) for x in z["references"]], z["lengths"],
                           **{name: z[name] for name, _ in _FIELDS})

    @staticmethod
    def concat(parts):
        first = parts[0]
        cols = {name: np.concatenate([getattr(p, name) for p in parts]) for name, _ in _FIELDS}
        return ReadColumns(first.references, first.lengths, **cols)
