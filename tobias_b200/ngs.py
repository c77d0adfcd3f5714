"""
Array-level mirror of the reference's read containers (tobias/utils/ngs.pyx:55-199) for the per-base path: a `ReadList` is a set of
alignment records held as columns (tobias_b200.reads.ReadColumns) instead of a Python list of OneRead objects; `from_bam`, `split_strands`
and `signal` keep the reference's semantics, and `signal` runs on the GPU through the C-ABI (tb_read_signal).  There is no CPU fallback.
"""
import numpy as np

from .reads import ReadColumns


class ReadList:
    def __init__(self, columns=None, read_shift=(4, -5)):
        self.columns = columns
        self.read_shift = tuple(read_shift)

    def __len__(self):
        return 0 if self.columns is None else len(self.columns)

    def from_bam(self, bam_obj, region):
        """bam_obj: ReadColumns, or a path to a .bam / .npz of records.  Records whose alignment overlaps the region (pysam fetch), minus
        unmapped and duplicate ones (ngs.pyx:150-160).  Returns a new ReadList, like the reference."""
        if isinstance(bam_obj, str):
            if bam_obj.endswith(".npz"):
                bam_obj = ReadColumns.load(bam_obj)
            else:
                from .io.bam import read_bam
                bam_obj = read_bam(bam_obj)
        cid = bam_obj.references.index(region.chrom)
        idx = bam_obj.fetch_index(cid, region.start, region.end)
        sub = bam_obj.select(idx)
        return ReadList(sub.select(sub.usable()), self.read_shift)

    def get_cutsites(self, read_shift):
        """OneRead.get_cutsite for every record (ngs.pyx:91-106): remembers the shift `signal` applies."""
        self.read_shift = tuple(read_shift)
        return self

    def split_strands(self):
        rev = self.columns.is_reverse
        return [ReadList(self.columns.select(~rev), self.read_shift), ReadList(self.columns.select(rev), self.read_shift)]

    def signal(self, region, ctx=None):
        """values[cut - start - 1] += 1 for every record of the list with start < cut <= end (ngs.pyx:186-199); float64[reg_len]."""
        from . import lib, signals
        n = region.end - region.start
        if self.columns is None or len(self.columns) == 0 or n <= 0:
            return np.zeros(max(n, 0))
        ctx = ctx or signals._default_ctx()
        rc = self.columns
        ctx.genome_create(rc.lengths)
        ctx.reads_upload_columns(rc)
        fwd, rev = ctx.pileup([rc.references.index(region.chrom)], [region.start], [region.end], self.read_shift, list_semantics=True)
        return (fwd.astype(np.float64) + rev.astype(np.float64))
