"""
TEST INFRASTRUCTURE ONLY -- functional stand-in for `pysam` (absent from the image) so that the
compiled, UNMODIFIED reference in oracle/_ref can run.  It provides exactly the attribute set the
reference's hot path touches (tobias/utils/ngs.pyx:72-87,153; tobias/tools/atacorrect.py:131-146;
tobias/utils/sequences.pyx:388):

  AlignmentFile(f, mode): .references .lengths .has_index() .fetch(chrom,start,end) .close()
  read: reference_name query_name reference_start reference_end query_alignment_start
        query_alignment_end template_length is_reverse query_length infer_query_length() flag
        cigartuples get_tags() is_unmapped is_duplicate
  FastaFile(f): .references .lengths .fetch(chrom,start,end) .close()

`fetch` follows htslib overlap semantics ([ref_start, ref_end) intersects [start, end), file order).
Alignment records come from the repo's columnar container (`.npz`) or its BAM codec (`.bam`).
"""
import os

from tobias_b200.reads import ReadColumns, NO_CIGAR, FLAG_UNMAPPED, FLAG_REVERSE, FLAG_DUPLICATE
from tobias_b200.io.fasta import FastaFile as _Fasta

__version__ = "shim"
_cache = {}


class AlignedSegment:
    __slots__ = ("reference_name", "query_name", "reference_start", "reference_end", "query_alignment_start",
                 "query_alignment_end", "template_length", "is_reverse", "query_length", "flag", "cigartuples",
                 "is_unmapped", "is_duplicate", "_qlen")

    def infer_query_length(self):
        return self._qlen

    def get_tags(self):
        return []


def _make_read(rc, i):
    r = AlignedSegment()
    flag = int(rc.flag[i])
    start, end = int(rc.ref_start[i]), int(rc.ref_end[i])
    ls, rs, ql = int(rc.left_soft[i]), int(rc.right_soft[i]), int(rc.query_length[i])
    r.reference_name = rc.references[int(rc.chrom_id[i])]
    r.query_name = "r%d" % i
    r.reference_start = start
    r.reference_end = end
    r.query_alignment_start = ls
    r.query_alignment_end = ql - rs
    r.template_length = int(rc.tlen[i])
    r.is_reverse = bool(flag & FLAG_REVERSE)
    r.is_unmapped = bool(flag & FLAG_UNMAPPED)
    r.is_duplicate = bool(flag & FLAG_DUPLICATE)
    r.query_length = ql
    r._qlen = ql
    r.flag = flag
    fo, lo = int(rc.first_op[i]), int(rc.last_op[i])
    if fo == NO_CIGAR:
        r.cigartuples = None
    else:
        # only the first/last operation and the clip lengths are consumed by the reference
        mid = max(ql - ls - rs, 1)
        tup = []
        if fo != 0:
            tup.append((fo, ls if fo == 4 else 1))
        tup.append((0, mid))
        if lo != 0:
            tup.append((lo, rs if lo == 4 else 1))
        r.cigartuples = tup
    return r


def _load_columns(path):
    key = (os.path.abspath(path), os.path.getmtime(path))
    if key not in _cache:
        if path.endswith(".npz"):
            rc = ReadColumns.load(path)
        else:
            from tobias_b200.io.bam import read_bam
            rc = read_bam(path)
        _cache.clear()
        _cache[key] = rc
    return _cache[key]


class AlignmentFile:
    def __init__(self, filename, mode="rb", **kw):
        self.filename = filename.encode() if isinstance(filename, str) else filename
        self._rc = _load_columns(filename if isinstance(filename, str) else filename.decode())
        self.references = tuple(self._rc.references)
        self.lengths = tuple(self._rc.lengths)
        self._name2id = {n: i for i, n in enumerate(self.references)}

    def has_index(self):
        return True

    def check_index(self):
        return True

    def fetch(self, chrom=None, start=None, end=None, **kw):
        rc = self._rc
        cid = self._name2id[chrom]
        start = 0 if start is None else max(0, int(start))
        end = self.lengths[cid] if end is None else int(end)
        for i in rc.fetch_index(cid, start, end):
            yield _make_read(rc, int(i))

    def close(self):
        pass


class FastaFile(_Fasta):
    pass


def index(*args, **kw):
    return None
