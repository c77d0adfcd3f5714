"""
TEST INFRASTRUCTURE ONLY -- stand-in for `pyBigWig` (absent from the image) for the compiled
reference in oracle/_ref.  Backed by the repo's own bigWig codec (tobias_b200/io/bigwig.py).
Covers what the reference calls: open(f[, "w"]), chroms(), values(chrom,s,e,numpy=True),
addHeader(list), addEntries(chrom, pos_list, values=list, span=1), close(); attribute numpy=1
(tobias/utils/regions.py:168; tobias/utils/utilities.py:141-215; tobias/tools/score_bigwig.py:37-39).
"""
import numpy as np

numpy = 1


def _codec():
    from tobias_b200.io import bigwig
    return bigwig


class _Reader:
    def __init__(self, path):
        self._r = _codec().BigWigReader(path)

    def chroms(self, chrom=None):
        d = dict(self._r.chroms)
        return d if chrom is None else d.get(chrom)

    def values(self, chrom, start, end, numpy=False):
        v = self._r.values(chrom, int(start), int(end))
        return v if numpy else v.tolist()

    def close(self):
        self._r.close()


class _Writer:
    def __init__(self, path):
        self._w = _codec().BigWigWriter(path)

    def addHeader(self, header, maxZooms=None):
        self._w.add_header(header)

    def addEntries(self, chrom, starts, values=None, span=1, ends=None, step=None, validate=True):
        if isinstance(chrom, (list, tuple)) or ends is not None or step is not None or span != 1:
            raise NotImplementedError("shim supports addEntries(chrom, positions, values=..., span=1) only")
        self._w.add_entries(chrom, np.asarray(starts, dtype=np.int64), np.asarray(values, dtype=np.float32))

    def close(self):
        self._w.close()


def open(path, mode="r"):
    return _Writer(path) if "w" in mode else _Reader(path)
