"""
Genomic interval handling for the two drivers (host side).

Behavioural mirror of the subset of tobias/utils/regions.py the hot path's drivers use
(`from_bed` :212-253, `loc_sort` :288-302, `merge` :343-373, `subtract` :433-483, `extend_reg` :73-82,
`check_boundary` :118-161, `split_region` :104-115, `chunks` :330-340, `keep_chroms` :389-400), written
from the documented behaviour rather than translated.  Quirks that change results are kept and
flagged with QUIRK.  Coordinates are BED-style: 0-based start, exclusive end.
"""
import math
import re
import sys

import numpy as np


class OneRegion(list):
    """[chrom, start, end, (name, score, strand, ...)] with attribute access, like the reference's OneRegion."""

    def __init__(self, lst=("", 0, 0)):
        super().__init__(lst)
        self.chrom = lst[0]
        self.start = int(lst[1])
        self.end = int(lst[2])
        self.name = lst[3] if len(lst) > 3 else ""
        self.score = lst[4] if len(lst) > 4 else ""
        self.strand = lst[5] if len(lst) > 5 else "."

    def __str__(self):
        return "\t".join(str(x) for x in self)

    def tup(self):
        return (self.chrom, self.start, self.end, self.name, self.strand)

    def _sync(self):
        self[1], self[2] = self.start, self.end

    def get_length(self):
        return self.end - self.start

    get_width = get_length

    def extend_reg(self, bp):
        """In place (and returns self, which callers rely on -- regions.py:73-82)."""
        self.start -= bp
        self.end += bp
        self._sync()
        return self

    def split_region(self, max_length):
        out = RegionList()
        s = self.start
        while s < self.end:
            out.append(OneRegion([self.chrom, s, min(s + max_length, self.end)]))
            s += max_length
        return out

    def check_boundary(self, boundaries, action="cut", logger=None):
        """Clip to [0, len(chrom)]; returns None when the contig is unknown or nothing is left."""
        if self.chrom not in boundaries:
            if action == "exit":
                sys.exit("Chromosome for region \"{0}\" is not found in list of available chromosomes".format(self))
            return None
        limit = int(boundaries[self.chrom])
        if self.start < 0 or self.end > limit:
            if action == "cut":
                self.start = max(0, self.start)
                self.end = min(limit, self.end)
                self._sync()
                if self.get_length() <= 0:
                    return None
            elif action == "remove":
                return None
            elif action == "exit":
                sys.exit("Region \"{0}\" is outside of the chromosome boundaries".format(self))
        return self


_BED_LINE = re.compile(r"[^\s]+\t\d+\t\d+\b.*")


class RegionList(list):
    def __init__(self, lst=()):
        super().__init__(lst)

    def __str__(self):
        return "\n".join(str(r) for r in self)

    # ---------------------------------------------------------------- I/O
    def from_list(self, lst):
        self.extend(lst)
        return self

    def from_bed(self, path, logger=None):
        out = RegionList()
        try:
            with open(path) as fh:
                lines = fh.readlines()
        except UnicodeDecodeError as e:
            sys.exit("Could not read input file '{0}' (compressed?). TOBIAS only reads .bed-files in raw text "
                     "format. Exception is: '{1}'".format(path, e))
        for i, line in enumerate(lines):
            if line.startswith("#"):
                continue
            if _BED_LINE.match(line) is None:
                sys.exit("Line {0} in {1} is not proper bed format:\n{2}".format(i + 1, path, line))
            cols = line.split("\t")
            cols[-1] = cols[-1].rstrip()
            cols[1], cols[2] = int(cols[1]), int(cols[2])
            if cols[1] >= cols[2]:
                sys.exit("Start position is larger than end position in line {0} in {1}:\n{2}".format(i + 1, path, line))
            out.append(OneRegion(cols))
        return out

    def from_chrom_lengths(self, chrom_lengths):
        for chrom in sorted(chrom_lengths):
            self.append(OneRegion([chrom, 0, chrom_lengths[chrom]]))
        return self

    def as_bed(self):
        return "".join("\t".join(str(c) for c in r) + "\n" for r in self)

    def write_bed(self, path):
        with open(path, "w") as fh:
            fh.write(self.as_bed())

    # ---------------------------------------------------------------- ordering / selection
    def count(self):
        return len(self)

    def loc_sort(self, contig_list=()):
        if len(contig_list) > 0:
            order = {c: i for i, c in enumerate(contig_list)}
            self.sort(key=lambda r: (order.get(r.chrom, -1), r.start, r.end, r.name))
        else:
            self.sort(key=lambda r: (r.chrom, r.start, r.end, r.name))

    def get_chroms(self):
        return list(dict.fromkeys(r.chrom for r in self))

    def keep_chroms(self, chroms):
        keep = set(chroms)
        self[:] = [r for r in self if r.chrom in keep]
        return self

    def remove_chroms(self, chroms):
        drop = set(chroms)
        self[:] = [r for r in self if r.chrom not in drop]
        return self

    def chunks(self, n):
        if len(self) == 0:
            return [self]
        per = int(math.ceil(len(self) / float(n)))
        return [RegionList(self[i:i + per]) for i in range(0, len(self), per)]

    def apply_method(self, method, *args):
        out = RegionList()
        for r in self:
            res = method(r, *args)
            if isinstance(res, RegionList):
                out.extend(res)
            elif isinstance(res, OneRegion):
                out.append(res)
        return out

    # ---------------------------------------------------------------- set algebra
    def merge(self):
        """Merge overlapping (start < previous end; book-ended intervals stay separate) regions in place.
        QUIRK (regions.py:361-366): the merged interval ends at the LATER region's end, not at the maximum
        of the two ends, so an interval nested inside its predecessor shortens the merged interval."""
        self.loc_sort()
        merged = []
        for r in self:
            if merged and merged[-1].chrom == r.chrom and r.start < merged[-1].end:
                prev = merged[-1]
                r.start = prev.start
                r._sync()
                merged[-1] = r
            else:
                merged.append(r)
        self[:] = merged
        return self

    def subtract(self, b_regions):
        """Remove everything covered by `b_regions` (in place; pieces are new OneRegion objects)."""
        self.loc_sort()
        b_regions.loc_sort()
        b = list(b_regions)
        out = []
        bi = 0
        nb = len(b)
        # the sweep below consumes `self` as a stack so that split remainders are revisited against later b's
        pending = list(reversed(self))
        while pending:
            a = pending.pop()
            while bi < nb and (b[bi].chrom < a.chrom or (b[bi].chrom == a.chrom and b[bi].end <= a.start)):
                bi += 1
            if bi >= nb or b[bi].chrom != a.chrom or b[bi].start >= a.end:
                out.append(a)
                continue
            bs, be = b[bi].start, b[bi].end
            if a.start >= bs and a.end <= be:
                continue                                             # swallowed
            if a.start < bs and a.end > be:
                out.append(OneRegion([a.chrom, a.start, bs]))
                pending.append(OneRegion([a.chrom, be, a.end]))
            elif a.end > be:
                pending.append(OneRegion([a.chrom, be, a.end]))
            else:
                out.append(OneRegion([a.chrom, a.start, bs]))
        self[:] = out
        return self


# ---------------------------------------------------------------------- array views for the C-ABI
def to_arrays(regions, chrom_index):
    """(chrom_id int32[n], start int64[n], end int64[n]) for a list of regions; chrom_index: name -> id."""
    n = len(regions)
    cid = np.fromiter((chrom_index[r.chrom] for r in regions), dtype=np.int32, count=n)
    st = np.fromiter((r.start for r in regions), dtype=np.int64, count=n)
    en = np.fromiter((r.end for r in regions), dtype=np.int64, count=n)
    return cid, st, en
