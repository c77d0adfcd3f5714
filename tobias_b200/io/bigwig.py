"""
Minimal bigWig codec (pyBigWig / libBigWig are absent from the image).

Writer: what `bigwig_writer` needs (tobias/utils/utilities.py:129-232): a header of (chrom, length) pairs in file order
and `addEntries(chrom, positions, values=..., span=1)` calls in sorted order; values are stored as float32, only the
positions handed in are written (the reference passes the non-zero positions, :193-200).  Entries become
variableStep sections (type 2, span 1) of <= 1024 items, zlib-compressed, followed by the chromosome B+ tree and a
packed R-tree over the sections -- the layout any bigWig reader (UCSC tools, pyBigWig, IGV) understands.  No zoom levels.

Reader: what `OneRegion.get_signal` / `calculate_scores` need (tobias/utils/regions.py:163-188,
tobias/tools/score_bigwig.py:37-39): `chroms()` and `values(chrom, start, end)` -> float32 array with NaN where the
file holds no data.  Decodes bedGraph (1), variableStep (2) and fixedStep (3) sections, compressed or not.
"""
import struct
import zlib

import numpy as np

BIGWIG_MAGIC = 0x888FFC26
CHROM_TREE_MAGIC = 0x78CA8C91
RTREE_MAGIC = 0x2468ACE0
ITEMS_PER_SECTION = 1024
RTREE_BLOCK = 256
_VAR_ITEM = np.dtype([("start", "<u4"), ("value", "<f4")])
_BED_ITEM = np.dtype([("start", "<u4"), ("end", "<u4"), ("value", "<f4")])


class BigWigWriter:
    def __init__(self, path):
        self.path = path
        self._fh = open(path, "wb")
        self._chroms = None
        self._sections = []          # (chrom_id, start, end, file_offset, size)
        self._last = (-1, -1)
        self._n_valid = 0
        self._min = np.inf
        self._max = -np.inf
        self._sum = 0.0
        self._sumsq = 0.0
        self._max_uncompressed = 0
        self._closed = False

    def add_header(self, header):
        self._chroms = [(str(c), int(n)) for c, n in header]
        self._id = {c: i for i, (c, _) in enumerate(self._chroms)}
        key = max(len(c.encode()) for c, _ in self._chroms)
        self._key_size = key
        # fixed-size prologue: header (64) + total summary (40) + chromosome tree (32 + 4 + n * (key + 8)) + section count (8)
        self._summary_offset = 64
        self._chrom_tree_offset = 64 + 40
        tree_size = 32 + 4 + len(self._chroms) * (key + 8)
        self._data_offset = self._chrom_tree_offset + tree_size
        self._fh.write(b"\0" * (self._data_offset + 8))

    def add_entries(self, chrom, positions, values, span=1):
        if self._chroms is None:
            raise RuntimeError("add_header must be called first")
        if span != 1:
            raise NotImplementedError("span must be 1")
        cid = self._id[chrom]
        pos = np.ascontiguousarray(positions, dtype=np.int64)
        val = np.ascontiguousarray(values, dtype=np.float32)
        if pos.shape != val.shape:
            raise ValueError("positions and values differ in length")
        if pos.size == 0:
            return
        if pos[0] < 0 or pos[-1] >= self._chroms[cid][1]:
            raise RuntimeError("position outside of chromosome bounds")
        if (cid, int(pos[0])) <= self._last or (pos.size > 1 and np.any(pos[1:] <= pos[:-1])):
            raise RuntimeError("entries must be added in sorted order (chromosome order of the header, ascending positions)")
        self._last = (cid, int(pos[-1]))
        v64 = val.astype(np.float64)
        self._n_valid += pos.size
        self._min = min(self._min, float(v64.min()))
        self._max = max(self._max, float(v64.max()))
        self._sum += float(v64.sum())
        self._sumsq += float((v64 * v64).sum())
        items = np.empty(pos.size, dtype=_VAR_ITEM)
        items["start"] = pos
        items["value"] = val
        fh = self._fh
        for a in range(0, pos.size, ITEMS_PER_SECTION):
            b = min(a + ITEMS_PER_SECTION, pos.size)
            start, end = int(pos[a]), int(pos[b - 1]) + 1
            raw = struct.pack("<IIIIIBBH", cid, start, end, 0, 1, 2, 0, b - a) + items[a:b].tobytes()
            self._max_uncompressed = max(self._max_uncompressed, len(raw))
            comp = zlib.compress(raw, 6)
            off = fh.tell()
            fh.write(comp)
            self._sections.append((cid, start, end, off, len(comp)))

    # ------------------------------------------------------------------ index
    def _write_rtree(self):
        fh = self._fh
        index_offset = fh.tell()
        secs = self._sections
        n = len(secs)
        if n == 0:
            fh.write(struct.pack("<IIQIIIIQII", RTREE_MAGIC, RTREE_BLOCK, 0, 0, 0, 0, 0, index_offset, 1, 0))
            fh.write(struct.pack("<BBH", 1, 0, 0))
            return index_offset
        # levels[0] = leaves' items; each upper level summarises RTREE_BLOCK children
        level_items = [[(s[0], s[1], s[0], s[2], s[3], s[4]) for s in secs]]
        while len(level_items[-1]) > RTREE_BLOCK:
            prev = level_items[-1]
            cur = []
            for a in range(0, len(prev), RTREE_BLOCK):
                grp = prev[a:a + RTREE_BLOCK]
                end_c, end_b = max((g[2], g[3]) for g in grp)
                cur.append((grp[0][0], grp[0][1], end_c, end_b, None, None))
            level_items.append(cur)
        level_items.reverse()                      # root level first; last = section items
        header_size = 48
        # node sizes per level: node = 4 bytes + count * item_size (24 non-leaf, 32 leaf)
        offsets = []                               # offsets[level][node] file offset
        pos = index_offset + header_size
        n_levels = len(level_items)
        for li, items in enumerate(level_items):
            item_size = 32 if li == n_levels - 1 else 24
            offs = []
            for a in range(0, len(items), RTREE_BLOCK):
                offs.append(pos)
                pos += 4 + min(RTREE_BLOCK, len(items) - a) * item_size
            offsets.append(offs)
        first, last = secs[0], max(secs, key=lambda s: (s[0], s[2]))
        fh.write(struct.pack("<IIQIIIIQII", RTREE_MAGIC, RTREE_BLOCK, n, first[0], first[1], last[0], last[2], index_offset, 1, 0))
        for li, items in enumerate(level_items):
            leaf = li == n_levels - 1
            for node, a in enumerate(range(0, len(items), RTREE_BLOCK)):
                grp = items[a:a + RTREE_BLOCK]
                assert fh.tell() == offsets[li][node]
                fh.write(struct.pack("<BBH", 1 if leaf else 0, 0, len(grp)))
                for k, it in enumerate(grp):
                    if leaf:
                        fh.write(struct.pack("<IIIIQQ", it[0], it[1], it[2], it[3], it[4], it[5]))
                    else:
                        fh.write(struct.pack("<IIIIQ", it[0], it[1], it[2], it[3], offsets[li + 1][a + k]))
        return index_offset

    def close(self):
        if self._closed:
            return
        self._closed = True
        fh = self._fh
        if self._chroms is None:
            fh.close()
            return
        index_offset = self._write_rtree()
        fh.seek(0)
        fh.write(struct.pack("<IHHQQQHHQQIQ", BIGWIG_MAGIC, 4, 0, self._chrom_tree_offset, self._data_offset, index_offset,
                             0, 0, 0, self._summary_offset, self._max_uncompressed, 0))
        n = self._n_valid
        fh.write(struct.pack("<Qdddd", n, self._min if n else 0.0, self._max if n else 0.0, self._sum, self._sumsq))
        nchrom = len(self._chroms)
        fh.write(struct.pack("<IIIIQQ", CHROM_TREE_MAGIC, max(nchrom, 1), self._key_size, 8, nchrom, 0))
        fh.write(struct.pack("<BBH", 1, 0, nchrom))
        for name, _ in sorted(self._chroms):       # B+ tree keys are sorted
            cid = self._id[name]
            fh.write(name.encode().ljust(self._key_size, b"\0") + struct.pack("<II", cid, self._chroms[cid][1]))
        fh.write(struct.pack("<Q", len(self._sections)))
        fh.close()


class BigWigReader:
    def __init__(self, path):
        self.path = path
        self._fh = open(path, "rb")
        hdr = self._fh.read(64)
        if len(hdr) < 64:
            raise ValueError("%s: not a bigWig file (too short)" % path)
        (magic, self.version, self.zoom_levels, chrom_off, self.data_offset, self.index_offset, _fc, _dfc, _asql,
         self.summary_offset, self.uncompress_buf, _ext) = struct.unpack("<IHHQQQHHQQIQ", hdr)
        if magic != BIGWIG_MAGIC:
            raise ValueError("%s: not a little-endian bigWig file" % path)
        self.chroms = {}
        self._ids = {}
        self._read_chrom_tree(chrom_off)
        self._index_header = None

    def _read_chrom_tree(self, off):
        fh = self._fh
        fh.seek(off)
        magic, block, key, val, count, _ = struct.unpack("<IIIIQQ", fh.read(32))
        if magic != CHROM_TREE_MAGIC:
            raise ValueError("bad chromosome tree magic")

        def node(offset):
            fh.seek(offset)
            leaf, _, cnt = struct.unpack("<BBH", fh.read(4))
            if leaf:
                for _i in range(cnt):
                    k = fh.read(key).rstrip(b"\0").decode()
                    cid, size = struct.unpack("<II", fh.read(8))
                    self.chroms[k] = size
                    self._ids[k] = cid
            else:
                children = []
                for _i in range(cnt):
                    fh.read(key)
                    children.append(struct.unpack("<Q", fh.read(8))[0])
                for c in children:
                    node(c)
        node(off + 32)

    def header(self):
        fh = self._fh
        fh.seek(self.summary_offset)
        n, mn, mx, sm, sq = struct.unpack("<Qdddd", fh.read(40))
        return {"version": self.version, "nLevels": self.zoom_levels, "nBasesCovered": n, "minVal": mn, "maxVal": mx,
                "sumData": sm, "sumSquared": sq}

    def _blocks(self, cid, start, end):
        fh = self._fh
        fh.seek(self.index_offset)
        magic, block, count, *_rest = struct.unpack("<IIQIIIIQII", fh.read(48))
        if magic != RTREE_MAGIC:
            raise ValueError("bad R-tree magic")
        out = []

        def overlaps(sc, sb, ec, eb):
            return (sc, sb) < (cid, end) and (ec, eb) > (cid, start)

        def node(offset):
            fh.seek(offset)
            leaf, _, cnt = struct.unpack("<BBH", fh.read(4))
            if leaf:
                raw = fh.read(32 * cnt)
                for i in range(cnt):
                    sc, sb, ec, eb, off, size = struct.unpack_from("<IIIIQQ", raw, 32 * i)
                    if overlaps(sc, sb, ec, eb):
                        out.append((off, size))
            else:
                raw = fh.read(24 * cnt)
                kids = []
                for i in range(cnt):
                    sc, sb, ec, eb, off = struct.unpack_from("<IIIIQ", raw, 24 * i)
                    if overlaps(sc, sb, ec, eb):
                        kids.append(off)
                for k in kids:
                    node(k)
        if count:
            node(self.index_offset + 48)
        return out

    def values(self, chrom, start, end):
        if chrom not in self.chroms:
            raise RuntimeError("Invalid interval bounds!")
        start, end = int(start), int(end)
        if start < 0 or end > self.chroms[chrom] or start >= end:
            raise RuntimeError("Invalid interval bounds!")
        cid = self._ids[chrom]
        out = np.full(end - start, np.nan, dtype=np.float32)
        fh = self._fh
        for off, size in self._blocks(cid, start, end):
            fh.seek(off)
            raw = fh.read(size)
            if self.uncompress_buf:
                raw = zlib.decompress(raw)
            c, s0, e0, step, span, typ, _, n = struct.unpack_from("<IIIIIBBH", raw, 0)
            if c != cid:
                continue
            body = raw[24:]
            if typ == 2:
                it = np.frombuffer(body, dtype=_VAR_ITEM, count=n)
                st = it["start"].astype(np.int64)
                en = st + span
                vals = it["value"]
            elif typ == 1:
                it = np.frombuffer(body, dtype=_BED_ITEM, count=n)
                st, en, vals = it["start"].astype(np.int64), it["end"].astype(np.int64), it["value"]
            elif typ == 3:
                vals = np.frombuffer(body, dtype="<f4", count=n)
                st = s0 + step * np.arange(n, dtype=np.int64)
                en = st + span
            else:
                raise ValueError("unknown bigWig section type %d" % typ)
            if typ == 2 and span == 1 or typ == 3 and span == 1:
                keep = (st >= start) & (st < end)
                out[st[keep] - start] = vals[keep]
            else:
                for a, b, v in zip(st.tolist(), en.tolist(), vals.tolist()):
                    a, b = max(a, start), min(b, end)
                    if a < b:
                        out[a - start:b - start] = v
        return out

    def close(self):
        self._fh.close()
