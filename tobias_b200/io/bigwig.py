"""
Minimal bigWig codec (pyBigWig / libBigWig are absent from the image).

Writer: what `bigwig_writer` needs (tobias/utils/utilities.py:129-232): a header of (chrom, length) pairs in file order
and `addEntries(chrom, positions, values=..., span=1)` calls in sorted order; values are stored as float32, only the
positions handed in are written (the reference passes the non-zero positions, :193-200).  Entries become
variableStep sections (type 2, span 1) of <= 1024 items, zlib-compressed in parallel on the host cores by a native helper
(bwcodec.c), followed by the chromosome B+ tree, a packed R-tree over the sections and ZOOM LEVELS (summary records of
64, 256, 1024, ... bases with their own R-trees, as pyBigWig / bedGraphToBigWig write them) -- the layout any bigWig reader
(UCSC tools, pyBigWig, IGV) understands, including the genome browsers that refuse files without zoom levels.

Reader: what `OneRegion.get_signal` / `calculate_scores` need (tobias/utils/regions.py:163-188,
tobias/tools/score_bigwig.py:37-39): `chroms()` and `values(chrom, start, end)` -> float32 array with NaN where the
file holds no data.  Decodes bedGraph (1), variableStep (2) and fixedStep (3) sections, compressed or not; `zoom_records(level)`
returns the summaries of a zoom level.
"""
import ctypes
import os
import struct
import subprocess
import zlib

import numpy as np

BIGWIG_MAGIC = 0x888FFC26
CHROM_TREE_MAGIC = 0x78CA8C91
RTREE_MAGIC = 0x2468ACE0
ITEMS_PER_SECTION = 1024
ZOOM_ITEMS_PER_BLOCK = 512
RTREE_BLOCK = 256
MAX_ZOOM = 8
FIRST_REDUCTION = 64
_VAR_ITEM = np.dtype([("start", "<u4"), ("value", "<f4")])
_BED_ITEM = np.dtype([("start", "<u4"), ("end", "<u4"), ("value", "<f4")])
_ZOOM_REC = np.dtype([("chrom", "<u4"), ("start", "<u4"), ("end", "<u4"), ("valid", "<u4"), ("min", "<f4"), ("max", "<f4"),
                      ("sum", "<f4"), ("sumsq", "<f4")])
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


FIXED_STEP = True          # False: variableStep sections only (comparison runs, tests)


def _codec():
    global _LIB
    if _LIB is None:
        so, src = os.path.join(_HERE, "_bwcodec.so"), os.path.join(_HERE, "bwcodec.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", so, src, "-lz"])
        _LIB = ctypes.CDLL(so)
        for f in ("tb_bw_bound", "tb_bw_sections", "tb_bw_sections_fixed", "tb_bw_zoom", "tb_bw_zoom0", "tb_bw_pack"):
            getattr(_LIB, f).restype = ctypes.c_int64
    return _LIB


def _vp(a):
    return ctypes.c_void_p(a.ctypes.data)


def _reduce_sorted(keys, valid, mn, mx, sm, sq):
    """Summaries of runs of equal keys (keys ascending): first index of every run and the reduced columns."""
    first = np.flatnonzero(np.concatenate([[True], keys[1:] != keys[:-1]]))
    return (first, np.add.reduceat(valid, first), np.minimum.reduceat(mn, first), np.maximum.reduceat(mx, first),
            np.add.reduceat(sm, first), np.add.reduceat(sq, first))


class BigWigWriter:
    def __init__(self, path, zoom=True, level=3):       # float tracks barely compress: level 3 is 4x faster than 6 and no larger (measured)
        self.path = path
        self._fh = open(path, "wb")
        self._chroms = None
        self._sec = []               # per add_entries call: (chrom id, starts, ends, file offsets, sizes) of its sections
        self._zoom0 = []             # level-0 summary columns per call: (chrom id array, bin, valid, min, max, sum, sumsq)
        self._last = (-1, -1)
        self._n_valid = 0
        self._min = np.inf
        self._max = -np.inf
        self._sum = 0.0
        self._sumsq = 0.0
        self._max_uncompressed = 0
        self._closed = False
        self._zoom = bool(zoom)
        self._level = int(level)

    def add_header(self, header):
        self._chroms = [(str(c), int(n)) for c, n in header]
        self._id = {c: i for i, (c, _) in enumerate(self._chroms)}
        key = max(len(c.encode()) for c, _ in self._chroms)
        self._key_size = key
        # fixed-size prologue: header (64) + zoom headers (24 x MAX_ZOOM) + total summary (40) + chromosome tree (32 + 4 + n * (key + 8)) + section count (8)
        self._summary_offset = 64 + 24 * MAX_ZOOM
        self._chrom_tree_offset = self._summary_offset + 40
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
        # ---- totals and the first zoom level (one summary per FIRST_REDUCTION bases that hold data): one native pass over the items
        lib = _codec()
        n_max = int(min(pos.size, (int(pos[-1]) - int(pos[0])) // FIRST_REDUCTION + 2))
        zb, zn = np.empty(n_max, np.int64), np.empty(n_max, np.int64)
        zmn, zmx, zsm, zsq = (np.empty(n_max, np.float64) for _ in range(4))
        tot = np.zeros(4, np.float64)
        nz = int(lib.tb_bw_zoom0(_vp(pos), _vp(val), ctypes.c_int64(pos.size), ctypes.c_int32(FIRST_REDUCTION), _vp(zb), _vp(zn), _vp(zmn), _vp(zmx),
                                 _vp(zsm), _vp(zsq), _vp(tot)))
        self._n_valid += pos.size
        self._min = min(self._min, float(tot[0]))
        self._max = max(self._max, float(tot[1]))
        self._sum += float(tot[2])
        self._sumsq += float(tot[3])
        if self._zoom:
            self._zoom0.append((cid, zb[:nz].copy(), zn[:nz].copy(), zmn[:nz].copy(), zmx[:nz].copy(), zsm[:nz].copy(), zsq[:nz].copy()))
        # ---- data sections: built and compressed in parallel by the native helper
        from .bam import parallel_ranges
        # runs of consecutive positions (what ATACorrect / ScoreBigwig write: every base of a region) go into fixedStep sections -- 4 bytes per
        # item instead of the 8 of variableStep, half the bytes through zlib; scattered positions keep variableStep sections of a fixed count
        brk = np.flatnonzero(np.diff(pos) != 1) + 1
        if FIXED_STEP and (brk.size + 1) * 16 <= pos.size:
            run_a = np.concatenate([[0], brk])
            run_b = np.concatenate([brk, [pos.size]])
            n_chunks = (run_b - run_a + ITEMS_PER_SECTION - 1) // ITEMS_PER_SECTION
            first = np.repeat(run_a, n_chunks)
            within = np.arange(int(n_chunks.sum()), dtype=np.int64) - np.repeat(np.cumsum(n_chunks) - n_chunks, n_chunks)
            sec_a = np.ascontiguousarray(first + within * ITEMS_PER_SECTION, dtype=np.int64)
            sec_b = np.ascontiguousarray(np.minimum(sec_a + ITEMS_PER_SECTION, np.repeat(run_b, n_chunks)), dtype=np.int64)
            n_sec = int(sec_a.shape[0])
            item_bytes = 4
        else:
            sec_a = sec_b = None
            n_sec = (pos.size + ITEMS_PER_SECTION - 1) // ITEMS_PER_SECTION
            item_bytes = 8
        slot = int(lib.tb_bw_bound(ctypes.c_int64(24 + item_bytes * ITEMS_PER_SECTION)))
        out = np.empty(n_sec * slot, dtype=np.uint8)
        off, ln = np.empty(n_sec, np.int64), np.empty(n_sec, np.int64)
        st, en = np.empty(n_sec, np.int32), np.empty(n_sec, np.int32)
        if sec_a is not None:
            parallel_ranges(lambda a, b: lib.tb_bw_sections_fixed(ctypes.c_int32(cid), _vp(pos), _vp(val), _vp(sec_a), _vp(sec_b),
                                                                  ctypes.c_int32(ITEMS_PER_SECTION), ctypes.c_int32(self._level), ctypes.c_int64(a),
                                                                  ctypes.c_int64(b), _vp(out), ctypes.c_int64(slot), _vp(ln), _vp(st), _vp(en)), 0, n_sec, 64)
        else:
            parallel_ranges(lambda a, b: lib.tb_bw_sections(ctypes.c_int32(cid), _vp(pos), _vp(val), ctypes.c_int64(pos.size),
                                                            ctypes.c_int32(ITEMS_PER_SECTION), ctypes.c_int32(self._level), ctypes.c_int64(a),
                                                            ctypes.c_int64(b), _vp(out), ctypes.c_int64(slot), _vp(ln), _vp(st), _vp(en)), 0, n_sec, 64)
        total = lib.tb_bw_pack(_vp(out), ctypes.c_int64(n_sec), ctypes.c_int64(slot), _vp(ln), _vp(off))
        if total < 0:
            raise RuntimeError("bigWig section compression failed (%d)" % total)
        self._max_uncompressed = max(self._max_uncompressed, 24 + item_bytes * min(pos.size, ITEMS_PER_SECTION))
        base = self._fh.tell()
        self._fh.write(out[:total].tobytes())
        self._sec.append((cid, st.astype(np.int64), en.astype(np.int64), off + base, ln))

    # ------------------------------------------------------------------ index
    def _write_rtree(self, sc, sb, ec, eb, off, size):
        """Packed R-tree (cirTree) over blocks given as columns: start chrom / base, end chrom / base, file offset, size."""
        fh = self._fh
        index_offset = fh.tell()
        n = int(sc.shape[0])
        if n == 0:
            fh.write(struct.pack("<IIQIIIIQII", RTREE_MAGIC, RTREE_BLOCK, 0, 0, 0, 0, 0, index_offset, 1, 0))
            fh.write(struct.pack("<BBH", 1, 0, 0))
            return index_offset
        # levels[-1] = the blocks themselves; each upper level summarises RTREE_BLOCK children
        levels = [(sc, sb, ec, eb)]
        while levels[-1][0].shape[0] > RTREE_BLOCK:
            psc, psb, pec, peb = levels[-1]
            first = np.arange(0, psc.shape[0], RTREE_BLOCK)
            key = pec.astype(np.int64) << 32 | peb.astype(np.int64)
            kmax = np.maximum.reduceat(key, first)
            levels.append((psc[first], psb[first], (kmax >> 32).astype(np.int64), (kmax & 0xFFFFFFFF).astype(np.int64)))
        levels.reverse()                           # root level first
        n_levels = len(levels)
        pos = index_offset + 48
        node_off = []                              # per level: file offset of every node
        for li, lv in enumerate(levels):
            item = 32 if li == n_levels - 1 else 24
            cnt = lv[0].shape[0]
            per_node = np.minimum(RTREE_BLOCK, cnt - np.arange(0, cnt, RTREE_BLOCK))
            sizes = 4 + per_node * item
            offs = pos + np.concatenate([[0], np.cumsum(sizes)[:-1]])
            node_off.append(offs)
            pos += int(sizes.sum())
        kall = ec.astype(np.int64) << 32 | eb.astype(np.int64)
        kl = int(kall.max())
        fh.write(struct.pack("<IIQIIIIQII", RTREE_MAGIC, RTREE_BLOCK, n, int(sc[0]), int(sb[0]), kl >> 32, kl & 0xFFFFFFFF, index_offset, 1, 0))
        leaf_dt = np.dtype([("sc", "<u4"), ("sb", "<u4"), ("ec", "<u4"), ("eb", "<u4"), ("off", "<u8"), ("size", "<u8")])
        node_dt = np.dtype([("sc", "<u4"), ("sb", "<u4"), ("ec", "<u4"), ("eb", "<u4"), ("off", "<u8")])
        for li, (lsc, lsb, lec, leb) in enumerate(levels):
            leaf = li == n_levels - 1
            cnt = lsc.shape[0]
            items = np.empty(cnt, dtype=leaf_dt if leaf else node_dt)
            items["sc"], items["sb"], items["ec"], items["eb"] = lsc, lsb, lec, leb
            if leaf:
                items["off"], items["size"] = off, size
            else:
                items["off"] = node_off[li + 1]
            raw = items.tobytes()
            isz = items.dtype.itemsize
            for node, a in enumerate(range(0, cnt, RTREE_BLOCK)):
                k = min(RTREE_BLOCK, cnt - a)
                assert fh.tell() == node_off[li][node]
                fh.write(struct.pack("<BBH", 1 if leaf else 0, 0, k) + raw[a * isz:(a + k) * isz])
        return index_offset

    def _write_zoom_levels(self):
        """(reduction, data offset, index offset) of every zoom level written."""
        if not self._zoom or not self._zoom0:
            return []
        lib = _codec()
        fh = self._fh
        cid = np.concatenate([np.full(z[1].shape[0], z[0], dtype=np.int64) for z in self._zoom0])
        bins, valid, mn, mx, sm, sq = (np.concatenate([z[k] for z in self._zoom0]) for k in range(1, 7))
        lens = np.array([n for _, n in self._chroms], dtype=np.int64)
        out_levels = []
        red = FIRST_REDUCTION
        while len(out_levels) < MAX_ZOOM:
            n = cid.shape[0]
            rec = np.empty(n, dtype=_ZOOM_REC)
            rec["chrom"], rec["start"] = cid, bins * red
            rec["end"] = np.minimum((bins + 1) * red, lens[cid])
            rec["valid"], rec["min"], rec["max"], rec["sum"], rec["sumsq"] = valid, mn, mx, sm, sq
            from .bam import parallel_ranges
            n_blk = (n + ZOOM_ITEMS_PER_BLOCK - 1) // ZOOM_ITEMS_PER_BLOCK
            slot = int(lib.tb_bw_bound(ctypes.c_int64(32 * ZOOM_ITEMS_PER_BLOCK)))
            buf = np.empty(n_blk * slot, dtype=np.uint8)
            boff, blen = np.empty(n_blk, np.int64), np.empty(n_blk, np.int64)
            raw = np.frombuffer(rec.tobytes(), dtype=np.uint8)
            parallel_ranges(lambda a, b: lib.tb_bw_zoom(_vp(raw), ctypes.c_int64(n), ctypes.c_int32(ZOOM_ITEMS_PER_BLOCK), ctypes.c_int32(self._level),
                                                        ctypes.c_int64(a), ctypes.c_int64(b), _vp(buf), ctypes.c_int64(slot), _vp(blen)), 0, n_blk, 64)
            total = lib.tb_bw_pack(_vp(buf), ctypes.c_int64(n_blk), ctypes.c_int64(slot), _vp(blen), _vp(boff))
            if total < 0:
                raise RuntimeError("bigWig zoom compression failed (%d)" % total)
            self._max_uncompressed = max(self._max_uncompressed, 32 * min(n, ZOOM_ITEMS_PER_BLOCK))
            data_offset = fh.tell()
            fh.write(struct.pack("<I", n))
            base = fh.tell()
            fh.write(buf[:total].tobytes())
            first = np.arange(0, n, ZOOM_ITEMS_PER_BLOCK)
            last = np.minimum(first + ZOOM_ITEMS_PER_BLOCK, n) - 1
            index_offset = self._write_rtree(cid[first], rec["start"][first].astype(np.int64), cid[last], rec["end"][last].astype(np.int64),
                                             boff + base, blen)
            out_levels.append((red, data_offset, index_offset))
            if n <= ZOOM_ITEMS_PER_BLOCK:
                break
            # next level: four times coarser, reduced from this one
            key = cid << 40 | (bins // 4)
            f2, valid, mn, mx, sm, sq = _reduce_sorted(key, valid, mn, mx, sm, sq)
            cid, bins = cid[f2], (bins // 4)[f2]
            red *= 4
        return out_levels

    def close(self):
        if self._closed:
            return
        self._closed = True
        fh = self._fh
        if self._chroms is None:
            fh.close()
            return
        if self._sec:
            sc = np.concatenate([np.full(s[1].shape[0], s[0], dtype=np.int64) for s in self._sec])
            sb, eb, off, size = (np.concatenate([s[k] for s in self._sec]) for k in range(1, 5))
        else:
            sc = sb = eb = off = size = np.zeros(0, dtype=np.int64)
        n_sections = int(sc.shape[0])
        index_offset = self._write_rtree(sc, sb, sc, eb, off, size)
        zooms = self._write_zoom_levels()
        fh.seek(0)
        fh.write(struct.pack("<IHHQQQHHQQIQ", BIGWIG_MAGIC, 4, len(zooms), self._chrom_tree_offset, self._data_offset, index_offset,
                             0, 0, 0, self._summary_offset, self._max_uncompressed, 0))
        for red, doff, ioff in zooms:
            fh.write(struct.pack("<IIQQ", red, 0, doff, ioff))
        fh.seek(self._summary_offset)
        n = self._n_valid
        fh.write(struct.pack("<Qdddd", n, self._min if n else 0.0, self._max if n else 0.0, self._sum, self._sumsq))
        nchrom = len(self._chroms)
        fh.write(struct.pack("<IIIIQQ", CHROM_TREE_MAGIC, max(nchrom, 1), self._key_size, 8, nchrom, 0))
        fh.write(struct.pack("<BBH", 1, 0, nchrom))
        for name, _ in sorted(self._chroms):       # B+ tree keys are sorted
            cid = self._id[name]
            fh.write(name.encode().ljust(self._key_size, b"\0") + struct.pack("<II", cid, self._chroms[cid][1]))
        fh.write(struct.pack("<Q", n_sections))
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

    def values_many(self, chrom, starts, ends):
        """values() of many intervals of one chromosome at once (starts ascending): the sections the intervals touch are read in one
        sweep, inflated on the thread pool and scattered into the concatenated output with array operations.  Returns (flat float32
        array, offsets: interval i occupies flat[off[i]:off[i + 1]])."""
        starts, ends = np.asarray(starts, dtype=np.int64), np.asarray(ends, dtype=np.int64)
        if chrom not in self.chroms or starts.size == 0:
            raise RuntimeError("Invalid interval bounds!")
        if starts.min() < 0 or ends.max() > self.chroms[chrom] or np.any(starts >= ends):
            raise RuntimeError("Invalid interval bounds!")
        off = np.concatenate([[0], np.cumsum(ends - starts)])
        out = np.full(int(off[-1]), np.nan, dtype=np.float32)
        cid = self._ids[chrom]
        blocks = sorted(set(self._blocks(cid, int(starts.min()), int(ends.max()))))
        if not blocks or np.any(starts[1:] < starts[:-1]):
            for i, (a, b) in enumerate(zip(starts.tolist(), ends.tolist())):           # unsorted intervals: one by one
                out[off[i]:off[i + 1]] = self.values(chrom, a, b)
            return out, off
        # the intervals usually cover a small part of the span: keep the sections that overlap one (section spans come with the R-tree leaves
        # only after inflation, so the test happens on the decoded header)
        from .bam import parallel_ranges
        fh = self._fh
        raws = []
        for o, size in blocks:
            fh.seek(o)
            raws.append(fh.read(size))
        if self.uncompress_buf:
            dec = [None] * len(raws)

            def inflate(a, b):
                for k in range(a, b):
                    dec[k] = zlib.decompress(raws[k])
            parallel_ranges(inflate, 0, len(raws), 64)
        else:
            dec = raws
        pos_l, val_l = [], []
        for raw in dec:
            c, s0, e0, step, span, typ, _, n = struct.unpack_from("<IIIIIBBH", raw, 0)
            if c != cid:
                continue
            if typ == 3 and span == 1 and step == 1:
                pos_l.append(np.arange(s0, s0 + n, dtype=np.int64))
                val_l.append(np.frombuffer(raw, dtype="<f4", count=n, offset=24))
            elif typ == 2 and span == 1:
                it = np.frombuffer(raw, dtype=_VAR_ITEM, count=n, offset=24)
                pos_l.append(it["start"].astype(np.int64))
                val_l.append(it["value"])
            else:                                                                      # bedGraph items / wider spans: the general reader
                for i, (a, b) in enumerate(zip(starts.tolist(), ends.tolist())):
                    out[off[i]:off[i + 1]] = self.values(chrom, a, b)
                return out, off
        if pos_l:
            pos, val = np.concatenate(pos_l), np.concatenate(val_l)                  # ascending: the sections of a chromosome are in file order
            if np.any(starts[1:] < ends[:-1]):
                # overlapping intervals: a position can belong to several of them -- one slice of the data per interval
                lo, hi = np.searchsorted(pos, starts), np.searchsorted(pos, ends)
                for i in range(starts.size):
                    out[off[i] + (pos[lo[i]:hi[i]] - starts[i])] = val[lo[i]:hi[i]]
            else:
                r = np.searchsorted(starts, pos, side="right") - 1                    # the interval starting at or before the position
                ok = r >= 0
                ok[ok] = pos[ok] < ends[r[ok]]
                r, pos, val = r[ok], pos[ok], val[ok]
                out[off[r] + (pos - starts[r])] = val
        return out, off

    def zoom_headers(self):
        """[(reduction level, data offset, index offset)] of the zoom levels."""
        self._fh.seek(64)
        raw = self._fh.read(24 * self.zoom_levels)
        return [struct.unpack_from("<IIQQ", raw, 24 * i)[::1] for i in range(self.zoom_levels)]

    def zoom_records(self, level):
        """All summary records of one zoom level (numpy structured array: chrom, start, end, valid, min, max, sum, sumsq), via its R-tree."""
        red, _res, _doff, ioff = self.zoom_headers()[level]
        save = self.index_offset
        self.index_offset = ioff
        try:
            blocks = []
            for cid in sorted(self._ids.values()):
                blocks += self._blocks(cid, 0, 0xFFFFFFFF)
        finally:
            self.index_offset = save
        out = []
        for off, size in sorted(set(blocks)):
            self._fh.seek(off)
            raw = self._fh.read(size)
            if self.uncompress_buf:
                raw = zlib.decompress(raw)
            out.append(np.frombuffer(raw, dtype=_ZOOM_REC))
        return red, (np.concatenate(out) if out else np.zeros(0, dtype=_ZOOM_REC))

    def close(self):
        self._fh.close()
