"""
Minimal BAM codec (pysam / htslib / samtools are absent from the image).

BamReader / read_bam(path, contigs=None) -> ReadColumns.  The file is memory-mapped; a native helper (bamdecode.c: zlib + OpenMP,
compiled with gcc on first use) scans the BGZF block table, inflates the blocks of a batch in parallel on the host cores and walks the
alignment records of the batch into the columns the hot path consumes.  Memory stays bounded (one batch of inflated blocks at a
time, a partial record carried into the next batch), records of contigs the caller does not keep are skipped during the walk, and --
with an index -- only the blocks of the wanted contigs are inflated at all: a `.bai` next to the BAM (its linear index gives the
virtual offset of every contig's first record) or the `.tbidx` sidecar a first full pass writes (contig -> uncompressed offset of its
first record).  One process per GPU therefore decodes its own contigs only.  Region queries are binary searches on the sorted columns
(ReadColumns.fetch_index).

write_bam(path, ReadColumns): coordinate-sorted BAM of synthetic records (used by the tests and by the synthetic data
generator); sequence and qualities are placeholders ('N', 0xFF).  No .bai is written.
"""
import ctypes
import mmap
import os
import struct
import subprocess
import zlib

import numpy as np

from ..reads import ReadColumns, NO_CIGAR, FLAG_UNMAPPED

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_EOF_BLOCK = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
_COLS = (("chrom_id", np.int32), ("ref_start", np.int32), ("ref_end", np.int32), ("left_soft", np.int32), ("right_soft", np.int32),
         ("query_length", np.int32), ("flag", np.uint16), ("first_op", np.uint8), ("last_op", np.uint8), ("tlen", np.int32))


def build_helper(force=False):
    so = os.path.join(_HERE, "_bamdecode.so")
    src = os.path.join(_HERE, "bamdecode.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", so, src, "-lz"])
    return so


def _helper():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_helper())
        _LIB.tb_bam_decode.restype = ctypes.c_int64
        _LIB.tb_bgzf_scan.restype = ctypes.c_int64
        _LIB.tb_bgzf_inflate.restype = ctypes.c_int64
    return _LIB


def _vp(a):
    return ctypes.c_void_p(a.ctypes.data)


def parallel_ranges(fn, lo, hi, min_chunk, threads=None):
    """fn(a, b) over [lo, hi) split into contiguous parts on a thread pool (the native helpers release the GIL).  Results in order."""
    from concurrent.futures import ThreadPoolExecutor
    threads = threads or min(32, os.cpu_count() or 1)
    n = hi - lo
    parts = max(1, min(threads, n // max(1, min_chunk)))
    if parts == 1:
        return [fn(lo, hi)]
    step = -(-n // parts)
    cuts = [(lo + i * step, min(hi, lo + (i + 1) * step)) for i in range(parts) if lo + i * step < hi]
    with ThreadPoolExecutor(len(cuts)) as ex:
        return list(ex.map(lambda ab: fn(*ab), cuts))


# ------------------------------------------------------------------------------------------------ reader
class BamReader:
    def __init__(self, path):
        self.path = path
        self._fh = open(path, "rb")
        size = os.fstat(self._fh.fileno()).st_size
        if size < 28:
            raise ValueError("%s: not a BGZF file (too short)" % path)
        self._mm = mmap.mmap(self._fh.fileno(), 0, access=mmap.ACCESS_READ)
        self.buf = np.frombuffer(self._mm, dtype=np.uint8)
        cap = size // 28 + 2                                           # a BGZF block is at least 28 bytes
        self.block_off, self.payload_off = np.empty(cap, np.int64), np.empty(cap, np.int64)
        self.payload_size, self.usize = np.empty(cap, np.int32), np.empty(cap, np.int32)
        nb = _helper().tb_bgzf_scan(_vp(self.buf), ctypes.c_int64(size), ctypes.c_int64(cap), _vp(self.block_off), _vp(self.payload_off),
                                    _vp(self.payload_size), _vp(self.usize))
        if nb < 0:
            raise ValueError("%s: not a BGZF file (bad block header)" % path)
        self.n_blocks = int(nb)
        for name in ("block_off", "payload_off", "payload_size", "usize"):
            setattr(self, name, getattr(self, name)[:nb])
        self.ustart = np.concatenate([[0], np.cumsum(self.usize, dtype=np.int64)])      # uncompressed offset of every block
        self.references, self.lengths, self.header_end = self._read_header()

    def close(self):
        self.buf = None
        try:
            self._mm.close()
        except BufferError:
            pass
        self._fh.close()

    # ---- pieces
    def _inflate(self, b0, b1):
        """uint8 array of the inflated blocks [b0, b1)."""
        out = np.empty(int(self.ustart[b1] - self.ustart[b0]), dtype=np.uint8)
        rel = np.ascontiguousarray(self.ustart[b0:b1 + 1] - self.ustart[b0])
        lib = _helper()

        def part(a, b):          # blocks [a, b): ctypes releases the GIL, so the parts inflate on different cores
            return lib.tb_bgzf_inflate(_vp(self.buf), _vp(self.payload_off), _vp(self.payload_size), _vp(self.usize), ctypes.c_int64(a),
                                       ctypes.c_int64(b - a), ctypes.c_void_p(rel.ctypes.data + 8 * (a - b0)), _vp(out))
        for rc in parallel_ranges(part, b0, b1, 256):
            if rc != 0:
                raise ValueError("%s: BGZF block %d does not inflate (corrupt file)" % (self.path, -rc - 1))
        return out

    def _read_header(self):
        b1, data = 0, b""
        while b1 < self.n_blocks:
            b1 = min(self.n_blocks, max(b1 * 2, 4))
            data = self._inflate(0, b1).tobytes()
            parsed = _parse_header(data)
            if parsed is not None:
                return parsed
        raise ValueError("%s: truncated BAM header" % self.path)

    def _index(self):
        """ref id -> uncompressed offset of its first record (-1: no records), from <bam>.tbidx, a .bai, or None."""
        n_ref = len(self.references)
        st = os.stat(self.path)
        side = self.path + ".tbidx"
        if os.path.exists(side):
            try:
                raw = np.fromfile(side, dtype=np.int64)
                if raw.shape[0] == 3 + n_ref and raw[0] == 0x58444954 and raw[1] == st.st_size and raw[2] == int(st.st_mtime):
                    return raw[3:].copy()
            except OSError:
                pass
        for bai in (self.path + ".bai", os.path.splitext(self.path)[0] + ".bai"):
            if os.path.exists(bai) and os.path.getmtime(bai) >= st.st_mtime - 1:
                try:
                    return self._index_from_bai(bai)
                except (ValueError, struct.error, IndexError):
                    pass
        return None

    def _index_from_bai(self, bai):
        raw = open(bai, "rb").read()
        if raw[:4] != b"BAI\1":
            raise ValueError("not a BAI file")
        n_ref = struct.unpack_from("<i", raw, 4)[0]
        p = 8
        first = np.full(len(self.references), -1, dtype=np.int64)
        for r in range(n_ref):
            n_bin = struct.unpack_from("<i", raw, p)[0]
            p += 4
            vmin = None
            for _ in range(n_bin):
                b, n_chunk = struct.unpack_from("<Ii", raw, p)
                p += 8
                if b != 37450:                                         # the metadata pseudo-bin holds no alignments
                    for k in range(n_chunk):
                        beg = struct.unpack_from("<Q", raw, p + 16 * k)[0]
                        vmin = beg if vmin is None else min(vmin, beg)
                p += 16 * n_chunk
            n_intv = struct.unpack_from("<i", raw, p)[0]
            p += 4 + 8 * n_intv
            if vmin is not None and r < first.shape[0]:
                blk = int(np.searchsorted(self.block_off, vmin >> 16, side="left"))
                if blk >= self.n_blocks or self.block_off[blk] != (vmin >> 16):
                    raise ValueError("BAI offset does not hit a block start")
                first[r] = self.ustart[blk] + (vmin & 0xFFFF)
        return first

    def write_index(self, first):
        st = os.stat(self.path)
        try:
            np.concatenate([np.array([0x58444954, st.st_size, int(st.st_mtime)], dtype=np.int64), first.astype(np.int64)]).tofile(self.path + ".tbidx")
        except OSError:
            pass                                                       # read-only directory: the next run scans again

    # ---- the read
    def has_index(self):
        return self._index() is not None

    def build_index(self, batch_bytes=256 << 20):
        """One pass over the file that keeps no record, only where every contig starts; writes <bam>.tbidx.  (What rank 0 does before
        the ranks of a multi-GPU run read their own contigs.)"""
        if self._index() is None:
            self.read(contigs=None, batch_bytes=batch_bytes, _index_only=True)
        return self._index()

    def read(self, contigs=None, batch_bytes=256 << 20, write_index=True, _index_only=False):
        """ReadColumns of the whole file, or of the named contigs only.  The contig ids of the result refer to the file's reference list."""
        n_ref = len(self.references)
        keep = None
        u0, want_max = self.header_end, n_ref - 1
        index = self._index()
        if _index_only:
            keep = np.zeros(max(n_ref, 1), dtype=np.uint8)
        elif contigs is not None:
            ids = sorted(self.references.index(c) for c in contigs if c in self.references)
            keep = np.zeros(max(n_ref, 1), dtype=np.uint8)
            keep[ids] = 1
            want_max = ids[-1] if ids else -1
            if index is not None:
                starts = [int(index[i]) for i in ids if index[i] >= 0]
                u0 = min(starts) if starts else None
        if u0 is None or want_max < 0:
            return ReadColumns(self.references, self.lengths, **{k: np.empty(0, dt) for k, dt in _COLS})
        first_off = np.full(max(n_ref, 1), -1, dtype=np.int64) if (index is None and (contigs is None or _index_only)) else None
        b = int(np.searchsorted(self.ustart, u0, side="right") - 1)
        carry = np.empty(0, dtype=np.uint8)
        skip = u0 - int(self.ustart[b])                                # bytes of block b before the first wanted record
        base = int(self.ustart[b]) + skip
        parts = {k: [] for k, _ in _COLS}
        consumed, last_ref = ctypes.c_int64(0), ctypes.c_int32(-2)
        lib = _helper()
        done = False
        while b < self.n_blocks and not done:
            b1 = b + 1
            while b1 < self.n_blocks and self.ustart[b1 + 1] - self.ustart[b] <= batch_bytes:
                b1 += 1
            fresh = self._inflate(b, b1)
            data = fresh[skip:] if carry.size == 0 else np.concatenate([carry, fresh[skip:]])
            skip = 0
            cap = data.shape[0] // 36 + 1                              # a record takes at least 36 bytes
            cols = {k: np.empty(cap, dt) for k, dt in _COLS}
            n = lib.tb_bam_decode(_vp(data), ctypes.c_int64(data.shape[0]), ctypes.c_int64(cap), ctypes.byref(consumed),
                                  _vp(keep) if keep is not None else None, ctypes.c_int32(n_ref), ctypes.byref(last_ref),
                                  _vp(first_off) if first_off is not None else None, ctypes.c_int64(base),
                                  *[_vp(cols[k]) for k, _ in _COLS])
            if n < 0:
                raise ValueError("%s: malformed alignment record at uncompressed offset %d" % (self.path, base + consumed.value))
            for k, _ in _COLS:
                parts[k].append(cols[k][:n].copy())
            carry = data[consumed.value:].copy()
            base += consumed.value
            b = b1
            if last_ref.value > want_max or (last_ref.value == -1 and consumed.value > 0):
                done = True                                            # coordinate order: nothing wanted lies behind (unplaced reads come last)
        if not done and carry.size:
            raise ValueError("%s: trailing bytes after the last complete BAM record (file truncated?)" % self.path)
        if first_off is not None and write_index:
            self.write_index(first_off[:n_ref])
        rc = ReadColumns(self.references, self.lengths, **{k: np.concatenate(parts[k]) if parts[k] else np.empty(0, dt) for k, dt in _COLS})
        if not rc.is_sorted():
            raise ValueError("%s is not coordinate sorted" % self.path)
        return rc


def read_bam_header(path):
    """(references, lengths) -- what the driver reads first (tobias/tools/atacorrect.py:131-140)."""
    r = BamReader(path)
    try:
        return r.references, r.lengths
    finally:
        r.close()


def read_bam(path, contigs=None, header_only=False, batch_bytes=256 << 20):
    r = BamReader(path)
    try:
        if header_only:
            return r.references, r.lengths
        return r.read(contigs, batch_bytes)
    finally:
        r.close()


def bgzf_decompress(buf):
    """The whole inflated stream of a BGZF buffer (tests; the reader itself works in batches)."""
    out = []
    off, n = 0, len(buf)
    while off + 18 <= n:
        xlen = struct.unpack_from("<H", buf, off + 10)[0]
        bsize = struct.unpack_from("<H", buf, off + 16)[0] + 1
        out.append(zlib.decompress(bytes(buf[off + 12 + xlen:off + bsize - 8]), -15))
        off += bsize
    return b"".join(out)


def _parse_header(data):
    if len(data) < 12:
        return None
    if data[:4] != b"BAM\1":
        raise ValueError("not a BAM file")
    l_text = struct.unpack_from("<i", data, 4)[0]
    p = 8 + l_text
    if len(data) < p + 4:
        return None
    n_ref = struct.unpack_from("<i", data, p)[0]
    p += 4
    names, lens = [], []
    for _ in range(n_ref):
        if len(data) < p + 4:
            return None
        l_name = struct.unpack_from("<i", data, p)[0]
        if len(data) < p + 4 + l_name + 4:
            return None
        names.append(data[p + 4:p + 4 + l_name - 1].decode())
        lens.append(struct.unpack_from("<i", data, p + 4 + l_name)[0])
        p += 8 + l_name
    return names, lens, p


def bgzf_compress(data, level=6, block_offsets=None):
    """block_offsets (a list): receives the file offset of every block (virtual offsets of a .bai)."""
    out, pos = [], 0
    for a in range(0, len(data), 0xFF00):
        chunk = data[a:a + 0xFF00]
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        comp = c.compress(chunk) + c.flush()
        bsize = len(comp) + 25
        out.append(struct.pack("<BBBBIBBHBBHH", 0x1F, 0x8B, 8, 4, 0, 0, 0xFF, 6, 66, 67, 2, bsize) + comp +
                   struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
        if block_offsets is not None:
            block_offsets.append(pos)
        pos += len(out[-1])
    out.append(_EOF_BLOCK)
    return b"".join(out)


# ------------------------------------------------------------------------------------------------ writer
def _reg2bin(beg, end):
    end -= 1
    if beg >> 14 == end >> 14:
        return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17:
        return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20:
        return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23:
        return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26:
        return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def write_bam(path, rc, level=4, bai=False):
    """bai=True also writes <path>.bai (bins of one chunk each + the 16-kb linear index), enough for every reader that seeks by contig."""
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % (n, ln) for n, ln in zip(rc.references, rc.lengths))
    parts = [b"BAM\1", struct.pack("<i", len(text)), text.encode(), struct.pack("<i", len(rc.references))]
    for n, ln in zip(rc.references, rc.lengths):
        nm = n.encode() + b"\0"
        parts.append(struct.pack("<i", len(nm)) + nm + struct.pack("<i", ln))
    recs = []
    rec_meta = []                                                  # (ref id, start, end, length of the encoded record)
    for i in range(len(rc)):
        cid, pos, end = int(rc.chrom_id[i]), int(rc.ref_start[i]), int(rc.ref_end[i])
        ls, rs, ql, fl = int(rc.left_soft[i]), int(rc.right_soft[i]), int(rc.query_length[i]), int(rc.flag[i])
        fo, lo = int(rc.first_op[i]), int(rc.last_op[i])
        cigar = []
        if fo != NO_CIGAR and not (fl & FLAG_UNMAPPED):
            if fo == 4:
                cigar.append((ls << 4) | 4)
            elif fo != 0:
                cigar.append((1 << 4) | fo)
            cigar.append(((end - pos) << 4) | 0)
            if lo == 4:
                cigar.append((rs << 4) | 4)
            elif lo != 0:
                cigar.append((1 << 4) | lo)
            # make the query length of the CIGAR agree with l_seq
            q = ls + rs + (end - pos)
            if q != ql:
                ql = q
        name = b"r%d\0" % i
        l_seq = ql
        seq = b"\xff" * ((l_seq + 1) // 2)
        qual = b"\xff" * l_seq
        core = struct.pack("<iiBBHHHiiii", cid, pos, len(name), 30, _reg2bin(pos, max(end, pos + 1)), len(cigar), fl, l_seq, -1, -1,
                           int(rc.tlen[i]))
        body = core + name + b"".join(struct.pack("<I", c) for c in cigar) + seq + qual
        recs.append(struct.pack("<i", len(body)) + body)
        rec_meta.append((cid, pos, max(end, pos + 1), len(body) + 4))
    head = b"".join(parts)
    blocks = []
    with open(path, "wb") as fh:
        fh.write(bgzf_compress(head + b"".join(recs), level, blocks))
    if bai:
        voff = lambda u: (blocks[u // 0xFF00] << 16) | (u % 0xFF00)
        n_ref = len(rc.references)
        bins = [dict() for _ in range(n_ref)]
        lin = [dict() for _ in range(n_ref)]
        u = len(head)
        for cid, pos, end, size in rec_meta:
            if 0 <= cid < n_ref:
                b = _reg2bin(pos, end)
                beg, fin = voff(u), voff(u + size)
                lo, hi = bins[cid].get(b, (beg, fin))
                bins[cid][b] = (min(lo, beg), max(hi, fin))
                for w in range(pos >> 14, ((end - 1) >> 14) + 1):
                    lin[cid][w] = min(lin[cid].get(w, beg), beg)
            u += size
        out = [b"BAI\1", struct.pack("<i", n_ref)]
        for cid in range(n_ref):
            out.append(struct.pack("<i", len(bins[cid])))
            for b, (lo, hi) in sorted(bins[cid].items()):
                out.append(struct.pack("<IiQQ", b, 1, lo, hi))
            n_intv = (max(lin[cid]) + 1) if lin[cid] else 0
            out.append(struct.pack("<i", n_intv))
            last = 0
            for w in range(n_intv):
                last = lin[cid].get(w, last)
                out.append(struct.pack("<Q", last))
        with open(path + ".bai", "wb") as fh:
            fh.write(b"".join(out))


def write_bam_fast(path, rc, level=1):
    """Vectorised writer for large synthetic record sets (bench.py --cli: 5 M records): same records as write_bam, but without stored
    sequences (l_seq = 0, '*': the query length is then inferred from the CIGAR, tobias/utils/ngs.pyx:84) and with fixed-width read
    names, so that all records with the same number of CIGAR operations have one size and are laid out with numpy; BGZF blocks are
    compressed on a thread pool."""
    from concurrent.futures import ThreadPoolExecutor
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % (n, ln) for n, ln in zip(rc.references, rc.lengths))
    parts = [b"BAM\1", struct.pack("<i", len(text)), text.encode(), struct.pack("<i", len(rc.references))]
    for n, ln in zip(rc.references, rc.lengths):
        nm = n.encode() + b"\0"
        parts.append(struct.pack("<i", len(nm)) + nm + struct.pack("<i", ln))
    head = b"".join(parts)
    n = len(rc)
    fo, lo = rc.first_op.astype(np.int64), rc.last_op.astype(np.int64)
    unm = (rc.flag & FLAG_UNMAPPED) != 0
    has = (fo != NO_CIGAR) & ~unm
    lead, trail = has & (fo != 0), has & (lo != 0)
    n_cig = has.astype(np.int64) + lead + trail
    name_len = 12                                                       # 'r' + 10 digits + NUL
    size = 4 + 32 + name_len + 4 * n_cig
    off = np.concatenate([[0], np.cumsum(size)])
    buf = np.zeros(int(off[-1]), dtype=np.uint8)

    def put(col_off, arr, dt):
        a = np.ascontiguousarray(arr.astype(dt)).view(np.uint8).reshape(n, -1)
        for k in range(a.shape[1]):
            buf[off[:-1] + col_off + k] = a[:, k]
    end = np.maximum(rc.ref_end.astype(np.int64), rc.ref_start.astype(np.int64) + 1)
    beg = rc.ref_start.astype(np.int64)
    e1 = end - 1
    bins = np.zeros(n, dtype=np.int64)
    done = np.zeros(n, dtype=bool)
    for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        m = ~done & ((beg >> shift) == (e1 >> shift))
        bins[m] = base + (beg[m] >> shift)
        done |= m
    put(0, size - 4, "<i4")
    put(4, rc.chrom_id, "<i4")
    put(8, rc.ref_start, "<i4")
    put(12, np.full(n, name_len), "u1")
    put(13, np.full(n, 30), "u1")
    put(14, bins, "<u2")
    put(16, n_cig, "<u2")
    put(18, rc.flag, "<u2")
    put(20, np.zeros(n), "<i4")                                         # l_seq = 0
    put(24, np.full(n, -1), "<i4")
    put(28, np.full(n, -1), "<i4")
    put(32, rc.tlen, "<i4")
    digits = np.arange(n, dtype=np.int64)
    buf[off[:-1] + 36] = ord("r")
    for k in range(10):
        buf[off[:-1] + 37 + k] = ord("0") + (digits // 10 ** (9 - k)) % 10
    ls, rs = rc.left_soft.astype(np.int64), rc.right_soft.astype(np.int64)
    op0 = np.where(fo == 4, (ls << 4) | 4, (1 << 4) | fo)
    opm = (end - beg) << 4
    op2 = np.where(lo == 4, (rs << 4) | 4, (1 << 4) | lo)
    cig0 = off[:-1] + 36 + name_len
    for mask, vals, slot in ((lead, op0, np.zeros(n, np.int64)), (has, opm, lead.astype(np.int64)), (trail, op2, lead.astype(np.int64) + 1)):
        idx = np.flatnonzero(mask)
        a = np.ascontiguousarray(vals[idx].astype("<u4")).view(np.uint8).reshape(-1, 4)
        for k in range(4):
            buf[cig0[idx] + 4 * slot[idx] + k] = a[:, k]
    data = head + buf.tobytes()

    def block(a):
        chunk = data[a:a + 0xFF00]
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        comp = c.compress(chunk) + c.flush()
        return struct.pack("<BBBBIBBHBBHH", 0x1F, 0x8B, 8, 4, 0, 0, 0xFF, 6, 66, 67, 2, len(comp) + 25) + comp + \
            struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk))
    with ThreadPoolExecutor(min(32, os.cpu_count() or 1)) as ex, open(path, "wb") as fh:
        for blk in ex.map(block, range(0, len(data), 0xFF00), chunksize=64):
            fh.write(blk)
        fh.write(_EOF_BLOCK)
