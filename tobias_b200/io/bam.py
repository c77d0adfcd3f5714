"""
Minimal BAM codec (pysam / htslib / samtools are absent from the image).

read_bam(path) -> ReadColumns: inflates the BGZF blocks (zlib, in a thread pool -- zlib releases the GIL) and parses
the alignment records with a small native helper (bamdecode.c, compiled with gcc on first use) into the columns the
hot path consumes.  The whole file is decoded once; region queries are binary searches on the sorted columns
(ReadColumns.fetch_index), so no .bai index is needed.

write_bam(path, ReadColumns): coordinate-sorted BAM of synthetic records (used by the tests and by the synthetic data
generator); sequence and qualities are placeholders ('N', 0xFF).  No .bai is written.
"""
import ctypes
import os
import struct
import subprocess
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from ..reads import ReadColumns, NO_CIGAR, FLAG_UNMAPPED

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_EOF_BLOCK = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def build_helper(force=False):
    so = os.path.join(_HERE, "_bamdecode.so")
    src = os.path.join(_HERE, "bamdecode.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", so, src])
    return so


def _helper():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_helper())
        _LIB.tb_bam_decode.restype = ctypes.c_int64
    return _LIB


# ------------------------------------------------------------------------------------------------ BGZF
def _bgzf_blocks(buf):
    """(payload offset, payload size, uncompressed size) of every BGZF block."""
    out = []
    off, n = 0, len(buf)
    while off + 18 <= n:
        if buf[off] != 0x1F or buf[off + 1] != 0x8B:
            raise ValueError("not a BGZF file (bad gzip magic at %d)" % off)
        xlen = struct.unpack_from("<H", buf, off + 10)[0]
        bsize = None
        p = off + 12
        while p < off + 12 + xlen:
            si1, si2, slen = buf[p], buf[p + 1], struct.unpack_from("<H", buf, p + 2)[0]
            if si1 == 66 and si2 == 67:
                bsize = struct.unpack_from("<H", buf, p + 4)[0]
            p += 4 + slen
        if bsize is None:
            raise ValueError("gzip member without BGZF extra field")
        total = bsize + 1
        isize = struct.unpack_from("<I", buf, off + total - 4)[0]
        out.append((off + 12 + xlen, total - xlen - 20, isize))
        off += total
    return out


def bgzf_decompress(buf, threads=None):
    blocks = [b for b in _bgzf_blocks(buf) if b[2] > 0]
    view = memoryview(buf)

    def inflate(b):
        return zlib.decompress(view[b[0]:b[0] + b[1]], -15)
    threads = threads or min(32, os.cpu_count() or 1)
    if len(blocks) > 64 and threads > 1:
        with ThreadPoolExecutor(threads) as ex:
            parts = list(ex.map(inflate, blocks, chunksize=256))
    else:
        parts = [inflate(b) for b in blocks]
    return b"".join(parts)


def bgzf_compress(data, level=6):
    out = []
    for a in range(0, len(data), 0xFF00):
        chunk = data[a:a + 0xFF00]
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        comp = c.compress(chunk) + c.flush()
        bsize = len(comp) + 25
        out.append(struct.pack("<BBBBIBBHBBHH", 0x1F, 0x8B, 8, 4, 0, 0, 0xFF, 6, 66, 67, 2, bsize) + comp +
                   struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
    out.append(_EOF_BLOCK)
    return b"".join(out)


# ------------------------------------------------------------------------------------------------ reader
def read_bam_header(path):
    """(references, lengths) -- what the driver reads first (tobias/tools/atacorrect.py:131-140)."""
    with open(path, "rb") as fh:
        head = fh.read(1 << 22)
    data = b""
    for (p, sz, _isz) in _bgzf_blocks(head[:len(head)]) if len(head) < (1 << 22) else _iter_blocks_prefix(head):
        data += zlib.decompress(head[p:p + sz], -15)
        parsed = _parse_header(data)
        if parsed is not None:
            return parsed[0], parsed[1]
    return read_bam(path, header_only=True)


def _iter_blocks_prefix(buf):
    off, n = 0, len(buf)
    while off + 18 <= n:
        xlen = struct.unpack_from("<H", buf, off + 10)[0]
        bsize = struct.unpack_from("<H", buf, off + 16)[0]
        total = bsize + 1
        if off + total > n:
            return
        yield (off + 12 + xlen, total - xlen - 20, 0)
        off += total


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


def read_bam(path, header_only=False):
    with open(path, "rb") as fh:
        raw = fh.read()
    data = bgzf_decompress(raw)
    del raw
    parsed = _parse_header(data)
    if parsed is None:
        raise ValueError("%s: truncated BAM header" % path)
    names, lens, p = parsed
    if header_only:
        return names, lens
    body = np.frombuffer(data, dtype=np.uint8)[p:]
    # upper bound on the record count: every record takes at least 36 bytes
    cap = body.shape[0] // 36 + 1
    cols = {"chrom_id": np.empty(cap, np.int32), "ref_start": np.empty(cap, np.int32), "ref_end": np.empty(cap, np.int32),
            "left_soft": np.empty(cap, np.int32), "right_soft": np.empty(cap, np.int32), "query_length": np.empty(cap, np.int32),
            "flag": np.empty(cap, np.uint16), "first_op": np.empty(cap, np.uint8), "last_op": np.empty(cap, np.uint8),
            "tlen": np.empty(cap, np.int32)}
    consumed = ctypes.c_int64(0)
    vp = ctypes.c_void_p
    n = _helper().tb_bam_decode(vp(body.ctypes.data), ctypes.c_int64(body.shape[0]), ctypes.c_int64(cap), ctypes.byref(consumed),
                                *[vp(cols[k].ctypes.data) for k in ("chrom_id", "ref_start", "ref_end", "left_soft", "right_soft",
                                                                   "query_length", "flag", "first_op", "last_op", "tlen")])
    if consumed.value != body.shape[0]:
        raise ValueError("%s: trailing bytes after the last complete BAM record (file truncated?)" % path)
    cols = {k: v[:n].copy() for k, v in cols.items()}
    keep = cols["chrom_id"] >= 0                      # records without a reference (unplaced unmapped reads) are never fetched
    if not keep.all():
        cols = {k: v[keep] for k, v in cols.items()}
    rc = ReadColumns(names, lens, **cols)
    if not rc.is_sorted():
        raise ValueError("%s is not coordinate sorted" % path)
    return rc


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


def write_bam(path, rc, level=4):
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % (n, ln) for n, ln in zip(rc.references, rc.lengths))
    parts = [b"BAM\1", struct.pack("<i", len(text)), text.encode(), struct.pack("<i", len(rc.references))]
    for n, ln in zip(rc.references, rc.lengths):
        nm = n.encode() + b"\0"
        parts.append(struct.pack("<i", len(nm)) + nm + struct.pack("<i", ln))
    recs = []
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
    with open(path, "wb") as fh:
        fh.write(bgzf_compress(b"".join(parts) + b"".join(recs), level))
