"""
Minimal FASTA reader/writer (pysam / htslib are absent from the image).

Replaces the two things the hot path needs from `pysam.FastaFile` in the reference:
contig names + lengths (tobias/tools/atacorrect.py:143-146) and `fetch(chrom, start, end)`
(tobias/utils/sequences.pyx:388).  Contigs are returned as uint8 ASCII arrays so that they can be
handed to the C-ABI (`tb_genome_upload`) without a Python-level per-base loop.
"""
import ctypes
import mmap
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PACK = None


def _packer():
    """Native helper (hostpack.c, compiled with gcc on first use): ASCII bases -> 2-bit words + N mask."""
    global _PACK
    if _PACK is None:
        so, src = os.path.join(_HERE, "_hostpack.so"), os.path.join(_HERE, "hostpack.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["gcc", "-O3", "-fPIC", "-shared", "-o", so, src])
        _PACK = ctypes.CDLL(so)
    return _PACK


def pack_ascii(ascii_arr, seq2=None, nmask=None):
    """(seq2 uint32[(n + 15) // 16], nmask uint32[(n + 31) // 32]) of an ASCII base array: the device genome layout
    (include/tobias_b200.h, tb_genome_upload_packed).  Optional preallocated (e.g. pinned) outputs."""
    a = np.ascontiguousarray(ascii_arr, dtype=np.uint8)
    n = int(a.shape[0])
    seq2 = np.empty((n + 15) // 16, dtype=np.uint32) if seq2 is None else seq2
    nmask = np.empty((n + 31) // 32, dtype=np.uint32) if nmask is None else nmask
    assert seq2.dtype == np.uint32 and nmask.dtype == np.uint32 and seq2.size >= (n + 15) // 16 and nmask.size >= (n + 31) // 32
    vp = ctypes.c_void_p
    from .bam import parallel_ranges
    lib = _packer()
    parallel_ranges(lambda g0, g1: lib.tb_pack_ascii(a.ctypes.data_as(vp), ctypes.c_int64(n), seq2.ctypes.data_as(vp), nmask.ctypes.data_as(vp),
                                                     ctypes.c_int64(g0), ctypes.c_int64(g1)), 0, (n + 31) // 32, 1 << 15)
    return seq2, nmask


def unpack_codes(seq2, nmask, n):
    out = np.empty(int(n), dtype=np.uint8)
    vp = ctypes.c_void_p
    _packer().tb_unpack_codes(np.ascontiguousarray(seq2).ctypes.data_as(vp), np.ascontiguousarray(nmask).ctypes.data_as(vp),
                              ctypes.c_int64(int(n)), out.ctypes.data_as(vp))
    return out


class PackedGenome:
    """The contigs of a FASTA file in the device layout, cached next to it as `<fasta>.tb2bit` (like a .fai: built on first use, reused
    while it is newer than the FASTA).  File: magic, contig count, per contig (name length, name, n_bases), then the word arrays."""
    MAGIC = b"TB2BIT01"

    def __init__(self, names, lengths, seq2, nmask):
        self.references, self.lengths, self.seq2, self.nmask = list(names), [int(x) for x in lengths], seq2, nmask

    @staticmethod
    def from_fasta(fasta, names=None, cache=True):
        path = fasta.filename + ".tb2bit"
        if cache and os.path.exists(path) and os.path.getmtime(path) >= os.path.getmtime(fasta.filename):
            pg = PackedGenome.load(path)
            if names is None or all(n in pg.references for n in names):
                return pg if names is None else pg.subset(names)
        allnames = list(fasta.references)
        seq2, nmask = [], []
        for n in allnames:
            a, b = pack_ascii(fasta.fetch_array(n))
            seq2.append(a)
            nmask.append(b)
        pg = PackedGenome(allnames, [fasta.get_reference_length(n) for n in allnames], seq2, nmask)
        if cache:
            try:
                pg.save(path)
            except OSError:
                pass                      # read-only genome directory: pack again next time
        return pg if names is None else pg.subset(names)

    def subset(self, names):
        idx = [self.references.index(n) for n in names]
        return PackedGenome([self.references[i] for i in idx], [self.lengths[i] for i in idx], [self.seq2[i] for i in idx],
                            [self.nmask[i] for i in idx])

    def save(self, path):
        tmp = path + ".tmp%d" % os.getpid()
        with open(tmp, "wb") as fh:
            fh.write(self.MAGIC + np.int64(len(self.references)).tobytes())
            for n, ln in zip(self.references, self.lengths):
                b = n.encode()
                fh.write(np.int64(len(b)).tobytes() + b + np.int64(ln).tobytes())
            for a, b in zip(self.seq2, self.nmask):
                fh.write(a.tobytes())
                fh.write(b.tobytes())
        os.replace(tmp, path)

    @staticmethod
    def load(path):
        buf = np.memmap(path, dtype=np.uint8, mode="r")
        if bytes(buf[:8]) != PackedGenome.MAGIC:
            raise ValueError("%s is not a tobias_b200 packed genome" % path)
        off = 8
        n = int(np.frombuffer(buf[off:off + 8], dtype=np.int64)[0])
        off += 8
        names, lengths = [], []
        for _ in range(n):
            ln = int(np.frombuffer(buf[off:off + 8], dtype=np.int64)[0])
            names.append(bytes(buf[off + 8:off + 8 + ln]).decode())
            lengths.append(int(np.frombuffer(buf[off + 8 + ln:off + 16 + ln], dtype=np.int64)[0]))
            off += 16 + ln
        seq2, nmask = [], []
        for ln in lengths:
            a, b = 4 * ((ln + 15) // 16), 4 * ((ln + 31) // 32)
            seq2.append(np.frombuffer(buf[off:off + a], dtype=np.uint32))
            nmask.append(np.frombuffer(buf[off + a:off + a + b], dtype=np.uint32))
            off += a + b
        return PackedGenome(names, lengths, seq2, nmask)


class FastaFile:
    def __init__(self, path):
        self.filename = path
        self._fh = open(path, "rb")
        size = os.fstat(self._fh.fileno()).st_size
        self._mm = mmap.mmap(self._fh.fileno(), 0, access=mmap.ACCESS_READ) if size > 0 else b""
        self._buf = np.frombuffer(self._mm, dtype=np.uint8)
        self._index = {}          # name -> (length, offset, line_bases, line_width)
        self.references = []
        fai = path + ".fai"
        if os.path.exists(fai) and os.path.getmtime(fai) >= os.path.getmtime(path):
            self._read_fai(fai)
        else:
            self._build_index()
        self.lengths = [self._index[r][0] for r in self.references]
        self.nreferences = len(self.references)

    # -------------------------------------------------------------- index
    def _read_fai(self, fai):
        with open(fai) as fh:
            for line in fh:
                c = line.rstrip("\n").split("\t")
                if len(c) < 5:
                    continue
                self.references.append(c[0])
                self._index[c[0]] = (int(c[1]), int(c[2]), int(c[3]), int(c[4]))

    def _build_index(self):
        buf = self._buf
        gt = np.flatnonzero(buf == ord(">"))
        gt = gt[(gt == 0) | (buf[np.maximum(gt, 1) - 1] == ord("\n"))]      # a header starts a line; '>' inside a description is text
        nl = np.flatnonzero(buf == ord("\n"))
        total = buf.shape[0]
        for k, h in enumerate(gt):
            e = nl[np.searchsorted(nl, h)] if nl.size and np.searchsorted(nl, h) < nl.size else total
            name = bytes(buf[h + 1:e]).decode().split()[0] if e > h + 1 else ""
            seq_start = min(e + 1, total)
            seq_end = int(gt[k + 1]) if k + 1 < gt.size else total
            a = np.searchsorted(nl, seq_start)
            b = np.searchsorted(nl, seq_end)
            n_newlines = b - a
            first_nl = int(nl[a]) if a < b else seq_end
            line_bases = first_nl - seq_start
            line_width = line_bases + 1
            # tolerate \r\n
            if line_bases > 0 and buf[first_nl - 1] == ord("\r"):
                line_bases -= 1
                n_cr = n_newlines
            else:
                n_cr = 0
            length = (seq_end - seq_start) - n_newlines - n_cr
            self.references.append(name)
            self._index[name] = (int(length), int(seq_start), int(max(line_bases, 1)), int(max(line_width, 2)))

    def write_fai(self):
        with open(self.filename + ".fai", "w") as fh:
            for r in self.references:
                ln, off, lb, lw = self._index[r]
                fh.write("%s\t%d\t%d\t%d\t%d\n" % (r, ln, off, lb, lw))

    # -------------------------------------------------------------- access
    def get_reference_length(self, chrom):
        return self._index[chrom][0]

    def fetch_array(self, chrom, start=None, end=None):
        """ASCII bytes (uint8 array, newlines removed) of chrom[start:end]."""
        length, off, lb, lw = self._index[chrom]
        start = 0 if start is None else max(0, int(start))
        end = length if end is None else min(length, int(end))
        if end <= start:
            return np.zeros(0, dtype=np.uint8)
        first_line, last_line = start // lb, (end - 1) // lb
        b0 = off + first_line * lw
        b1 = min(off + (last_line + 1) * lw, self._buf.shape[0])
        raw = self._buf[b0:b1]
        nlines = last_line - first_line + 1
        if raw.shape[0] == nlines * lw:
            seq = raw.reshape(nlines, lw)[:, :lb].reshape(-1)
        else:  # last line shorter than line_width (end of contig / file)
            full = (raw.shape[0] // lw)
            head = raw[:full * lw].reshape(full, lw)[:, :lb].reshape(-1)
            tail = raw[full * lw:]
            tail = tail[(tail != ord("\n")) & (tail != ord("\r"))]
            seq = np.concatenate([head, tail])
        s = start - first_line * lb
        return np.ascontiguousarray(seq[s:s + (end - start)])

    def fetch(self, chrom, start=None, end=None):
        return self.fetch_array(chrom, start, end).tobytes().decode("ascii")

    def close(self):
        self._buf = None
        try:
            if self._mm:
                self._mm.close()
        except (BufferError, ValueError):
            pass
        self._fh.close()


def write_fasta(path, contigs, line_width=60):
    """contigs: iterable of (name, uint8 ASCII array or str).  Also writes the .fai."""
    fai = []
    with open(path, "wb") as fh:
        for name, seq in contigs:
            if isinstance(seq, str):
                seq = np.frombuffer(seq.encode("ascii"), dtype=np.uint8)
            header = (">%s\n" % name).encode()
            fh.write(header)
            off = fh.tell()
            n = seq.shape[0]
            full = n // line_width
            if full:
                block = np.empty((full, line_width + 1), dtype=np.uint8)
                block[:, :line_width] = seq[:full * line_width].reshape(full, line_width)
                block[:, line_width] = ord("\n")
                fh.write(block.tobytes())
            if n % line_width:
                fh.write(seq[full * line_width:].tobytes() + b"\n")
            fai.append((name, n, off, line_width, line_width + 1))
    with open(path + ".fai", "w") as fh:
        for rec in fai:
            fh.write("%s\t%d\t%d\t%d\t%d\n" % rec)
