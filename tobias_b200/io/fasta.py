"""
Minimal FASTA reader/writer (pysam / htslib are absent from the image).

Replaces the two things the hot path needs from `pysam.FastaFile` in the reference:
contig names + lengths (tobias/tools/atacorrect.py:143-146) and `fetch(chrom, start, end)`
(tobias/utils/sequences.pyx:388).  Contigs are returned as uint8 ASCII arrays so that they can be
handed to the C-ABI (`tb_genome_upload`) without a Python-level per-base loop.
"""
import mmap
import os

import numpy as np


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
