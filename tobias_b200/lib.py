"""
ctypes binding of libtobias_b200.so (C-ABI declared in include/tobias_b200.h).

`load()` loads the nvcc-built library from tobias_b200/csrc/ and raises if it is missing or if no
CUDA device can be initialised -- there is no CPU fallback anywhere in this package.  `Context` wraps
one `tb_ctx` (one GPU, one stream set; one process per GPU) with numpy-typed methods.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "csrc", "libtobias_b200.so")

TB_READ_REVERSE, TB_READ_5P_MATCH, TB_READ_SKIP = 1, 2, 4
SCORE_KINDS = {"footprint": 0, "sum": 1, "mean": 2, "none": 3, "FOS": 4}
TRACKS = ("uncorrected", "bias", "expected", "corrected")


class TobiasB200Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__("libtobias_b200 error %d: %s" % (code, message))
        self.code = code


class ScoreParams(C.Structure):
    _fields_ = [("kind", C.c_int), ("absolute", C.c_int), ("use_min", C.c_int), ("use_max", C.c_int),
                ("min_limit", C.c_double), ("max_limit", C.c_double), ("window", C.c_int), ("smooth", C.c_int),
                ("flank_min", C.c_int), ("flank_max", C.c_int), ("fp_min", C.c_int), ("fp_max", C.c_int),
                ("flank", C.c_int), ("as_f64", C.c_int)]


_vp, _i, _i64, _d = C.c_void_p, C.c_int, C.c_int64, C.c_double
_PROTOS = {
    "tb_abi_version": (_i, []),
    "tb_init": (_i, [_i, C.POINTER(_vp)]),
    "tb_destroy": (None, [_vp]),
    "tb_last_error": (C.c_char_p, [_vp]),
    "tb_device_info": (_i, [_vp, C.POINTER(_i), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i)]),
    "tb_last_kernel_ms": (_i, [_vp, C.POINTER(C.c_float)]),
    "tb_launch_count": (_i, [_vp, C.POINTER(_i64)]),
    "tb_set_option": (_i, [_vp, C.c_char_p, _i64]),
    "tb_timer_begin": (_i, [_vp]),
    "tb_timer_end": (_i, [_vp, C.POINTER(C.c_float)]),
    "tb_host_alloc": (_i, [_vp, _i64, C.POINTER(_vp)]),
    "tb_host_free": (_i, [_vp, _vp]),
    "tb_genome_create": (_i, [_vp, _i, _vp]),
    "tb_genome_upload": (_i, [_vp, _i, _i64, _vp, _i64]),
    "tb_genome_upload_async": (_i, [_vp, _i, _i64, _vp, _i64]),
    "tb_genome_upload_packed": (_i, [_vp, _i, _vp, _vp, _i64, _i]),
    "tb_genome_wait": (_i, [_vp]),
    "tb_genome_fetch": (_i, [_vp, _i, _i64, _i64, _vp]),
    "tb_reads_upload": (_i, [_vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "tb_reads_upload_compact": (_i, [_vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "tb_genome_upload_packed_runs": (_i, [_vp, _i, _vp, _i64, _vp, _vp, _i64, _i]),
    "tb_count_reads": (_i, [_vp, _i64, _vp, _vp, _vp, _i, _i, C.POINTER(_i64)]),
    "tb_pileup": (_i, [_vp, _i64, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "tb_read_signal": (_i, [_vp, _i64, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "tb_bias_counts": (_i, [_vp, _i64, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "tb_nccl_unique_id": (_i, [_vp]),
    "tb_nccl_init": (_i, [_vp, _vp, _i, _i]),
    "tb_allreduce_i64": (_i, [_vp, _vp, _i64]),
    "tb_allreduce_f64": (_i, [_vp, _vp, _i64]),
    "tb_set_bias_model": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "tb_bias_model_info": (_i, [_vp, _i, _vp, _vp, _vp]),
    "tb_score_sequence": (_i, [_vp, _i, _i, _i64, _i64, _vp]),
    "tb_correct_run": (_i, [_vp, _i64, _vp, _vp, _vp, _i, _i, _d, _i, _i, _i]),
    "tb_stage_ms": (_i, [_vp, C.POINTER(C.c_float), _i]),
    "tb_correct_download": (_i, [_vp, _i, _i, _vp, _vp]),
    "tb_correct_download_begin": (_i, [_vp, _i, _i, _vp, _vp]),
    "tb_correct_download_wait": (_i, [_vp]),
    "tb_correct_set_sink": (_i, [_vp, _i, _i, _vp, _vp]),
    "tb_score_set_sink": (_i, [_vp, _vp]),
    "tb_score_sink_wait": (_i, [_vp]),
    "tb_correct_fit_trace": (_i, [_vp, C.POINTER(_i64), _vp]),
    "tb_signal_upload": (_i, [_vp, _vp, _i64, _vp]),
    "tb_score_run": (_i, [_vp, C.POINTER(ScoreParams)]),
    "tb_score_download": (_i, [_vp, _vp]),
    "tb_score_gather": (_i, [_vp, _i64, _vp, _vp]),
    "tb_score_regions": (_i, [_vp, _vp, _i64, _vp, C.POINTER(ScoreParams), _vp]),
    "tb_signal_from_corrected": (_i, [_vp, _i]),
    "tb_signal_from_corrected_regions": (_i, [_vp, _i, _i64, _vp, _vp, _vp]),
    "tb_rolling": (_i, [_vp, _vp, _i64, _i, _i, _vp]),
    "tb_signal_summary": (_i, [_vp, _i, _vp]),
    "tb_signal_aggregate": (_i, [_vp, _vp, _vp, _i, _vp]),
    "tb_add_kmers": (_i, [_vp, _i, _i, _i, _i64, _vp, _vp, _vp, _vp, _vp]),
}

_lib = None
_lib_path = None


def exported_symbols():
    return sorted(_PROTOS)


def load(path=None):
    """Load the C-ABI library (default: the in-tree CUDA build) and declare every prototype.  Raises when the
    library is missing: build it with `python -m tobias_b200.build` (or __graft_entry__.build())."""
    global _lib, _lib_path
    if path is None and _lib is not None:
        return _lib                      # whatever was loaded first stays (the CUDA build unless a test loaded the emulator)
    path = os.path.abspath(path or DEFAULT_LIB)
    if _lib is not None and _lib_path == path:
        return _lib
    if not os.path.exists(path):
        raise FileNotFoundError("%s not found: the CUDA extension is not built (python -m tobias_b200.build); "
                                "tobias_b200 has no CPU fallback" % path)
    lib = C.CDLL(path)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib, _lib_path = lib, path
    return lib


def loaded_path():
    return _lib_path


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_vp)


def _arr(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class Context:
    """One tb_ctx: a GPU-resident genome + read set + bias model and the kernels that run over them."""

    def __init__(self, device=0, lib_path=None):
        self._lib = load(lib_path)
        h = _vp()
        rc = self._lib.tb_init(int(device), C.byref(h))
        if rc != 0:
            raise TobiasB200Error(rc, (self._lib.tb_last_error(None) or b"").decode())
        self._h = h
        self.device = device
        self.n_contigs = 0
        self.contig_lengths = []

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "_h", None):
            for p in getattr(self, "_pinned", []):
                self._lib.tb_host_free(self._h, p)
            self._pinned = []
            self._lib.tb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != 0:
            raise TobiasB200Error(rc, (self._lib.tb_last_error(self._h) or b"").decode())

    def device_info(self):
        sm, fr, tot, emu = _i(), _i64(), _i64(), _i()
        self._ck(self._lib.tb_device_info(self._h, C.byref(sm), C.byref(fr), C.byref(tot), C.byref(emu)))
        return {"sm_count": sm.value, "free_bytes": fr.value, "total_bytes": tot.value, "is_emulator": bool(emu.value)}

    def last_kernel_ms(self):
        ms = C.c_float()
        self._ck(self._lib.tb_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value

    def launch_count(self):
        n = _i64()
        self._ck(self._lib.tb_launch_count(self._h, C.byref(n)))
        return n.value

    def set_option(self, key, value):
        """Kernel-variant / chunk-size switch (comparison runs and tests); value < 0 restores the default."""
        self._ck(self._lib.tb_set_option(self._h, key.encode(), int(value)))

    def timer_begin(self):
        self._ck(self._lib.tb_timer_begin(self._h))

    def timer_end(self):
        ms = C.c_float()
        self._ck(self._lib.tb_timer_end(self._h, C.byref(ms)))
        return ms.value

    def pinned_empty(self, shape, dtype):
        """numpy array backed by page-locked host memory owned by this context (freed on close)."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        p = _vp()
        self._ck(self._lib.tb_host_alloc(self._h, n, C.byref(p)))
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(p)
        buf = (C.c_uint8 * max(n, 1)).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    # ------------------------------------------------------------------ genome / reads
    def genome_create(self, lengths):
        lens = _arr(lengths, np.int64)
        self._ck(self._lib.tb_genome_create(self._h, int(lens.shape[0]), _ptr(lens)))
        self.n_contigs = int(lens.shape[0])
        self.contig_lengths = [int(x) for x in lens]

    def genome_upload(self, contig, ascii_arr, offset=0):
        a = _arr(ascii_arr, np.uint8)
        self._ck(self._lib.tb_genome_upload(self._h, int(contig), int(offset), _ptr(a), int(a.shape[0])))

    def genome_upload_async(self, contig, ascii_arr, offset=0):
        """Enqueues the upload and returns; `ascii_arr` (ideally from pinned_empty) must stay alive and untouched until genome_wait()."""
        a = _arr(ascii_arr, np.uint8)
        self._pending_uploads = getattr(self, "_pending_uploads", [])
        self._pending_uploads.append(a)
        self._ck(self._lib.tb_genome_upload_async(self._h, int(contig), int(offset), _ptr(a), int(a.shape[0])))

    def genome_upload_packed(self, contig, seq2, nmask, n_bases, wait=True):
        """A contig in the device layout (tobias_b200.io.fasta.pack_ascii / PackedGenome).  wait=False: enqueued on the copy stream; the
        arrays (ideally pinned) must stay alive and untouched until genome_wait()."""
        a, b = _arr(seq2, np.uint32), _arr(nmask, np.uint32)
        if not wait:
            self._pending_uploads = getattr(self, "_pending_uploads", [])
            self._pending_uploads += [a, b]
        self._ck(self._lib.tb_genome_upload_packed(self._h, int(contig), _ptr(a), _ptr(b), int(n_bases), int(not wait)))

    def genome_upload_packed_runs(self, contig, seq2, run_start, run_end, n_bases, wait=True):
        """A contig as 2-bit words plus the runs of N positions [run_start, run_end) (mask_runs): the device builds the mask."""
        a, rs, re_ = _arr(seq2, np.uint32), _arr(run_start, np.int64), _arr(run_end, np.int64)
        if not wait:
            self._pending_uploads = getattr(self, "_pending_uploads", [])
            self._pending_uploads += [a, rs, re_]
        self._ck(self._lib.tb_genome_upload_packed_runs(self._h, int(contig), _ptr(a), int(rs.shape[0]), _ptr(rs), _ptr(re_), int(n_bases),
                                                        int(not wait)))

    def genome_wait(self):
        self._ck(self._lib.tb_genome_wait(self._h))
        self._pending_uploads = []

    def genome_load(self, contigs):
        """contigs: list of uint8 ASCII arrays, one per contig, in contig-id order."""
        self.genome_create([len(c) for c in contigs])
        for i, c in enumerate(contigs):
            self.genome_upload(i, c)

    def genome_fetch(self, contig, start, n):
        out = np.empty(int(n), dtype=np.uint8)
        self._ck(self._lib.tb_genome_fetch(self._h, int(contig), int(start), int(n), _ptr(out)))
        return out

    def reads_upload(self, contig, ref_start, ref_end, clip5, flags):
        c, s, e = _arr(contig, np.int32), _arr(ref_start, np.int32), _arr(ref_end, np.int32)
        k, f = _arr(clip5, np.int32), _arr(flags, np.uint8)
        self._ck(self._lib.tb_reads_upload(self._h, int(c.shape[0]), _ptr(c), _ptr(s), _ptr(e), _ptr(k), _ptr(f)))

    def reads_upload_compact(self, rec_off, ref_start, ref_len, clip5, flags):
        """Compact record columns (compact_read_columns): 9 bytes per record over the bus instead of 17."""
        o, s, ln = _arr(rec_off, np.int64), _arr(ref_start, np.int32), _arr(ref_len, np.uint16)
        k, f = _arr(clip5, np.uint16), _arr(flags, np.uint8)
        self._ck(self._lib.tb_reads_upload_compact(self._h, int(s.shape[0]), _ptr(o), _ptr(s), _ptr(ln), _ptr(k), _ptr(f)))

    def reads_upload_columns(self, rc):
        """Upload a tobias_b200.reads.ReadColumns."""
        flags = (rc.is_reverse.astype(np.uint8) * TB_READ_REVERSE
                 | rc.five_prime_is_match().astype(np.uint8) * TB_READ_5P_MATCH
                 | (~rc.usable()).astype(np.uint8) * TB_READ_SKIP)
        self.reads_upload(rc.chrom_id, rc.ref_start, rc.ref_end, rc.clip5(), flags)

    # ------------------------------------------------------------------ K1
    @staticmethod
    def _regions(contig, start, end):
        c, s, e = _arr(contig, np.int32), _arr(start, np.int64), _arr(end, np.int64)
        if not (c.shape == s.shape == e.shape) or c.ndim != 1:
            raise ValueError("contig/start/end must be 1-D arrays of equal length")
        return c, s, e

    def count_reads(self, contig, start, end, read_shift=(4, -5)):
        c, s, e = self._regions(contig, start, end)
        out = _i64()
        self._ck(self._lib.tb_count_reads(self._h, int(c.shape[0]), _ptr(c), _ptr(s), _ptr(e), int(read_shift[0]),
                                          int(read_shift[1]), C.byref(out)))
        return out.value

    def pileup(self, contig, start, end, read_shift=(4, -5), list_semantics=False):
        """list_semantics: ReadList.signal (ngs.pyx:186-199) -- every resident record, start < cut <= end."""
        c, s, e = self._regions(contig, start, end)
        total = int((e - s).sum())
        fwd, rev = np.zeros(total, dtype=np.uint32), np.zeros(total, dtype=np.uint32)
        fn = self._lib.tb_read_signal if list_semantics else self._lib.tb_pileup
        self._ck(fn(self._h, int(c.shape[0]), _ptr(c), _ptr(s), _ptr(e), int(read_shift[0]),
                                     int(read_shift[1]), _ptr(fwd), _ptr(rev)))
        return fwd, rev

    # ------------------------------------------------------------------ K2
    def bias_counts(self, contig, start, end, k_flank=12, bg_shift=100, read_shift=(4, -5), cap=10, is_dwm=True):
        c, s, e = self._regions(contig, start, end)
        L = 2 * k_flank + 1
        shape = (2, 2, 5, 5, L, L) if is_dwm else (2, 2, 5, L)
        counts = np.zeros(shape, dtype=np.int64)
        scalars = np.zeros(5, dtype=np.int64)
        self._ck(self._lib.tb_bias_counts(self._h, int(c.shape[0]), _ptr(c), _ptr(s), _ptr(e), int(k_flank), int(bg_shift),
                                          int(read_shift[0]), int(read_shift[1]), int(cap), int(bool(is_dwm)),
                                          _ptr(counts), _ptr(scalars)))
        return counts, scalars

    def nccl_unique_id(self):
        buf = np.zeros(128, dtype=np.uint8)
        rc = self._lib.tb_nccl_unique_id(_ptr(buf))
        if rc != 0:
            raise TobiasB200Error(rc, (self._lib.tb_last_error(None) or b"").decode())
        return buf

    def nccl_init(self, unique_id, rank, world_size):
        uid = _arr(unique_id, np.uint8)
        self._ck(self._lib.tb_nccl_init(self._h, _ptr(uid), int(rank), int(world_size)))

    def allreduce(self, buf):
        """In-place sum of an int64 / float64 host array over the ranks (NCCL)."""
        if buf.dtype == np.int64:
            self._ck(self._lib.tb_allreduce_i64(self._h, _ptr(buf), int(buf.size)))
        elif buf.dtype == np.float64:
            self._ck(self._lib.tb_allreduce_f64(self._h, _ptr(buf), int(buf.size)))
        else:
            raise TypeError("allreduce supports int64 and float64 arrays")
        return buf

    # ------------------------------------------------------------------ K3
    def set_bias_model(self, strand, is_dwm, L, bias_pwm_log, bg_pwm_log=None, bias_dwm_log=None, bg_dwm_log=None):
        t = [None if x is None else _arr(x, np.float64) for x in (bias_pwm_log, bg_pwm_log, bias_dwm_log, bg_dwm_log)]
        self._ck(self._lib.tb_set_bias_model(self._h, int(strand), int(bool(is_dwm)), int(L), _ptr(t[0]), _ptr(t[1]),
                                             _ptr(t[2]), _ptr(t[3])))

    def bias_model_info(self, strand):
        """(is_dwm, L, product_form) of the model tb_set_bias_model stored for a strand."""
        a, b, c = C.c_int(-1), C.c_int(0), C.c_int(0)
        self._ck(self._lib.tb_bias_model_info(self._h, int(strand), C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def score_sequence(self, strand, contig, start, end):
        out = np.empty(int(end - start), dtype=np.float64)
        self._ck(self._lib.tb_score_sequence(self._h, int(strand), int(contig), int(start), int(end), _ptr(out)))
        return out

    def stage_ms(self, n=4):
        buf = (C.c_float * n)()
        self._ck(self._lib.tb_stage_ms(self._h, buf, n))
        return [float(x) for x in buf]

    def correct_run(self, contig, start, end, k_flank=12, window=100, correction_factor=1.0, read_shift=(4, -5),
                    lm_exact_order=True):
        c, s, e = self._regions(contig, start, end)
        self._ck(self._lib.tb_correct_run(self._h, int(c.shape[0]), _ptr(c), _ptr(s), _ptr(e), int(k_flank), int(window),
                                          float(correction_factor), int(read_shift[0]), int(read_shift[1]),
                                          int(lm_exact_order)))
        self._cor_total = int((e - s).sum())
        self._cor_L = 2 * k_flank + 1

    def correct_download(self, split_strands=False, as_f64=False, tracks=TRACKS, prepost=True, out=None, wait=True):
        """dict track -> array (split_strands=False) or track -> (forward, reverse); plus 'prepost' [2][3][5][L].
        wait=False: the copies are only enqueued (they overlap with kernels launched afterwards); the arrays are complete
        after correct_download_wait()."""
        dt = np.float64 if as_f64 else np.float32
        ptrs = ((_vp * 2) * 4)()
        out_bufs, out = out, {}          # `out`: optional dict track -> preallocated (e.g. pinned) array, strands combined
        for t, name in enumerate(TRACKS):
            if name not in tracks:
                continue
            if split_strands:
                a, b = np.empty(self._cor_total, dtype=dt), np.empty(self._cor_total, dtype=dt)
                ptrs[t][0], ptrs[t][1] = a.ctypes.data, b.ctypes.data
                out[name] = (a, b)
            else:
                a = out_bufs[name] if out_bufs is not None else np.empty(self._cor_total, dtype=dt)
                assert a.dtype == dt and a.size >= self._cor_total and a.flags.c_contiguous
                ptrs[t][0] = a.ctypes.data
                out[name] = a[:self._cor_total]
        pp = np.zeros((2, 3, 5, self._cor_L), dtype=np.float64) if prepost else None
        fn = self._lib.tb_correct_download if wait else self._lib.tb_correct_download_begin
        self._ck(fn(self._h, int(bool(split_strands)), int(bool(as_f64)), ptrs, _ptr(pp)))
        if prepost:
            out["prepost"] = pp
        self._pending_download = None if wait else (ptrs, out)         # keeps the destination arrays alive
        return out

    def correct_download_wait(self):
        self._ck(self._lib.tb_correct_download_wait(self._h))
        self._pending_download = None

    def correct_set_sink(self, out, n_out, k_flank=12, as_f64=False, prepost=True):
        """Streaming download (strands combined): `out` maps track names to preallocated (pinned) arrays of at least n_out elements; the
        NEXT correct_run copies its tracks there chunk by chunk while it computes.  Returns the dict of views (+ 'prepost'), complete
        after correct_download_wait()."""
        dt = np.float64 if as_f64 else np.float32
        ptrs = ((_vp * 2) * 4)()
        views = {}
        for t, name in enumerate(TRACKS):
            if name in out:
                a = out[name]
                assert a.dtype == dt and a.size >= n_out and a.flags.c_contiguous
                ptrs[t][0] = a.ctypes.data
                views[name] = a[:n_out]
        pp = np.zeros((2, 3, 5, 2 * k_flank + 1), dtype=np.float64) if prepost else None
        self._ck(self._lib.tb_correct_set_sink(self._h, 0, int(bool(as_f64)), ptrs, _ptr(pp)))
        if prepost:
            views["prepost"] = pp
        self._pending_download = (ptrs, views, out)                    # keeps the destination arrays alive
        return views

    def score_set_sink(self, out):
        """Streaming download of the NEXT score_run's output into `out` (preallocated, pinned for overlap); complete after score_sink_wait()."""
        assert out.flags.c_contiguous
        self._ck(self._lib.tb_score_set_sink(self._h, _ptr(out)))
        self._pending_score_sink = out

    def score_sink_wait(self):
        self._ck(self._lib.tb_score_sink_wait(self._h))
        self._pending_score_sink = None

    def correct_fit_trace(self):
        n = _i64()
        self._ck(self._lib.tb_correct_fit_trace(self._h, C.byref(n), None))
        out = np.zeros((2, n.value, 6), dtype=np.float64)
        if n.value:
            self._ck(self._lib.tb_correct_fit_trace(self._h, C.byref(n), _ptr(out)))
        return out

    # ------------------------------------------------------------------ K4
    @staticmethod
    def score_params(kind="footprint", flank=0, absolute=False, min_limit=None, max_limit=None, window=100, smooth=1,
                     flank_min=10, flank_max=30, fp_min=20, fp_max=50, as_f64=False):
        return ScoreParams(SCORE_KINDS[kind], int(bool(absolute)), int(min_limit is not None), int(max_limit is not None),
                           float(min_limit or 0.0), float(max_limit or 0.0), int(window), int(smooth), int(flank_min),
                           int(flank_max), int(fp_min), int(fp_max), int(flank), int(bool(as_f64)))

    def signal_upload(self, signal, ext_len):
        sig, ln = _arr(signal, np.float32), _arr(ext_len, np.int64)
        if int(ln.sum()) != sig.shape[0]:
            raise ValueError("signal length does not match sum(ext_len)")
        self._ck(self._lib.tb_signal_upload(self._h, _ptr(sig), int(ln.shape[0]), _ptr(ln)))
        self._sc_ext = ln

    def signal_from_corrected(self, flank, out_len):
        self._ck(self._lib.tb_signal_from_corrected(self._h, int(flank)))
        self._sc_ext = _arr(out_len, np.int64) + 2 * int(flank)

    def signal_from_corrected_regions(self, flank, contig, start, end):
        """Scoring regions of one's own choice over the corrected track of the last correct_run (0 where the track holds nothing)."""
        c, s, e = self._regions(contig, start, end)
        self._ck(self._lib.tb_signal_from_corrected_regions(self._h, int(flank), int(c.shape[0]), _ptr(c), _ptr(s), _ptr(e)))
        self._sc_ext = (e - s) + 2 * int(flank)

    def score_run(self, params):
        self._ck(self._lib.tb_score_run(self._h, C.byref(params)))
        self._sc_params = params

    def score_download(self, out=None):
        p = self._sc_params
        total = int((self._sc_ext - 2 * p.flank).sum())
        dt = np.float64 if p.as_f64 else np.float32
        if out is None:
            out = np.empty(total, dtype=dt)
        assert out.dtype == dt and out.size >= total and out.flags.c_contiguous
        out = out[:total]
        self._ck(self._lib.tb_score_download(self._h, _ptr(out)))
        return out

    def score_gather(self, index):
        """Scores of the last score_run at positions of its concatenated output space (float64), gathered on the device."""
        idx = _arr(index, np.int64)
        out = np.empty(idx.shape[0], dtype=np.float64)
        self._ck(self._lib.tb_score_gather(self._h, int(idx.shape[0]), _ptr(idx), _ptr(out)))
        return out

    SUMMARY_OPS = {"start": 0, "mid": 1, "end": 2, "min": 3, "max": 4, "mean": 5, "sum": 6}

    def signal_summary(self, op):
        """Per-region summary of the uploaded signal (ScoreBed's score functions): float64 [n_regions]."""
        out = np.empty(self._sc_ext.shape[0], dtype=np.float64)
        self._ck(self._lib.tb_signal_summary(self._h, self.SUMMARY_OPS[op], _ptr(out)))
        return out

    def signal_aggregate(self, reverse=None, keep=None, log_transform=False):
        """Column means over the uploaded equal-width regions (PlotAggregate): float64 [width]."""
        out = np.empty(int(self._sc_ext[0]), dtype=np.float64)
        rv = None if reverse is None else _arr(reverse, np.uint8)
        kp = None if keep is None else _arr(keep, np.uint8)
        self._ck(self._lib.tb_signal_aggregate(self._h, _ptr(rv), _ptr(kp), int(bool(log_transform)), _ptr(out)))
        return out

    def score_regions(self, signal, ext_len, params):
        self.signal_upload(signal, ext_len)
        self.score_run(params)
        return self.score_download()

    def add_kmers(self, kmers, amounts, counts, neg_counts=None, is_dwm=True, is_background=False):
        """Batch form of SequenceMatrix.add_sequence / add_background: adds into `counts` (and `neg_counts`) in place, returns the total added."""
        k = _arr(kmers, np.int64).reshape(-1, counts.shape[-1])
        a = _arr(amounts, np.float64).reshape(-1)
        assert k.shape[0] == a.shape[0] and counts.dtype == np.float64 and counts.flags.c_contiguous
        total = np.zeros(1, dtype=np.float64)
        self._ck(self._lib.tb_add_kmers(self._h, int(counts.shape[-1]), int(bool(is_dwm)), int(bool(is_background)), int(k.shape[0]), _ptr(k),
                                        _ptr(a), _ptr(counts), _ptr(neg_counts), _ptr(total)))
        return float(total[0])

    def rolling(self, arr, w, operation="sum"):
        a = _arr(arr, np.float64)
        out = np.empty_like(a)
        op = {"sum": 0, "mean": 1, "max": 2, "min": 3, "prod": 4}[operation]
        self._ck(self._lib.tb_rolling(self._h, _ptr(a), int(a.shape[0]), int(w), op, _ptr(out)))
        return out



def compact_read_columns(contig, ref_start, ref_end, clip5, flags, n_contigs):
    """(rec_off, ref_start, ref_len u16, clip5 u16, flags) for Context.reads_upload_compact, or None when the records do not fit the
    compact form (not grouped by ascending contig, a contig outside the genome, a reference span or clip of 65536 or more)."""
    c = np.asarray(contig)
    if c.size and (c.min() < 0 or c.max() >= n_contigs or np.any(np.diff(c) < 0)):
        return None
    ln = np.asarray(ref_end, dtype=np.int64) - np.asarray(ref_start, dtype=np.int64)
    k = np.asarray(clip5)
    if c.size and (ln.min() < 0 or ln.max() > 65535 or k.min() < 0 or k.max() > 65535):
        return None
    rec_off = np.zeros(n_contigs + 1, dtype=np.int64)
    rec_off[1:] = np.cumsum(np.bincount(c, minlength=n_contigs))
    return rec_off, np.ascontiguousarray(ref_start, dtype=np.int32), ln.astype(np.uint16), k.astype(np.uint16), np.ascontiguousarray(flags, dtype=np.uint8)


def mask_runs(nmask, n_bases):
    """Runs [start, end) of set bits in a bit mask of n_bases positions (uint32 words, 32 positions per word, bit 0 first)."""
    w = np.ascontiguousarray(nmask, dtype=np.uint32)
    nz = np.flatnonzero(w)
    if nz.size == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    bits = np.unpackbits(w[nz].view(np.uint8).reshape(-1, 4), axis=1, bitorder="little")      # the non-zero words only
    pos = (nz[:, None].astype(np.int64) * 32 + np.arange(32, dtype=np.int64)[None, :])[bits.astype(bool)]
    pos = pos[pos < n_bases]
    if pos.size == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    brk = np.flatnonzero(np.diff(pos) != 1)
    starts = np.concatenate([pos[:1], pos[brk + 1]])
    ends = np.concatenate([pos[brk], pos[-1:]]) + 1
    return starts.astype(np.int64), ends.astype(np.int64)
