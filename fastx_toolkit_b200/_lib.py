"""ctypes binding of include/fxg.h (one function per C entry point, same names and argument order)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FXG_OK = 0
QBINS = 109


class FxgError(RuntimeError):
    def __init__(self, code, detail=""):
        self.code = code
        super().__init__("libfxg error %d: %s" % (code, detail))


class Batch(C.Structure):
    """struct fxg_batch"""
    _fields_ = [("seq", C.c_void_p), ("qual", C.c_void_p), ("len", C.c_void_p),
                ("uniform_len", C.c_int32), ("stride", C.c_int32), ("n", C.c_int64)]


class BarcodeTable(C.Structure):
    """struct fxg_barcode_table (entries / entry_len are host arrays kept alive by the caller)"""
    _fields_ = [("entries", C.c_void_p), ("entry_len", C.c_void_p), ("n_entries", C.c_int32), ("barcode_len", C.c_int32),
                ("allowed_mismatches", C.c_int32)]


class Stage(C.Structure):
    """struct fxg_stage (op: 0 trim, 1 filter, 2 clip, 3 collapse)"""
    _fields_ = [("op", C.c_int32), ("a0", C.c_int32), ("a1", C.c_int32), ("clip", C.c_void_p), ("collapser", C.c_void_p)]


class ClipOpts(C.Structure):
    """struct fxg_clip_opts"""
    _fields_ = [("adapter", C.c_char_p), ("min_length", C.c_int32), ("keep_delta", C.c_int32),
                ("discard_non_clipped", C.c_int32), ("discard_clipped", C.c_int32), ("discard_unknown", C.c_int32),
                ("min_adapter_len", C.c_int32)]


class TextReport(C.Structure):
    """struct fxg_text_report"""
    _fields_ = [("n_records", C.c_int64), ("n_out_records", C.c_int64), ("consumed_bytes", C.c_int64), ("out_bytes", C.c_int64),
                ("max_len", C.c_int32), ("anomaly", C.c_int32), ("anomaly_record", C.c_int64),
                ("min_len", C.c_int32), ("reserved", C.c_int32), ("clip_class", C.c_int64 * 6), ("n_reads", C.c_int64), ("n_out_reads", C.c_int64),
                ("raw_out_bytes", C.c_int64), ("out_crc32_pure", C.c_uint32), ("deflated", C.c_uint32)]


class DCollapseReport(C.Structure):
    """struct fxg_dcollapse_report"""
    _fields_ = [("n_unique", C.c_int64), ("first_bad_read", C.c_int64), ("n_reads_local", C.c_int64), ("rows_received", C.c_int64),
                ("n_unique_local", C.c_int64), ("bytes_sent", C.c_int64), ("ms", C.c_float * 5), ("reserved", C.c_float)]


class Report(C.Structure):
    """struct fxg_report"""
    _fields_ = [("n_in", C.c_int64), ("n_out", C.c_int64), ("first_bad_read", C.c_int64), ("aux", C.c_int64 * 6)]


def lib_path():
    return os.path.join(_HERE, "libfxg.so")


def lib():
    """Load libfxg.so.  Missing library is a hard error (no fallback path exists)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not os.path.exists(p):
        raise FxgError(-1, "%s not built: run `make lib` (nvcc, sm_100a). There is no CPU fallback." % p)
    L = C.CDLL(p)
    vp, i32, i64, u64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_size_t
    BP, RP = C.POINTER(Batch), C.POINTER(Report)
    sig = {
        "fxg_init": (i32, [i32, C.POINTER(vp)]),
        "fxg_destroy": (None, [vp]),
        "fxg_strerror": (C.c_char_p, [i32]),
        "fxg_last_error": (C.c_char_p, [vp]),
        "fxg_device_info": (i32, [vp, C.POINTER(i32), C.POINTER(sz), C.POINTER(i32), C.POINTER(i32)]),
        "fxg_set_stream": (i32, [vp, vp]),
        "fxg_use_own_stream": (i32, [vp]),
        "fxg_sync": (i32, [vp]),
        "fxg_get_report": (i32, [vp, RP]),
        "fxg_report_reset": (i32, [vp]),
        "fxg_kernel_launches": (i64, [vp]),
        "fxg_alloc_pinned": (vp, [sz]),
        "fxg_free_pinned": (None, [vp]),
        "fxg_host_register": (i32, [vp, sz]),
        "fxg_host_unregister": (i32, [vp]),
        "fxg_alloc_device": (vp, [vp, sz]),
        "fxg_free_device": (None, [vp, vp]),
        "fxg_memcpy_h2d": (i32, [vp, vp, vp, sz]),
        "fxg_memcpy_d2h": (i32, [vp, vp, vp, sz]),
        "fxg_memset_dev": (i32, [vp, vp, i32, sz]),
        "fxg_set_tuning": (i32, [vp, i32, i32, i32]),
        "fxg_synth_dev": (i32, [vp, vp, vp, i64, i64, i64, C.c_int32, C.c_int32, u64, i32, i32]),
        "fxg_trim_dev": (i32, [vp, BP, i32, i32, i32, vp, i64]),
        "fxg_trim_host": (i32, [vp, BP, i32, i32, i32, vp, RP]),
        "fxg_filter_dev": (i32, [vp, BP, i32, i32, i32, vp, i64]),
        "fxg_filter_host": (i32, [vp, BP, i32, i32, i32, vp, RP]),
        "fxg_revcomp_dev": (i32, [vp, BP, i32, vp, vp, i64]),
        "fxg_revcomp_host": (i32, [vp, BP, i32, vp, vp, RP]),
        "fxg_stats_accum_dev": (i32, [vp, BP, i32, vp, C.c_int32, vp, i64]),
        "fxg_stats_accum_host": (i32, [vp, BP, i32, vp, C.c_int32, vp, RP]),
        "fxg_clip_dev": (i32, [vp, BP, vp, i32, C.POINTER(ClipOpts), vp, vp, vp, i64]),
        "fxg_clip_host": (i32, [vp, BP, vp, i32, C.POINTER(ClipOpts), vp, vp, RP]),
        "fxg_hash_dev": (i32, [vp, BP, vp]),
        "fxg_comm_init_all": (i32, [i32, C.POINTER(i32), C.POINTER(vp)]),
        "fxg_comm_allreduce_u64": (i32, [vp, C.POINTER(vp), sz]),
        "fxg_comm_unique_id": (i32, [vp]),
        "fxg_comm_init_rank": (i32, [i32, i32, i32, vp, C.POINTER(vp)]),
        "fxg_comm_nranks": (i32, [vp]),
        "fxg_comm_nlocal": (i32, [vp]),
        "fxg_comm_rank": (i32, [vp, i32]),
        "fxg_comm_device": (i32, [vp, i32]),
        "fxg_comm_set_stream": (i32, [vp, i32, vp, i32]),
        "fxg_comm_sync": (i32, [vp]),
        "fxg_comm_allgather": (i32, [vp, C.POINTER(vp), C.POINTER(vp), sz]),
        "fxg_comm_alltoallv": (i32, [vp, C.POINTER(vp), C.POINTER(i64), C.POINTER(i64), C.POINTER(vp), C.POINTER(i64), C.POINTER(i64), sz]),
        "fxg_comm_gatherv": (i32, [vp, C.POINTER(vp), C.POINTER(i64), C.POINTER(i64), vp, i32, sz]),
        "fxg_comm_bytes_sent": (i64, [vp]),
        "fxg_comm_collectives": (i64, [vp]),
        "fxg_dcollapse_new": (i32, [vp, C.c_int32, C.POINTER(vp)]),
        "fxg_dcollapse_free": (None, [vp]),
        "fxg_dcollapse_run": (i32, [vp, BP, C.POINTER(i64), C.POINTER(vp), C.POINTER(vp), i32, C.POINTER(DCollapseReport)]),
        "fxg_collapse_reserve": (i32, [vp, i64, C.c_int32]),
        "fxg_dcollapse_fetch_local": (i32, [vp, i32, vp, vp, vp, vp, vp]),
        "fxg_dcollapse_fetch_order": (i32, [vp, vp, vp, vp, vp]),
        "fxg_dcollapse_error": (C.c_char_p, [vp]),
        "fxg_dcollapse_launches": (i64, [vp]),
        "fxg_comm_free": (None, [vp]),
        "fxg_comm_error": (C.c_char_p, [vp]),
        "fxg_validate_dev": (i32, [vp, BP, i32, i64]),
        "fxg_validate_host": (i32, [vp, BP, i32, RP]),
        "fxg_mask_dev": (i32, [vp, BP, i32, i32, i32, vp, vp, i64]),
        "fxg_mask_host": (i32, [vp, BP, i32, i32, i32, vp, vp, RP]),
        "fxg_artifacts_dev": (i32, [vp, BP, i32, vp, i64]),
        "fxg_artifacts_host": (i32, [vp, BP, i32, vp, RP]),
        "fxg_barcode_dev": (i32, [vp, BP, vp, vp]),
        "fxg_barcode_host": (i32, [vp, BP, vp, vp, RP]),
        "fxg_pipeline_dev": (i32, [vp, BP, i32, vp, i32, vp, vp]),
        "fxg_has_n_dev": (i32, [vp, BP, i32, vp, i64]),
        "fxg_has_n_host": (i32, [vp, BP, i32, vp, RP]),
        "fxg_text_new": (i32, [vp, i32, sz, C.POINTER(vp)]),
        "fxg_text_free": (None, [vp]),
        "fxg_text_run_host": (i32, [vp, i32, vp, sz, i32, i32, i32, vp, C.POINTER(TextReport)]),
        "fxg_text_decide_host": (i32, [vp, i32, vp, sz, i32, i32, i32, vp, vp, C.POINTER(TextReport)]),
        "fxg_text_clip_host": (i32, [vp, vp, sz, i32, C.POINTER(ClipOpts), i32, i32, vp, C.POINTER(TextReport)]),
        "fxg_text_stats_host": (i32, [vp, vp, sz, i32, vp, C.c_int32, C.POINTER(TextReport)]),
        "fxg_text_set_format": (i32, [vp, i32]),
        "fxg_text_set_deflate": (i32, [vp, i32]),
        "fxg_crc32_concat": (C.c_uint32, [C.c_uint32, C.c_uint32, C.c_uint64]),
        "fxg_crc32_finish": (C.c_uint32, [C.c_uint32, C.c_uint64]),
        "fxg_text_collapse_host": (i32, [vp, vp, sz, i32, vp, i64, C.POINTER(TextReport)]),
        "fxg_text_numeric_chunks": (i64, [vp]),
        "fxg_text_fasta_chunks": (i64, [vp]),
        "fxg_text_error": (C.c_char_p, [vp]),
        "fxg_text_launches": (i64, [vp]),
        "fxg_collapse_new": (i32, [i32, i64, C.c_int32, C.POINTER(vp)]),
        "fxg_collapse_free": (None, [vp]),
        "fxg_collapse_add": (i32, [vp, BP, vp, vp, i64]),
        "fxg_collapse_add_next": (i32, [vp, BP]),
        "fxg_collapse_finish": (i32, [vp, i32, C.POINTER(i64), C.POINTER(i64)]),
        "fxg_collapse_fetch": (i32, [vp, vp, vp, vp, vp, vp]),
        "fxg_collapse_error": (C.c_char_p, [vp]),
        "fxg_collapse_launches": (i64, [vp]),
        "fxg_collapse_stride": (C.c_int32, [vp]),
        "fxg_collapse_order_dev": (i32, [i32, vp, vp, vp, i64, vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _LIB = L
    return L


def _ptr(x):
    """Device/host address of a torch tensor, numpy array, int or None."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "ctypes"):
        return x.ctypes.data
    raise TypeError(type(x))


class Collapser:
    """fxg_collapser: exact dedup + reference output order on one GPU."""

    def __init__(self, device, max_reads, stride):
        self.L = lib()
        h = C.c_void_p()
        rc = self.L.fxg_collapse_new(device, max_reads, stride, C.byref(h))
        if rc != FXG_OK:
            raise FxgError(rc, "fxg_collapse_new: " + self.L.fxg_strerror(rc).decode())
        self.h, self.stride, self.n_unique, self.first_bad = h, stride, 0, -1

    def _ck(self, rc):
        if rc != FXG_OK:
            raise FxgError(rc, self.L.fxg_collapse_error(self.h).decode() or self.L.fxg_strerror(rc).decode())

    def add(self, b, weight=None, first=None, index_base=0):
        self._ck(self.L.fxg_collapse_add(self.h, C.byref(b), _ptr(weight), _ptr(first), index_base))

    def finish(self, order=True):
        u, bad = C.c_int64(), C.c_int64()
        self._ck(self.L.fxg_collapse_finish(self.h, 1 if order else 0, C.byref(u), C.byref(bad)))
        self.n_unique, self.first_bad = u.value, bad.value
        return u.value

    def fetch(self, out_seq=None, out_len=None, out_count=None, out_first=None, out_hash=None):
        self._ck(self.L.fxg_collapse_fetch(self.h, _ptr(out_seq), _ptr(out_len), _ptr(out_count), _ptr(out_first), _ptr(out_hash)))

    def launches(self):
        return int(self.L.fxg_collapse_launches(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.L.fxg_collapse_free(self.h)
            self.h = None

    __del__ = close


class Comm:
    """fxg_comm: the native NCCL communicator over the GPUs this process drives.
    Comm.all(devices)              one process, several GPUs (ncclCommInitAll)
    Comm.rank(device, n, r, id)    one GPU per process; `id` = Comm.unique_id() made on one rank and handed to the others"""

    def __init__(self, handle):
        self.L, self.h = lib(), handle

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(128)
        rc = lib().fxg_comm_unique_id(buf)
        if rc != FXG_OK:
            raise FxgError(rc, "fxg_comm_unique_id: " + lib().fxg_comm_error(None).decode())
        return buf.raw

    @classmethod
    def all(cls, devices):
        h = C.c_void_p()
        arr = (C.c_int * len(devices))(*devices)
        rc = lib().fxg_comm_init_all(len(devices), arr, C.byref(h))
        if rc != FXG_OK:
            raise FxgError(rc, "fxg_comm_init_all: " + lib().fxg_comm_error(None).decode())
        return cls(h)

    @classmethod
    def rank(cls, device, nranks, rank, uid):
        h = C.c_void_p()
        rc = lib().fxg_comm_init_rank(device, nranks, rank, C.c_char_p(uid), C.byref(h))
        if rc != FXG_OK:
            raise FxgError(rc, "fxg_comm_init_rank: " + lib().fxg_comm_error(None).decode())
        return cls(h)

    def _ck(self, rc):
        if rc != FXG_OK:
            raise FxgError(rc, self.L.fxg_comm_error(self.h).decode() or self.L.fxg_strerror(rc).decode())

    @property
    def nranks(self):
        return self.L.fxg_comm_nranks(self.h)

    @property
    def nlocal(self):
        return self.L.fxg_comm_nlocal(self.h)

    def set_stream(self, local_index, stream_handle, adopt=True):
        self._ck(self.L.fxg_comm_set_stream(self.h, local_index, C.c_void_p(stream_handle), 1 if adopt else 0))

    def sync(self):
        self._ck(self.L.fxg_comm_sync(self.h))

    def allreduce_u64(self, bufs, count):
        """bufs: one device tensor / address per local GPU (u64 counters), summed in place across all ranks"""
        arr = (C.c_void_p * len(bufs))(*[_ptr(b) for b in bufs])
        self._ck(self.L.fxg_comm_allreduce_u64(self.h, arr, count))

    def bytes_sent(self):
        return int(self.L.fxg_comm_bytes_sent(self.h))

    def collectives(self):
        return int(self.L.fxg_comm_collectives(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.L.fxg_comm_free(self.h)
            self.h = None


class DCollapser:
    """fxg_dcollapse: fastx_collapser across the GPUs of a Comm (owner = std::hash mod nranks)."""

    def __init__(self, comm, stride):
        self.L, self.comm, self.stride = lib(), comm, stride
        h = C.c_void_p()
        rc = self.L.fxg_dcollapse_new(comm.h, stride, C.byref(h))
        if rc != FXG_OK:
            raise FxgError(rc, "fxg_dcollapse_new: " + self.L.fxg_strerror(rc).decode())
        self.h = h

    def _ck(self, rc):
        if rc != FXG_OK:
            raise FxgError(rc, self.L.fxg_dcollapse_error(self.h).decode() or self.L.fxg_strerror(rc).decode())

    def run(self, batches, index_bases, weights=None, root=0, firsts=None):
        """batches: one Batch per local GPU (device slabs); returns DCollapseReport"""
        n = len(batches)
        arr = (Batch * n)(*batches)
        bases = (C.c_int64 * n)(*index_bases)
        w = None
        if weights is not None:
            w = (C.c_void_p * n)(*[_ptr(x) for x in weights])
        f = None
        if firsts is not None:
            f = (C.c_void_p * n)(*[_ptr(x) for x in firsts])
        rep = DCollapseReport()
        self._ck(self.L.fxg_dcollapse_run(self.h, arr, bases, w, f, root, C.byref(rep)))
        return rep

    def fetch_local(self, local_index, out_seq=None, out_len=None, out_count=None, out_first=None, out_hash=None):
        self._ck(self.L.fxg_dcollapse_fetch_local(self.h, local_index, _ptr(out_seq), _ptr(out_len), _ptr(out_count), _ptr(out_first), _ptr(out_hash)))

    def fetch_order(self, perm_owner=None, perm_index=None, ordered_first=None, ordered_count=None):
        self._ck(self.L.fxg_dcollapse_fetch_order(self.h, _ptr(perm_owner), _ptr(perm_index), _ptr(ordered_first), _ptr(ordered_count)))

    def launches(self):
        return int(self.L.fxg_dcollapse_launches(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.L.fxg_dcollapse_free(self.h)
            self.h = None


class TextPipe:
    """fxg_text: FASTQ text -> (K-LINES, K-RECS, K-PACK, op, K-EMIT) -> FASTQ text."""

    def __init__(self, ctx, max_chunk_bytes):
        self.L, self.cap = lib(), max_chunk_bytes
        h = C.c_void_p()
        rc = self.L.fxg_text_new(ctx.h, ctx.device, max_chunk_bytes, C.byref(h))
        if rc != FXG_OK:
            raise FxgError(rc, "fxg_text_new: " + self.L.fxg_strerror(rc).decode())
        self.h = h

    def run(self, op, text, q_offset, a0, a1):
        """text: bytes / numpy uint8 array.  Returns (output bytes, TextReport)."""
        import numpy as np
        src = np.frombuffer(text, np.uint8) if isinstance(text, (bytes, bytearray)) else text
        out = np.empty(self.cap + self.cap // 4 + 64, np.uint8)
        rep = TextReport()
        rc = self.L.fxg_text_run_host(self.h, op, src.ctypes.data, src.size, q_offset, a0, a1, out.ctypes.data, C.byref(rep))
        if rc != FXG_OK:
            raise FxgError(rc, self.L.fxg_text_error(self.h).decode())
        return out[: rep.out_bytes].tobytes(), rep

    def set_deflate(self, on):
        rc = self.L.fxg_text_set_deflate(self.h, 1 if on else 0)
        if rc != FXG_OK:
            raise FxgError(rc, self.L.fxg_text_error(self.h).decode())

    def set_format(self, fasta):
        rc = self.L.fxg_text_set_format(self.h, 1 if fasta else 0)
        if rc != FXG_OK:
            raise FxgError(rc, "fxg_text_set_format")

    def collapse(self, text, q_offset, collapser, first_base=-1):
        """fxg_text_collapse_host: add the chunk's reads to a Collapser.  Returns TextReport."""
        import numpy as np
        src = np.frombuffer(text, np.uint8) if isinstance(text, (bytes, bytearray)) else text
        rep = TextReport()
        rc = self.L.fxg_text_collapse_host(self.h, src.ctypes.data, src.size, q_offset, collapser.h, first_base, C.byref(rep))
        if rc != FXG_OK:
            raise FxgError(rc, self.L.fxg_text_error(self.h).decode())
        return rep

    def stats(self, text, q_offset, hist_dev, max_cycles):
        import numpy as np
        src = np.frombuffer(text, np.uint8) if isinstance(text, (bytes, bytearray)) else text
        rep = TextReport()
        rc = self.L.fxg_text_stats_host(self.h, src.ctypes.data, src.size, q_offset, _ptr(hist_dev), max_cycles, C.byref(rep))
        if rc != FXG_OK:
            raise FxgError(rc, self.L.fxg_text_error(self.h).decode())
        return rep

    def numeric_chunks(self):
        return int(self.L.fxg_text_numeric_chunks(self.h))

    def fasta_chunks(self):
        return int(self.L.fxg_text_fasta_chunks(self.h))

    def clip(self, text, q_offset, opts, show_adapter_only=0, expect_len=0):
        """fxg_text_clip_host: fastx_clipper on a chunk of equal-length reads.  Returns (output bytes, TextReport)."""
        import numpy as np
        src = np.frombuffer(text, np.uint8) if isinstance(text, (bytes, bytearray)) else text
        out = np.empty(self.cap + self.cap // 4 + 64, np.uint8)
        rep = TextReport()
        rc = self.L.fxg_text_clip_host(self.h, src.ctypes.data, src.size, q_offset, C.byref(opts), show_adapter_only, expect_len,
                                       out.ctypes.data, C.byref(rep))
        if rc != FXG_OK:
            raise FxgError(rc, self.L.fxg_text_error(self.h).decode())
        return out[: rep.out_bytes].tobytes(), rep

    def close(self):
        if getattr(self, "h", None):
            self.L.fxg_text_free(self.h)
            self.h = None

    __del__ = close


def collapse_order_dev(device, hash_dev, first_dev, count_dev, n_unique, perm_dev):
    rc = lib().fxg_collapse_order_dev(device, _ptr(hash_dev), _ptr(first_dev), _ptr(count_dev), n_unique, _ptr(perm_dev))
    if rc != FXG_OK:
        raise FxgError(rc, "fxg_collapse_order_dev")


class Context:
    """One fxg_ctx (one GPU).  Methods mirror the C entry points 1:1 and raise FxgError on failure."""

    def __init__(self, device=0):
        self.L = lib()
        h = C.c_void_p()
        rc = self.L.fxg_init(device, C.byref(h))
        if rc != FXG_OK:
            raise FxgError(rc, self.L.fxg_strerror(rc).decode() + " — " + self.L.fxg_last_error(None).decode()
                           + " (a B200/sm_100 GPU is required; there is no CPU fallback)")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.L.fxg_destroy(self.h)
            self.h = None

    __del__ = close

    def _ck(self, rc):
        if rc != FXG_OK:
            raise FxgError(rc, self.L.fxg_last_error(self.h).decode() or self.L.fxg_strerror(rc).decode())

    # ---- plumbing
    def set_stream(self, stream_handle):
        """Adopt a cudaStream_t handle (e.g. torch.cuda.current_stream().cuda_stream; 0 = default stream)."""
        self._ck(self.L.fxg_set_stream(self.h, C.c_void_p(stream_handle)))

    def use_own_stream(self):
        self._ck(self.L.fxg_use_own_stream(self.h))

    def set_tuning(self, tile_reads=0, stages=0, ctas_per_sm=0):
        self._ck(self.L.fxg_set_tuning(self.h, tile_reads, stages, ctas_per_sm))

    def sync(self):
        self._ck(self.L.fxg_sync(self.h))
        return self.report()

    def report(self):
        r = Report()
        self._ck(self.L.fxg_get_report(self.h, C.byref(r)))
        return r

    def report_reset(self):
        self._ck(self.L.fxg_report_reset(self.h))

    def launches(self):
        return int(self.L.fxg_kernel_launches(self.h))

    def device_info(self):
        sm, hbm, mj, mn = C.c_int(), C.c_size_t(), C.c_int(), C.c_int()
        self._ck(self.L.fxg_device_info(self.h, C.byref(sm), C.byref(hbm), C.byref(mj), C.byref(mn)))
        return dict(sm_count=sm.value, hbm_bytes=hbm.value, cc=(mj.value, mn.value))

    @staticmethod
    def batch(seq, qual, n, stride, uniform_len=0, lens=None):
        return Batch(_ptr(seq), _ptr(qual), _ptr(lens), uniform_len, stride, n)

    # ---- ops (device pointers)
    def synth_dev(self, seq, qual, n, length, stride, seed, kind=0, q_offset=33, first_read=0, n_total=None):
        self._ck(self.L.fxg_synth_dev(self.h, _ptr(seq), _ptr(qual), n, first_read, n_total or (first_read + n),
                                      length, stride, seed, kind, q_offset))

    def trim_dev(self, b, q_offset, threshold, min_len, out_len, index_base=0):
        self._ck(self.L.fxg_trim_dev(self.h, C.byref(b), q_offset, threshold, min_len, _ptr(out_len), index_base))

    def filter_dev(self, b, q_offset, min_quality, min_percent, keep, index_base=0):
        self._ck(self.L.fxg_filter_dev(self.h, C.byref(b), q_offset, min_quality, min_percent, _ptr(keep), index_base))

    def revcomp_dev(self, b, q_offset, out_seq, out_qual, index_base=0):
        self._ck(self.L.fxg_revcomp_dev(self.h, C.byref(b), q_offset, _ptr(out_seq), _ptr(out_qual), index_base))

    def stats_accum_dev(self, b, q_offset, hist, max_cycles, weight=None, index_base=0):
        self._ck(self.L.fxg_stats_accum_dev(self.h, C.byref(b), q_offset, _ptr(hist), max_cycles, _ptr(weight), index_base))

    def clip_dev(self, b, widths, q_offset, opts, out_len, out_class=None, out_cut=None, index_base=0):
        self._ck(self.L.fxg_clip_dev(self.h, C.byref(b), _ptr(widths), q_offset, C.byref(opts), _ptr(out_len),
                                     _ptr(out_class), _ptr(out_cut), index_base))

    def validate_dev(self, b, q_offset, index_base=0):
        self._ck(self.L.fxg_validate_dev(self.h, C.byref(b), q_offset, index_base))

    def mask_dev(self, b, q_offset, min_quality, mask_char, out_seq, masked_flag, index_base=0):
        self._ck(self.L.fxg_mask_dev(self.h, C.byref(b), q_offset, min_quality, mask_char, _ptr(out_seq), _ptr(masked_flag), index_base))

    def artifacts_dev(self, b, q_offset, keep, index_base=0):
        self._ck(self.L.fxg_artifacts_dev(self.h, C.byref(b), q_offset, _ptr(keep), index_base))

    def barcode_dev(self, b, table, best):
        self._ck(self.L.fxg_barcode_dev(self.h, C.byref(b), C.byref(table), _ptr(best)))

    def barcode_host(self, b, table, best):
        r = Report()
        self._ck(self.L.fxg_barcode_host(self.h, C.byref(b), C.byref(table), _ptr(best), C.byref(r)))
        return r

    def pipeline_dev(self, b, q_offset, stages, final_len):
        """stages: list of Stage; returns the number of surviving reads (blocking)"""
        arr = (Stage * len(stages))(*stages)
        alive = C.c_int64(0)
        self._ck(self.L.fxg_pipeline_dev(self.h, C.byref(b), q_offset, C.cast(arr, C.c_void_p), len(stages), _ptr(final_len), C.byref(alive)))
        return alive.value

    def has_n_dev(self, b, q_offset, has_n, index_base=0):
        self._ck(self.L.fxg_has_n_dev(self.h, C.byref(b), q_offset, _ptr(has_n), index_base))

    def has_n_host(self, b, q_offset, has_n):
        r = Report()
        self._ck(self.L.fxg_has_n_host(self.h, C.byref(b), q_offset, _ptr(has_n), C.byref(r)))
        return r

    def mask_host(self, b, q_offset, min_quality, mask_char, out_seq, masked_flag):
        r = Report()
        self._ck(self.L.fxg_mask_host(self.h, C.byref(b), q_offset, min_quality, mask_char, _ptr(out_seq), _ptr(masked_flag), C.byref(r)))
        return r

    def artifacts_host(self, b, q_offset, keep):
        r = Report()
        self._ck(self.L.fxg_artifacts_host(self.h, C.byref(b), q_offset, _ptr(keep), C.byref(r)))
        return r

    def validate_host(self, b, q_offset):
        r = Report()
        self._ck(self.L.fxg_validate_host(self.h, C.byref(b), q_offset, C.byref(r)))
        return r

    def hash_dev(self, b, hash_out):
        self._ck(self.L.fxg_hash_dev(self.h, C.byref(b), _ptr(hash_out)))

    # ---- ops (host pointers; copies are inside the call)
    def stats_accum_host(self, b, q_offset, hist_dev, max_cycles, weight=None):
        r = Report()
        self._ck(self.L.fxg_stats_accum_host(self.h, C.byref(b), q_offset, _ptr(hist_dev), max_cycles, _ptr(weight), C.byref(r)))
        return r

    def clip_host(self, b, widths, q_offset, opts, out_len, out_class=None):
        r = Report()
        self._ck(self.L.fxg_clip_host(self.h, C.byref(b), _ptr(widths), q_offset, C.byref(opts), _ptr(out_len),
                                      _ptr(out_class), C.byref(r)))
        return r

    def trim_host(self, b, q_offset, threshold, min_len, out_len):
        r = Report()
        self._ck(self.L.fxg_trim_host(self.h, C.byref(b), q_offset, threshold, min_len, _ptr(out_len), C.byref(r)))
        return r

    def filter_host(self, b, q_offset, min_quality, min_percent, keep):
        r = Report()
        self._ck(self.L.fxg_filter_host(self.h, C.byref(b), q_offset, min_quality, min_percent, _ptr(keep), C.byref(r)))
        return r

    def revcomp_host(self, b, q_offset, out_seq, out_qual):
        r = Report()
        self._ck(self.L.fxg_revcomp_host(self.h, C.byref(b), q_offset, _ptr(out_seq), _ptr(out_qual), C.byref(r)))
        return r
