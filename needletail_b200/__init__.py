"""needletail_b200 — host-side mirror of needletail's public surface on top of libntgpu's C ABI.

The product is ``libntgpu.so`` (include/ntgpu.h).  This module is the thin Python face of it, with
the function names of the reference's own Python module (src/python.rs:429-438:
``parse_fastx_file``, ``parse_fastx_string``, ``normalize_seq``, ``reverse_complement``,
``decode_phred`` — the last one is a byte subtraction and stays on the host, as in the reference)
plus batch forms of the ``Sequence`` trait methods
(src/sequence.rs:156-253) and the fused hot-path call.  There is no CPU fallback: importing works
anywhere, but every call needs a B200 (``Context()`` raises ``NtgError`` otherwise).
"""
import ctypes as C
import os, time
import zlib

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get("NTGPU_SO", os.path.join(_HERE, "libntgpu.so"))   # NTGPU_SO: experiment builds only

# ntg_status (include/ntgpu.h) — 1..7 == needletail::errors::ParseErrorKind (src/errors.rs:28-43)
OK = 0
ERROR_KINDS = {1: "Io", 2: "UnknownFormat", 3: "InvalidStart", 4: "InvalidSeparator", 5: "UnequalLengths",
               6: "UnexpectedEnd", 7: "EmptyFile", 16: "InvalidArgument", 17: "Cuda", 18: "Nccl", 19: "NoMemory",
               20: "Unsupported"}
FORMATS = {0: None, 1: "fasta", 2: "fastq"}
LINE_ENDINGS = {0: None, 1: "unix", 2: "windows"}
TALLY_FIELDS = ["n_records", "n_bases", "n_kmers", "n_not_rc", "kmer_sum_lo", "kmer_sum_hi", "n_query",
                "n_minimizers", "minimizer_sum"]


class NtgError(Exception):
    """Library failure (CUDA, NCCL, bad argument)."""

    def __init__(self, status, msg=""):
        self.status = status
        self.kind = ERROR_KINDS.get(status, str(status))
        super().__init__(f"{self.kind}: {msg}")


class NeedletailError(Exception):
    """Parse error — mirrors needletail.NeedletailError (src/python.rs:53-60, src/errors.rs:46-56)."""

    def __init__(self, kind, line=0, id=None, format=None):
        self.kind, self.line, self.id, self.format = kind, line, id, format
        super().__init__(f"{kind} at line {line}" + (f" (record {id!r})" if id else ""))


class _Record(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("start", "id_b", "id_e", "seq_b", "seq_e", "qual_b", "qual_e", "all_e",
                                          "num_bases", "line")]


class _ParseError(C.Structure):
    _fields_ = [("kind", C.c_int32), ("format", C.c_int32), ("line", C.c_uint64), ("record_index", C.c_uint64),
                ("has_id", C.c_int32), ("id", C.c_char * 236)]


class _Records(C.Structure):
    _fields_ = [("format", C.c_int32), ("line_ending", C.c_int32), ("n_records", C.c_uint64),
                ("records", C.POINTER(_Record)), ("error", _ParseError), ("final_line", C.c_uint64),
                ("final_byte", C.c_uint64), ("_priv", C.c_void_p)]


class _Items(C.Structure):
    _fields_ = [("n_seqs", C.c_uint64), ("n_items", C.c_uint64), ("item_offs", C.POINTER(C.c_uint64)),
                ("pos", C.POINTER(C.c_uint32)), ("was_rc", C.POINTER(C.c_uint8)), ("val_lo", C.POINTER(C.c_uint64)),
                ("val_hi", C.POINTER(C.c_uint64)), ("_priv", C.c_void_p)]


class _TallyConfig(C.Structure):
    _fields_ = [("k", C.c_uint32), ("m", C.c_uint32), ("allow_iupac", C.c_uint32), ("has_query", C.c_uint32),
                ("query", C.c_uint8 * 64), ("flags", C.c_uint32), ("qmask_score", C.c_uint32)]


class _Tallies(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in TALLY_FIELDS] + [("reserved", C.c_uint64 * 7)]


_lib = None


def load_library():
    """dlopen libntgpu.so (built in-tree by needletail_b200/build.py) and declare the ABI."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise NtgError(17, f"{_SO} is missing: run `python needletail_b200/build.py` (nvcc, sm_100a)")
    L = C.CDLL(_SO)
    vp, u64, u32, sz, cp = C.c_void_p, C.c_uint64, C.c_uint32, C.c_size_t, C.c_char_p
    P = C.POINTER
    sig = {
        "ntg_abi_version": ([], C.c_int),
        "ntg_device_count": ([P(C.c_int)], C.c_int),
        "ntg_create": ([C.c_int, P(vp)], C.c_int),
        "ntg_destroy": ([vp], None),
        "ntg_last_error": ([vp], cp),
        "ntg_device_info": ([vp, P(C.c_int), P(sz), P(C.c_int), P(C.c_int)], C.c_int),
        "ntg_sync": ([vp], C.c_int),
        "ntg_launch_count": ([vp], u64),
        "ntg_release_scratch": ([vp], C.c_int),
        "ntg_alloc_pinned": ([sz, P(vp)], C.c_int),
        "ntg_free_pinned": ([vp], C.c_int),
        "ntg_device_alloc": ([vp, sz, P(u64)], C.c_int),
        "ntg_device_free": ([vp, u64], C.c_int),
        "ntg_memcpy_h2d": ([vp, u64, vp, sz], C.c_int),
        "ntg_memcpy_d2h": ([vp, vp, u64, sz], C.c_int),
        "ntg_event_record": ([vp, C.c_int], C.c_int),
        "ntg_event_elapsed_ms": ([vp, C.c_int, C.c_int, P(C.c_float)], C.c_int),
        "ntg_parse_fastx": ([vp, vp, sz, P(P(_Records))], C.c_int),
        "ntg_records_free": ([P(_Records)], None),
        "ntg_write_records": ([vp, vp, sz, C.c_int, vp, sz, vp, C.c_int, vp, sz, P(sz)], C.c_int),
        "ntg_parse_fastx_chunk": ([vp, vp, sz, C.c_int, C.c_int, P(P(_Records)), P(u64)], C.c_int),
        "ntg_stream_open": ([vp, P(_TallyConfig), P(vp)], C.c_int),
        "ntg_stream_feed": ([vp, vp, sz], C.c_int),
        "ntg_stream_acquire": ([vp, P(vp), P(sz)], C.c_int),
        "ntg_stream_commit": ([vp, sz], C.c_int),
        "ntg_stream_feed_gz": ([vp, vp, sz, C.c_int], C.c_int),
        "ntg_stream_finish": ([vp, P(_Tallies), P(_ParseError)], C.c_int),
        "ntg_inflate_bgzf": ([vp, vp, sz, vp, sz, P(sz)], C.c_int),
        "ntg_spectrum_create": ([vp, u32, u64, P(vp)], C.c_int),
        "ntg_spectrum_destroy": ([vp], None),
        "ntg_spectrum_clear": ([vp], C.c_int),
        "ntg_spectrum_add_fastx": ([vp, vp, sz, P(_Tallies), P(_ParseError)], C.c_int),
        "ntg_spectrum_add_fastx_device": ([vp, u64, sz, P(_Tallies), P(_ParseError)], C.c_int),
        "ntg_spectrum_count": ([vp, cp, P(u64)], C.c_int),
        "ntg_spectrum_export": ([vp, vp, vp, u64, P(u64)], C.c_int),
        "ntg_spectrum_histogram": ([vp, vp, u32], C.c_int),
        "ntg_spectrum_reduce": ([vp], C.c_int),
        "ntg_spectrum_kmers": ([vp], u64),
        "ntg_stream_bytes": ([vp], u64),
        "ntg_stream_close": ([vp], None),
        "ntg_tally_fastx_file": ([vp, cp, P(_TallyConfig), C.c_int, P(_Tallies), P(_ParseError)], C.c_int),
        "ntg_normalize": ([vp, vp, vp, sz, C.c_int, vp, vp, vp], C.c_int),
        "ntg_strip_returns": ([vp, vp, vp, sz, vp, vp, vp], C.c_int),
        "ntg_reverse_complement": ([vp, vp, vp, sz, vp], C.c_int),
        "ntg_quality_mask": ([vp, vp, vp, vp, vp, sz, C.c_uint8, vp], C.c_int),
        "ntg_kmers": ([vp, vp, vp, sz, u32, P(P(_Items))], C.c_int),
        "ntg_items_free": ([P(_Items)], None),
        "ntg_canonical_kmers": ([vp, vp, vp, vp, sz, u32, P(P(_Items))], C.c_int),
        "ntg_bit_kmers": ([vp, vp, vp, sz, u32, C.c_int, P(P(_Items))], C.c_int),
        "ntg_bit_minimizers": ([vp, vp, vp, sz, u32, u32, P(P(_Items))], C.c_int),
        "ntg_bitkmer_reverse_complement": ([vp, vp, sz, u32, vp], C.c_int),
        "ntg_bitkmer_canonical": ([vp, vp, sz, u32, vp, vp], C.c_int),
        "ntg_bitkmer_minimizer": ([vp, vp, sz, u32, u32, vp], C.c_int),
        "ntg_tally_fastx": ([vp, vp, sz, P(_TallyConfig), P(_Tallies), P(_ParseError)], C.c_int),
        "ntg_tally_fastx_device": ([vp, u64, sz, P(_TallyConfig), P(_Tallies), P(_ParseError)], C.c_int),
        "ntg_tally_fastx_device_enqueue": ([vp, u64, sz, P(_TallyConfig)], C.c_int),
        "ntg_tally_fastx_device_collect": ([vp, P(_Tallies), P(_ParseError), P(C.c_float)], C.c_int),
        "ntg_synth_fastq_device": ([vp, u64, u64, u64, u64, u32, u32], C.c_int),
        "ntg_synth_fasta_device": ([vp, u64, u64, u64, u64, u32, u32], C.c_int),
        "ntg_comm_unique_id": ([vp], C.c_int),
        "ntg_comm_init": ([vp, C.c_int, C.c_int, vp], C.c_int),
        "ntg_comm_allreduce_tallies": ([vp, P(_Tallies)], C.c_int),
        "ntg_comm_destroy": ([vp], C.c_int),
    }
    for name, (args, res) in sig.items():
        f = getattr(L, name)      # raises AttributeError if the .so does not export what the header declares
        f.argtypes, f.restype = args, res
    L._declared = sorted(sig)
    _lib = L
    return L


def _as_u8(data):
    if isinstance(data, np.ndarray):
        assert data.dtype == np.uint8
        return np.ascontiguousarray(data)
    return np.frombuffer(bytes(data), dtype=np.uint8)


def _batch(seqs):
    """list of bytes -> (concatenated uint8 array, uint64 offsets)"""
    offs = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if seqs:
        offs[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    cat = np.frombuffer(b"".join(bytes(s) for s in seqs), dtype=np.uint8) if seqs else np.zeros(0, dtype=np.uint8)
    return cat, offs


def _ptr(a):
    return a.ctypes.data if a.size else None


class Items:
    """CSR result of canonical_kmers / bit_kmers / bit_minimizers (one row per emitted item)."""

    def __init__(self, it):
        n, ni = it.n_seqs, it.n_items
        self.item_offs = np.ctypeslib.as_array(it.item_offs, shape=(n + 1,)).copy()
        self.pos = np.ctypeslib.as_array(it.pos, shape=(ni,)).copy() if ni else np.zeros(0, np.uint32)
        self.was_rc = (np.ctypeslib.as_array(it.was_rc, shape=(ni,)).copy() if ni else np.zeros(0, np.uint8)) if it.was_rc else None
        self.val_lo = (np.ctypeslib.as_array(it.val_lo, shape=(ni,)).copy() if ni else np.zeros(0, np.uint64)) if it.val_lo else None
        self.val_hi = (np.ctypeslib.as_array(it.val_hi, shape=(ni,)).copy() if ni else np.zeros(0, np.uint64)) if it.val_hi else None

    def of(self, i):
        a, b = int(self.item_offs[i]), int(self.item_offs[i + 1])
        return slice(a, b)


class Context:
    """One CUDA device + its streams (ntg_ctx).  Not thread-safe, like a `&mut` reader."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = C.c_void_p()
        st = self.lib.ntg_create(device, C.byref(h))
        if st != OK:
            raise NtgError(st, "ntg_create failed (no CUDA device / not sm_100?) — libntgpu has no CPU path")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.ntg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st):
        if st != OK:
            raise NtgError(st, self.lib.ntg_last_error(self.h).decode(errors="replace"))

    # ---- info / memory / timing
    def device_info(self):
        sm, mem, ma, mi = C.c_int(), C.c_size_t(), C.c_int(), C.c_int()
        self._ck(self.lib.ntg_device_info(self.h, C.byref(sm), C.byref(mem), C.byref(ma), C.byref(mi)))
        return dict(sm_count=sm.value, total_mem=mem.value, cc=(ma.value, mi.value))

    def launch_count(self):
        return int(self.lib.ntg_launch_count(self.h))

    def sync(self):
        self._ck(self.lib.ntg_sync(self.h))

    def device_alloc(self, nbytes):
        d = C.c_uint64()
        self._ck(self.lib.ntg_device_alloc(self.h, nbytes, C.byref(d)))
        return d.value

    def device_free(self, dptr):
        self._ck(self.lib.ntg_device_free(self.h, dptr))

    def h2d(self, dptr, arr):
        arr = _as_u8(arr)
        self._ck(self.lib.ntg_memcpy_h2d(self.h, dptr, arr.ctypes.data, arr.size))

    def d2h(self, dptr, nbytes):
        out = np.empty(nbytes, dtype=np.uint8)
        self._ck(self.lib.ntg_memcpy_d2h(self.h, out.ctypes.data, dptr, nbytes))
        return out

    def event_record(self, slot):
        self._ck(self.lib.ntg_event_record(self.h, slot))

    def event_elapsed_ms(self, a, b):
        ms = C.c_float()
        self._ck(self.lib.ntg_event_elapsed_ms(self.h, a, b, C.byref(ms)))
        return ms.value

    # ---- (1) record scanner
    def parse(self, data):
        """-> Parsed (records before the first error, plus .error) — ntg_parse_fastx"""
        arr = _as_u8(data)
        out = C.POINTER(_Records)()
        self._ck(self.lib.ntg_parse_fastx(self.h, _ptr(arr), arr.size, C.byref(out)))
        try:
            return Parsed(arr, out.contents)
        finally:
            self.lib.ntg_records_free(out)

    def write_records(self, data, parsed, keep=None, line_ending=None):
        """Text of the kept records of a Parsed table (filtered output) — ntg_write_records; line_ending "unix" / "windows"
        (default: the input's, like SequenceRecord::write)."""
        arr = _as_u8(data)
        table = np.ascontiguousarray(parsed.table, dtype=np.uint64)
        le = {"unix": 1, "windows": 2, None: {"windows": 2}.get(parsed.line_ending, 1)}[line_ending]
        fmt = {"fasta": 1, "fastq": 2}[parsed.format]
        kp = None if keep is None else np.ascontiguousarray(keep, dtype=np.uint8)
        assert kp is None or kp.size == len(table)
        n = C.c_size_t()
        args = (self.h, _ptr(arr), arr.size, fmt, table.ctypes.data if len(table) else None, len(table), kp.ctypes.data if kp is not None and kp.size else None, le)
        st = self.lib.ntg_write_records(*args, None, 0, C.byref(n))
        if st not in (OK, 16):
            self._ck(st)
        out = np.empty(max(1, n.value), dtype=np.uint8)
        self._ck(self.lib.ntg_write_records(*args, out.ctypes.data, out.size, C.byref(n)))
        return out[:n.value].tobytes()

    def release_scratch(self):
        """Hand the scanner's / batch calls' cached device scratch and pinned result buffers back — ntg_release_scratch"""
        self._ck(self.lib.ntg_release_scratch(self.h))

    def parse_chunks(self, data, window=1 << 30, with_records=True):
        """Generator of Parsed, one per window of `window` bytes (< 4 GiB) — ntg_parse_fastx_chunk: the incremental reader behind
        FastxReader for inputs of any size.  Offsets, lines and error positions are made absolute; the last Parsed carries
        the end-of-stream error, if any."""
        arr = _as_u8(data)
        n, pos, fmt, line_base, rec_base = arr.size, 0, 0, 0, 0
        while True:
            ln = min(window, n - pos)
            at_eof = pos + ln == n
            out = C.POINTER(_Records)(); consumed = C.c_uint64()
            win = arr[pos:pos + ln]
            t0 = time.perf_counter()
            self._ck(self.lib.ntg_parse_fastx_chunk(self.h, _ptr(win), ln, fmt, int(at_eof), C.byref(out), C.byref(consumed)))
            self.parse_seconds += time.perf_counter() - t0          # time inside the C ABI (this generator adds table copies)
            try:
                p = Parsed(win, out.contents, with_records)
            finally:
                self.lib.ntg_records_free(out)
            fmt = {"fasta": 1, "fastq": 2}.get(p.format, 0)
            if len(p.table) and (pos or line_base):
                for c in (0, 1, 2, 3, 4, 7) + ((5, 6) if p.format == "fastq" else ()):
                    p.table[:, c] += np.uint64(pos)
                p.table[:, 9] += np.uint64(line_base)
                for r in p.records:
                    r.byte += pos; r.line += line_base
            p.err_line += line_base; p.err_record_index += rec_base
            p.final_byte += pos; p.final_line += line_base
            done = at_eof or p.err_kind is not None
            if not done and consumed.value == 0:
                if window >= 0xF0000000:
                    raise NtgError(20, "a record larger than 3.75 GiB")
                window = min(window * 2, 0xF0000000)
                continue
            yield p
            if done:
                return
            line_base = p.final_line - 1
            rec_base += len(p.table)
            pos += consumed.value

    # ---- (2) Sequence trait, batch form
    def normalize(self, seqs, iupac=False):
        """list[bytes] -> (list[bytes], changed flags) — sequence::normalize per sequence"""
        cat, offs = _batch(seqs)
        out = np.empty(max(1, cat.size), dtype=np.uint8)
        ooffs = np.zeros(len(seqs) + 1, dtype=np.uint64)
        ch = np.zeros(max(1, len(seqs)), dtype=np.uint8)
        self._ck(self.lib.ntg_normalize(self.h, _ptr(cat), offs.ctypes.data, len(seqs), int(iupac), out.ctypes.data,
                                        ooffs.ctypes.data, ch.ctypes.data))
        return [out[int(ooffs[i]):int(ooffs[i + 1])].tobytes() for i in range(len(seqs))], [bool(x) for x in ch[:len(seqs)]]

    def strip_returns(self, seqs):
        cat, offs = _batch(seqs)
        out = np.empty(max(1, cat.size), dtype=np.uint8)
        ooffs = np.zeros(len(seqs) + 1, dtype=np.uint64)
        ch = np.zeros(max(1, len(seqs)), dtype=np.uint8)
        self._ck(self.lib.ntg_strip_returns(self.h, _ptr(cat), offs.ctypes.data, len(seqs), out.ctypes.data,
                                            ooffs.ctypes.data, ch.ctypes.data))
        return [out[int(ooffs[i]):int(ooffs[i + 1])].tobytes() for i in range(len(seqs))], [bool(x) for x in ch[:len(seqs)]]

    def reverse_complement(self, seqs):
        cat, offs = _batch(seqs)
        out = np.empty(max(1, cat.size), dtype=np.uint8)
        self._ck(self.lib.ntg_reverse_complement(self.h, _ptr(cat), offs.ctypes.data, len(seqs), out.ctypes.data))
        return [out[int(offs[i]):int(offs[i + 1])].tobytes() for i in range(len(seqs))]

    def quality_mask(self, seqs, quals, score):
        if len(seqs) != len(quals):
            raise ValueError("quality_mask: one quality string per sequence")
        cat, offs = _batch(seqs)
        qcat, qoffs = _batch(quals)
        out = np.empty(max(1, cat.size), dtype=np.uint8)
        self._ck(self.lib.ntg_quality_mask(self.h, _ptr(cat), _ptr(qcat), offs.ctypes.data, qoffs.ctypes.data, len(seqs), score, out.ctypes.data))
        return [out[int(offs[i]):int(offs[i + 1])].tobytes() for i in range(len(seqs))]

    def _items(self, fn, *args):
        out = C.POINTER(_Items)()
        self._ck(fn(self.h, *args, C.byref(out)))
        try:
            return Items(out.contents)
        finally:
            self.lib.ntg_items_free(out)

    def canonical_kmers(self, seqs, k, rcs=None):
        cat, offs = _batch(seqs)
        rc = _batch(rcs)[0] if rcs is not None else None
        return self._items(self.lib.ntg_canonical_kmers, _ptr(cat), _ptr(rc) if rc is not None else None,
                           offs.ctypes.data, len(seqs), k)

    def kmers(self, seqs, k):
        """Sequence::kmers: list (one entry per sequence) of the k-byte windows, as bytes — ntg_kmers gives the positions"""
        cat, offs = _batch(seqs)
        it = self._items(self.lib.ntg_kmers, _ptr(cat), offs.ctypes.data, len(seqs), k)
        return [[bytes(seqs[i][int(p):int(p) + k]) for p in it.pos[it.of(i)]] for i in range(len(seqs))]

    def bit_kmers(self, seqs, k, canonical=False):
        cat, offs = _batch(seqs)
        return self._items(self.lib.ntg_bit_kmers, _ptr(cat), offs.ctypes.data, len(seqs), k, int(canonical))

    def bit_minimizers(self, seqs, k, m):
        cat, offs = _batch(seqs)
        return self._items(self.lib.ntg_bit_minimizers, _ptr(cat), offs.ctypes.data, len(seqs), k, m)

    def bitkmer_reverse_complement(self, vals, k):
        v = np.ascontiguousarray(vals, dtype=np.uint64); out = np.empty_like(v)
        self._ck(self.lib.ntg_bitkmer_reverse_complement(self.h, _ptr(v), v.size, k, out.ctypes.data))
        return out

    def bitkmer_canonical(self, vals, k):
        v = np.ascontiguousarray(vals, dtype=np.uint64); out = np.empty_like(v); fl = np.zeros(max(1, v.size), np.uint8)
        self._ck(self.lib.ntg_bitkmer_canonical(self.h, _ptr(v), v.size, k, out.ctypes.data, fl.ctypes.data))
        return out, fl[:v.size]

    def bitkmer_minimizer(self, vals, k, m):
        v = np.ascontiguousarray(vals, dtype=np.uint64); out = np.empty_like(v)
        self._ck(self.lib.ntg_bitkmer_minimizer(self.h, _ptr(v), v.size, k, m, out.ctypes.data))
        return out

    # ---- (3) fused hot path
    parse_seconds = 0.0      # seconds spent inside ntg_parse_fastx_chunk (parse_chunks)
    tally_flags = 0          # NTG_TALLY_* bits sent with every tally call (1 = no FASTQ line-phase speculation; diagnostic)

    def _cfg(self, k, m, iupac, query, qmask=0):
        cfg = _TallyConfig(k=k, m=m, allow_iupac=int(iupac), has_query=int(query is not None), flags=self.tally_flags, qmask_score=qmask)
        if query is not None:
            q = bytes(query)
            for i, b in enumerate(q[:64]):
                cfg.query[i] = b
        return cfg

    @staticmethod
    def _tally_result(t, e):
        d = {f: int(getattr(t, f)) for f in TALLY_FIELDS}
        d["err_kind"] = ERROR_KINDS.get(e.kind) if e.kind else None
        d["err_line"] = int(e.line)
        d["fallback"] = int(t.reserved[0]) & 0xFFFFFFFF      # 0: the single-pass fused kernel produced the tallies
        d["spec_missed"] = int(t.reserved[1]) & 0xFFFFFFFF   # != 0: a speculated FASTQ line phase was wrong and the pass re-ran without speculation
        d["fast_path"] = bool(int(t.reserved[1]) >> 32)       # the record-owned short-read FASTQ kernel produced the tallies
        # NTG_STATS builds of the tile kernel only (per-CTA cycle accounting, see fused.cuh); zero otherwise
        d["stats"] = {"cta_cycles_sum": int(t.reserved[2]), "p0_or_lookback_cycles": int(t.reserved[3]), "p1_cycles_or_lookbacks": int(t.reserved[4]),
                      "barrier_wait_cycles_t0": int(t.reserved[5]), "walk_cycles_t0": int(t.reserved[6])}
        return d

    def tally(self, data, k, m=0, iupac=False, query=None, qmask=0):
        """Host bytes -> tallies dict (the end-to-end call: H2D inside) — ntg_tally_fastx.  qmask: quality_mask(score) first."""
        arr = _as_u8(data)
        cfg = self._cfg(k, m, iupac, query, qmask); t = _Tallies(); e = _ParseError()
        self._ck(self.lib.ntg_tally_fastx(self.h, _ptr(arr), arr.size, C.byref(cfg), C.byref(t), C.byref(e)))
        return self._tally_result(t, e)

    def tally_ptr(self, host_ptr, nbytes, k, m=0, iupac=False, query=None):
        cfg = self._cfg(k, m, iupac, query); t = _Tallies(); e = _ParseError()
        self._ck(self.lib.ntg_tally_fastx(self.h, host_ptr, nbytes, C.byref(cfg), C.byref(t), C.byref(e)))
        return self._tally_result(t, e)

    def tally_device(self, dptr, nbytes, k, m=0, iupac=False, query=None, qmask=0):
        cfg = self._cfg(k, m, iupac, query, qmask); t = _Tallies(); e = _ParseError()
        self._ck(self.lib.ntg_tally_fastx_device(self.h, dptr, nbytes, C.byref(cfg), C.byref(t), C.byref(e)))
        return self._tally_result(t, e)

    def tally_device_enqueue(self, dptr, nbytes, k, m=0, iupac=False, query=None, allreduce=False):
        """allreduce=True (after comm_init): the tallies of all ranks are summed by one in-stream ncclAllReduce (NTG_TALLY_ALLREDUCE)"""
        cfg = self._cfg(k, m, iupac, query)
        if allreduce:
            cfg.flags |= 2
        self._ck(self.lib.ntg_tally_fastx_device_enqueue(self.h, dptr, nbytes, C.byref(cfg)))

    def tally_file(self, path, k, m=0, iupac=False, query=None, threads=1):
        """parse_fastx_file for the tally path: plain or gzip (BGZF inflates on `threads` workers) — ntg_tally_fastx_file"""
        cfg = self._cfg(k, m, iupac, query); t = _Tallies(); e = _ParseError()
        self._ck(self.lib.ntg_tally_fastx_file(self.h, os.fsencode(path), C.byref(cfg), threads, C.byref(t), C.byref(e)))
        return self._tally_result(t, e)

    def inflate_bgzf(self, blob):
        """BGZF bytes -> text, inflated by the device-side DEFLATE decoder — ntg_inflate_bgzf"""
        arr = _as_u8(blob)
        n = C.c_size_t()
        st = self.lib.ntg_inflate_bgzf(self.h, _ptr(arr), arr.size, None, 0, C.byref(n))
        if st not in (OK, 16):
            self._ck(st)
        out = np.empty(max(1, n.value), dtype=np.uint8)
        self._ck(self.lib.ntg_inflate_bgzf(self.h, _ptr(arr), arr.size, out.ctypes.data, out.size, C.byref(n)))
        return out[:n.value].tobytes()

    def spectrum(self, k, capacity=0):
        return Spectrum(self, k, capacity)

    def stream(self, k, m=0, iupac=False, query=None):
        """A tally session over a stream of unknown length (parse_fastx_reader<R: Read>) — ntg_stream_*"""
        return TallyStream(self, self._cfg(k, m, iupac, query))

    def tally_device_collect(self):
        t = _Tallies(); e = _ParseError(); ms = C.c_float()
        self._ck(self.lib.ntg_tally_fastx_device_collect(self.h, C.byref(t), C.byref(e), C.byref(ms)))
        d = self._tally_result(t, e)
        d["fused_kernel_ms"] = ms.value
        d["not_reduced"] = bool(int(t.reserved[0]) >> 32)
        return d

    # ---- (4) synthetic inputs
    def synth_fastq_device(self, dptr, seed, rec0, nrec, read_len, n_thresh=0):
        self._ck(self.lib.ntg_synth_fastq_device(self.h, dptr, seed, rec0, nrec, read_len, n_thresh))

    def synth_fasta_device(self, dptr, seed, rec0, nrec, read_len, n_thresh=0):
        self._ck(self.lib.ntg_synth_fasta_device(self.h, dptr, seed, rec0, nrec, read_len, n_thresh))

    # ---- (5) multi-GPU
    def comm_unique_id(self):
        buf = (C.c_uint8 * 128)()
        st = self.lib.ntg_comm_unique_id(buf)
        if st != OK:
            raise NtgError(st, "ncclGetUniqueId")
        return bytes(buf)

    def comm_init(self, n_ranks, rank, uid):
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        self._ck(self.lib.ntg_comm_init(self.h, n_ranks, rank, buf))

    def comm_allreduce_tallies(self, d):
        t = _Tallies()
        for f in TALLY_FIELDS:
            setattr(t, f, d[f])
        self._ck(self.lib.ntg_comm_allreduce_tallies(self.h, C.byref(t)))
        return {f: int(getattr(t, f)) for f in TALLY_FIELDS}


class Spectrum:
    """ntg_spectrum: counts per distinct canonical k-mer (dense histogram for k <= 14, hash table of `capacity` slots above)."""

    def __init__(self, ctx, k, capacity=0):
        self.ctx, self.k, self.h = ctx, k, C.c_void_p()
        ctx._ck(ctx.lib.ntg_spectrum_create(ctx.h, k, capacity, C.byref(self.h)))

    def add(self, data):
        arr = _as_u8(data); t = _Tallies(); e = _ParseError()
        self.ctx._ck(self.ctx.lib.ntg_spectrum_add_fastx(self.h, _ptr(arr), arr.size, C.byref(t), C.byref(e)))
        return self.ctx._tally_result(t, e)

    def add_device(self, dptr, nbytes):
        t = _Tallies(); e = _ParseError()
        self.ctx._ck(self.ctx.lib.ntg_spectrum_add_fastx_device(self.h, dptr, nbytes, C.byref(t), C.byref(e)))
        return self.ctx._tally_result(t, e)

    def count(self, kmer):
        v = C.c_uint64()
        self.ctx._ck(self.ctx.lib.ntg_spectrum_count(self.h, bytes(kmer), C.byref(v)))
        return int(v.value)

    def items(self):
        """-> (keys uint64, counts uint32) of the distinct k-mers, sorted by key"""
        n = C.c_uint64()
        self.ctx._ck(self.ctx.lib.ntg_spectrum_export(self.h, None, None, 0, C.byref(n)))
        keys = np.empty(max(1, n.value), np.uint64); counts = np.empty(max(1, n.value), np.uint32)
        self.ctx._ck(self.ctx.lib.ntg_spectrum_export(self.h, keys.ctypes.data, counts.ctypes.data, n.value, C.byref(n)))
        keys, counts = keys[:n.value], counts[:n.value]
        o = np.argsort(keys, kind="stable")
        return keys[o], counts[o]

    def histogram(self, n_bins=256):
        h = np.zeros(n_bins, np.uint64)
        self.ctx._ck(self.ctx.lib.ntg_spectrum_histogram(self.h, h.ctypes.data, n_bins))
        return h

    def reduce(self):
        self.ctx._ck(self.ctx.lib.ntg_spectrum_reduce(self.h))

    def kmers_added(self):
        return int(self.ctx.lib.ntg_spectrum_kmers(self.h))

    def clear(self):
        self.ctx._ck(self.ctx.lib.ntg_spectrum_clear(self.h))

    def close(self):
        if self.h:
            self.ctx.lib.ntg_spectrum_destroy(self.h)
            self.h = C.c_void_p()


class TallyStream:
    """ntg_stream: feed(bytes) / feed_gz(bytes, threads) any number of times, then finish() -> tallies dict."""

    def __init__(self, ctx, cfg):
        self.ctx, self.h = ctx, C.c_void_p()
        ctx._ck(ctx.lib.ntg_stream_open(ctx.h, C.byref(cfg), C.byref(self.h)))

    def feed(self, data):
        arr = _as_u8(data)
        self.ctx._ck(self.ctx.lib.ntg_stream_feed(self.h, _ptr(arr), arr.size))

    def feed_ptr(self, host_ptr, nbytes):
        self.ctx._ck(self.ctx.lib.ntg_stream_feed(self.h, host_ptr, nbytes))

    def feed_gz(self, data, threads=1):
        arr = _as_u8(data)
        self.ctx._ck(self.ctx.lib.ntg_stream_feed_gz(self.h, _ptr(arr), arr.size, threads))

    def feed_gz_ptr(self, host_ptr, nbytes, threads=1):
        self.ctx._ck(self.ctx.lib.ntg_stream_feed_gz(self.h, host_ptr, nbytes, threads))

    def bytes_fed(self):
        return int(self.ctx.lib.ntg_stream_bytes(self.h))

    def finish(self):
        t = _Tallies(); e = _ParseError()
        try:
            self.ctx._ck(self.ctx.lib.ntg_stream_finish(self.h, C.byref(t), C.byref(e)))
        finally:
            self.close()
        return self.ctx._tally_result(t, e)

    def close(self):
        if self.h:
            self.ctx.lib.ntg_stream_close(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def _snippet(seq, max_len=20):
    # python.rs:37-45 get_seq_snippet
    return seq[:max_len - 4] + "\u2026" + seq[-3:] if len(seq) > max_len else seq


class Record:
    """needletail's Python Record (src/python.rs:88-287): ``Record(id, seq, qual=None)`` with ``id``, ``seq``, ``qual``
    as str, ``name`` / ``description``, ``is_fasta()`` / ``is_fastq()``, ``normalize(iupac=False)``, ``==``, ``hash``,
    ``len``, ``str`` (the record as FASTA / FASTQ text) and ``repr``.  Records yielded by a reader additionally carry
    the reference's Rust-side accessors: ``raw_seq`` (bytes, line breaks included), ``all``, ``num_bases`` and the
    record's ``line`` / ``byte`` position (src/parser/record.rs:66-154)."""

    __slots__ = ("id", "seq", "qual", "raw_seq", "all", "num_bases", "line", "byte")

    def __init__(self, id, seq, qual=None):
        if qual is not None and len(qual) != len(seq):
            raise ValueError("Sequence and quality strings must have the same length")     # python.rs:207-214
        self.id, self.seq, self.qual = id, seq, qual
        self.raw_seq = seq.encode()
        self.all, self.num_bases, self.line, self.byte = None, len(seq), None, None

    @classmethod
    def _from_table(cls, data, r, fmt):
        b = data
        self = cls.__new__(cls)
        self.id = b[r.id_b:r.id_e].tobytes().decode("utf-8", errors="replace")
        self.raw_seq = b[r.seq_b:r.seq_e].tobytes()
        # Record.seq is SequenceRecord::seq(): raw_seq minus all \r\n (src/parser/record.rs:84-89, python.rs:136-142)
        self.seq = self.raw_seq.replace(b"\n", b"").replace(b"\r", b"").decode("utf-8", errors="replace")
        self.qual = b[r.qual_b:r.qual_e].tobytes().decode("utf-8", errors="replace") if fmt == "fastq" else None
        self.all = b[r.start:r.all_e].tobytes()
        self.num_bases = int(r.num_bases)
        self.line = int(r.line)
        self.byte = int(r.start)
        return self

    def _first_ws(self):
        for i, c in enumerate(self.id):
            if c.isspace():
                return i
        return -1

    @property
    def name(self):                                   # python.rs:148-154: the id up to its first whitespace character
        i = self._first_ws()
        return self.id if i < 0 else self.id[:i]

    @property
    def description(self):                            # python.rs:157-163: what follows it, left-trimmed; None without whitespace
        i = self._first_ws()
        return None if i < 0 else self.id[i:].lstrip()

    def is_fasta(self):
        return self.qual is None

    def is_fastq(self):
        return self.qual is not None

    def normalize(self, iupac=False, ctx=None):
        """In place, like python.rs:197-202 (the normalisation itself is ntg_normalize on the device)."""
        self.seq = normalize_seq(self.seq, iupac, ctx)

    def __eq__(self, other):
        return isinstance(other, Record) and (self.id, self.seq, self.qual) == (other.id, other.seq, other.qual)

    def __hash__(self):
        return hash((self.id, self.seq)) if self.qual is None else hash((self.id, self.seq, self.qual))

    def __len__(self):
        return len(self.seq)

    def __str__(self):
        if self.qual is None:
            return ">%s\n%s\n" % (self.id, self.seq)
        return "@%s\n%s\n+\n%s\n" % (self.id, self.seq, self.qual)

    def __repr__(self):
        name = self.name
        id_snippet = name + "\u2026" if name != self.id else name
        return "Record(id=%s, seq=%s, qual=%s)" % (id_snippet, _snippet(self.seq), "None" if self.qual is None else _snippet(self.qual))


class Parsed:
    def __init__(self, data, rs, with_records=True):
        self.format = FORMATS[rs.format]
        self.line_ending = LINE_ENDINGS[rs.line_ending]
        self.final_line, self.final_byte = int(rs.final_line), int(rs.final_byte)
        n = int(rs.n_records)
        self.table = np.ctypeslib.as_array(C.cast(rs.records, C.POINTER(C.c_uint64)), shape=(n, 10)).copy() if n else np.zeros((0, 10), np.uint64)
        self.records = [Record._from_table(data, rs.records[i], self.format) for i in range(n)] if with_records else []
        e = rs.error
        self.err_kind = ERROR_KINDS.get(e.kind) if e.kind else None
        self.err_line = int(e.line)
        self.err_record_index = int(e.record_index)
        self.err_id = e.id if e.has_id else None

    def error(self):
        return NeedletailError(self.err_kind, self.err_line, self.err_id, self.format) if self.err_kind else None


# ------------------------------------------------------------------------------------------------
# module-level API with the reference's Python names (src/python.rs:291-427)
_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    return _default_ctx


def _decompress(raw):
    """Host side of the boundary: compression sniff (src/parser/mod.rs:27-35,95-147).  gzip is
    multi-member like flate2::MultiGzDecoder; bz2/xz via the stdlib; zstd is not available here."""
    if len(raw) >= 2:
        magic = raw[:2]
        if magic == b"\x1f\x8b":
            out, d = [], raw
            while d:
                z = zlib.decompressobj(16 + zlib.MAX_WBITS)
                out.append(z.decompress(d)); out.append(z.flush())
                d = z.unused_data
            return b"".join(out)
        if magic == b"BZ":
            import bz2
            return bz2.decompress(raw)
        if magic == b"\xfd7":
            import lzma
            return lzma.decompress(raw)
        if magic == b"\x28\xb5":
            raise NtgError(20, "zstd input: no decoder in this environment")
    return raw


class FastxReader:
    """Iterator over records (python.rs:62-86); raises NeedletailError at the first invalid record,
    after yielding the valid ones before it — the reference's iteration order.  The input is scanned window by window
    (Context.parse_chunks), so its size is not limited by the 32-bit offsets of one scanner call."""

    WINDOW = 1 << 30

    def __init__(self, data, ctx=None):
        ctx = ctx or default_context()
        raw = _decompress(bytes(data))
        if len(data) >= 2 and len(raw) < 1 and raw != data:
            raise NeedletailError("EmptyFile")
        self._chunks = ctx.parse_chunks(raw, self.WINDOW)
        self._p = next(self._chunks)
        if self._p.err_kind in ("EmptyFile", "UnknownFormat") and not self._p.records:
            raise self._p.error()          # parse_fastx_reader fails up front (mod.rs:88-91,44)
        self._i = 0

    def __iter__(self):
        return self

    def __next__(self):
        while True:
            if self._i < len(self._p.records):
                r = self._p.records[self._i]
                self._i += 1
                return r
            if self._p.err_kind and self._i == len(self._p.records):
                self._i += 1
                raise self._p.error()
            nxt = next(self._chunks, None) if not self._p.err_kind else None
            if nxt is None:
                raise StopIteration
            self._p, self._i = nxt, 0


def parse_fastx_file(path, ctx=None):
    try:
        with open(os.fspath(path), "rb") as f:
            data = f.read()
    except OSError as e:
        raise NeedletailError("Io") from e
    return FastxReader(data, ctx)


def parse_fastx_string(content, ctx=None):
    return FastxReader(content.encode() if isinstance(content, str) else content, ctx)


def normalize_seq(seq, iupac=False, ctx=None):
    out, _ = (ctx or default_context()).normalize([seq.encode() if isinstance(seq, str) else seq], iupac)
    return out[0].decode()


def reverse_complement(seq, ctx=None):
    out = (ctx or default_context()).reverse_complement([seq.encode() if isinstance(seq, str) else seq])
    return out[0].decode()


def write_fasta(id, seq, line_ending="unix"):
    """record::write_fasta (src/parser/record.rs:207-220) as bytes: host formatting of ONE record (batches: Context.write_records)"""
    e = b"\r\n" if line_ending == "windows" else b"\n"
    return b">" + bytes(id) + e + bytes(seq) + e


def write_fastq(id, seq, qual=None, line_ending="unix"):
    """record::write_fastq (src/parser/record.rs:222-247): a missing quality is written as 'I' per base"""
    e = b"\r\n" if line_ending == "windows" else b"\n"
    return b"@" + bytes(id) + e + bytes(seq) + e + b"+" + e + (bytes(qual) if qual is not None else b"I" * len(seq)) + e


def decode_phred(qual, base_64=False):
    """quality::decode_phred (src/quality.rs:15-28) with the signature and error of the reference's Python module
    (src/python.rs:416-427): Phred+33 (or +64) characters -> tuple of scores; a character below the offset raises
    ValueError.  Host arithmetic — off the device path, like the reference's."""
    q = np.frombuffer(qual.encode() if isinstance(qual, str) else bytes(qual), dtype=np.uint8)
    offset = 64 if base_64 else 33
    bad = np.flatnonzero(q < offset)
    if bad.size:
        raise ValueError("Invalid Phred quality: character '%s' cannot be decoded with offset '%d'" % (chr(int(q[bad[0]])), offset))
    return tuple(int(v) for v in (q - offset))
