"""ctypes face of the CPU oracle (oracle/libntoracle.so).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "oracle", "libntoracle.so")

ERR = {0: None, 1: "Io", 2: "UnknownFormat", 3: "InvalidStart", 4: "InvalidSeparator",
       5: "UnequalLengths", 6: "UnexpectedEnd", 7: "EmptyFile"}
FMT = {0: None, 1: "fasta", 2: "fastq"}
LE = {0: None, 1: "unix", 2: "windows"}
TALLY_FIELDS = ["n_records", "n_bases", "n_kmers", "n_not_rc", "kmer_sum_lo", "kmer_sum_hi",
                "n_query", "n_minimizers", "minimizer_sum"]

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
        L = C.CDLL(_SO)
        u8p, u64p, sz = C.c_char_p, C.POINTER(C.c_uint64), C.c_size_t
        L.ntref_normalize.argtypes = [u8p, sz, C.c_int, u8p, C.POINTER(sz)]; L.ntref_normalize.restype = C.c_int
        L.ntref_complement.argtypes = [C.c_uint8]; L.ntref_complement.restype = C.c_uint8
        L.ntref_reverse_complement.argtypes = [u8p, sz, u8p]
        L.ntref_strip_returns.argtypes = [u8p, sz, u8p, C.POINTER(sz)]; L.ntref_strip_returns.restype = C.c_int
        L.ntref_str_canonical.argtypes = [u8p, sz, u8p]
        L.ntref_str_minimizer.argtypes = [u8p, sz, sz, u8p]
        L.ntref_quality_mask.argtypes = [u8p, u8p, sz, C.c_uint8, u8p]
        L.ntref_decode_phred.argtypes = [u8p, sz, C.c_int, u8p]; L.ntref_decode_phred.restype = C.c_int
        L.ntref_canonical_kmers.argtypes = [u8p, sz, u8p, sz, C.c_uint, C.c_void_p, C.c_void_p, sz]
        L.ntref_canonical_kmers.restype = sz
        L.ntref_bit_kmers.argtypes = [u8p, sz, C.c_uint, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, sz]
        L.ntref_bit_kmers.restype = sz
        L.ntref_bit_reverse_complement.argtypes = [C.c_uint64, C.c_uint]; L.ntref_bit_reverse_complement.restype = C.c_uint64
        L.ntref_bit_canonical.argtypes = [C.c_uint64, C.c_uint, C.POINTER(C.c_int)]; L.ntref_bit_canonical.restype = C.c_uint64
        L.ntref_bit_minimizer.argtypes = [C.c_uint64, C.c_uint, C.c_uint]; L.ntref_bit_minimizer.restype = C.c_uint64
        L.ntref_bitmer_to_bytes.argtypes = [C.c_uint64, C.c_uint, u8p]
        L.ntref_bytes_to_bitmer.argtypes = [u8p, sz]; L.ntref_bytes_to_bitmer.restype = C.c_uint64
        L.ntref_parse_fastx.argtypes = [C.c_void_p, sz, C.c_void_p, sz, C.c_void_p, C.c_char_p, sz]
        L.ntref_parse_fastx.restype = sz
        L.ntref_parse_fastx_incremental.argtypes = [C.c_void_p, sz, sz, sz, C.c_void_p, sz, C.c_void_p, C.c_char_p, sz]
        L.ntref_parse_fastx_incremental.restype = sz
        L.ntref_tally_fastx.argtypes = [C.c_void_p, sz, C.c_uint, C.c_uint, C.c_int, u8p, C.c_void_p]
        L.ntref_tally_fastx.restype = C.c_int
        L.ntref_bench_fastq.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_uint, C.c_uint, C.c_int, C.c_void_p]
        L.ntref_bench_fastq.restype = C.c_double
        L.ntref_gen_fastq.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, sz, C.c_uint32, C.c_uint]
        L.ntref_gen_fasta.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, sz, C.c_uint32, C.c_uint]
        _lib = L
    return _lib


def _buf(n):
    return C.create_string_buffer(max(1, n))


def normalize(seq: bytes, iupac: bool):
    """-> (normalized bytes, changed)"""
    out = _buf(len(seq)); n = C.c_size_t(0)
    ch = lib().ntref_normalize(seq, len(seq), int(iupac), out, C.byref(n))
    return out.raw[:n.value], bool(ch)


def complement(c: int) -> int:
    return lib().ntref_complement(c)


def reverse_complement(seq: bytes) -> bytes:
    out = _buf(len(seq)); lib().ntref_reverse_complement(seq, len(seq), out); return out.raw[:len(seq)]


def strip_returns(seq: bytes):
    out = _buf(len(seq)); n = C.c_size_t(0)
    ch = lib().ntref_strip_returns(seq, len(seq), out, C.byref(n))
    return out.raw[:n.value], bool(ch)


def str_canonical(seq: bytes) -> bytes:
    out = _buf(len(seq)); lib().ntref_str_canonical(seq, len(seq), out); return out.raw[:len(seq)]


def str_minimizer(seq: bytes, length: int) -> bytes:
    out = _buf(length); lib().ntref_str_minimizer(seq, len(seq), length, out); return out.raw[:length]


def quality_mask(seq: bytes, qual: bytes, score: int) -> bytes:
    out = _buf(len(seq)); lib().ntref_quality_mask(seq, qual, len(seq), score, out); return out.raw[:len(seq)]


def decode_phred(q: bytes, base64=False):
    out = _buf(len(q))
    ok = lib().ntref_decode_phred(q, len(q), int(base64), out)
    return tuple(out.raw[:len(q)]) if ok else None


def canonical_kmers(seq: bytes, k: int, rc: bytes = None):
    """-> list of (pos, kmer bytes, was_rc) exactly like CanonicalKmers (src/kmer.rs:114-129)"""
    if rc is None:
        rc = reverse_complement(seq)
    cap = max(1, len(seq))
    pos = np.zeros(cap, dtype=np.uint64); fl = np.zeros(cap, dtype=np.uint8)
    c = lib().ntref_canonical_kmers(seq, len(seq), rc, len(rc), k, pos.ctypes.data, fl.ctypes.data, cap)
    out = []
    for i in range(c):
        p = int(pos[i])
        if fl[i]:
            out.append((p, rc[len(rc) - p - k: len(rc) - p], True))
        else:
            out.append((p, seq[p:p + k], False))
    return out


def canonical_kmers_arrays(seq: bytes, k: int, rc: bytes = None):
    if rc is None:
        rc = reverse_complement(seq)
    cap = max(1, len(seq))
    pos = np.zeros(cap, dtype=np.uint64); fl = np.zeros(cap, dtype=np.uint8)
    c = lib().ntref_canonical_kmers(seq, len(seq), rc, len(rc), k, pos.ctypes.data, fl.ctypes.data, cap)
    return pos[:c], fl[:c]


def bit_kmers(seq: bytes, k: int, canonical: bool):
    """-> (pos u64[], kmer u64[], was_rc u8[])"""
    cap = max(1, len(seq))
    pos = np.zeros(cap, dtype=np.uint64); km = np.zeros(cap, dtype=np.uint64); fl = np.zeros(cap, dtype=np.uint8)
    c = lib().ntref_bit_kmers(seq, len(seq), k, int(canonical), pos.ctypes.data, km.ctypes.data, fl.ctypes.data, cap)
    return pos[:c], km[:c], fl[:c]


def bit_reverse_complement(v, k):
    return lib().ntref_bit_reverse_complement(v, k)


def bit_canonical(v, k):
    f = C.c_int(0); r = lib().ntref_bit_canonical(v, k, C.byref(f)); return r, bool(f.value)


def bit_minimizer(v, k, m):
    return lib().ntref_bit_minimizer(v, k, m)


def bitmer_to_bytes(v, k):
    out = _buf(k); lib().ntref_bitmer_to_bytes(v, k, out); return out.raw[:k]


def bytes_to_bitmer(s: bytes):
    return lib().ntref_bytes_to_bitmer(s, len(s))


class Parsed:
    pass


def parse_fastx(data: bytes, cap=None, capacity=None, max_read=0):
    """Whole-buffer parse (capacity=None) or the reference's incremental readers with a `capacity`-byte buffer.
    -> object with .format .line_ending .records (list of dict) .err (kind,line,id)"""
    arr = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(1, dtype=np.uint8)
    if cap is None:
        cap = data.count(b"\n") + 2
    recs = np.zeros((cap, 12), dtype=np.uint64)
    info = np.zeros(8, dtype=np.uint64)
    eid = C.create_string_buffer(512)
    if capacity is None:
        c = lib().ntref_parse_fastx(arr.ctypes.data, len(data), recs.ctypes.data, cap, info.ctypes.data, eid, 512)
    else:
        c = lib().ntref_parse_fastx_incremental(arr.ctypes.data, len(data), capacity, max_read, recs.ctypes.data, cap,
                                                info.ctypes.data, eid, 512)
    p = Parsed()
    p.format = FMT[int(info[0])]; p.line_ending = LE[int(info[1])]
    p.err_kind = ERR[int(info[2])]; p.err_line = int(info[3])
    p.err_id = eid.value if info[4] else None
    p.final_line = int(info[5]); p.final_byte = int(info[6])
    p.table = recs[:c]
    p.records = []
    for r in recs[:c]:
        r = [int(x) for x in r]
        p.records.append(dict(start=r[0], id=data[r[1]:r[2]], raw_seq=data[r[3]:r[4]],
                              qual=(data[r[5]:r[6]] if p.format == "fastq" else None),
                              all=data[r[0]:r[7]], num_bases=r[8], line=r[9], byte=r[10]))
    return p


def tally_fastx(data: bytes, k: int, m: int = 0, iupac: bool = False, query: bytes = None):
    arr = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(1, dtype=np.uint8)
    out = np.zeros(9, dtype=np.uint64)
    e = lib().ntref_tally_fastx(arr.ctypes.data, len(data), k, m, int(iupac), query, out.ctypes.data)
    d = {f: int(v) for f, v in zip(TALLY_FIELDS, out)}
    d["err_kind"] = ERR[e]
    return d


def gen_fastq(seed, rec0, nrec, L, n_thresh=0, nthreads=1):
    out = np.empty(nrec * (2 * L + 16), dtype=np.uint8)
    lib().ntref_gen_fastq(out.ctypes.data, seed, rec0, nrec, L, n_thresh, nthreads)
    return out


def gen_fasta(seed, rec0, nrec, L, n_thresh=0, nthreads=1):
    out = np.empty(nrec * (L + 12), dtype=np.uint8)
    lib().ntref_gen_fasta(out.ctypes.data, seed, rec0, nrec, L, n_thresh, nthreads)
    return out


def bench_fastq(buf: np.ndarray, rec_bytes: int, nthreads: int, k: int, m: int, iupac=False):
    """-> (tallies dict, seconds).  buf holds whole fixed-size records."""
    nrec = buf.size // rec_bytes
    offs = np.array([(nrec * i // nthreads) * rec_bytes for i in range(nthreads + 1)], dtype=np.uint64)
    out = np.zeros(9, dtype=np.uint64)
    secs = lib().ntref_bench_fastq(buf.ctypes.data, offs.ctypes.data, nthreads, k, m, int(iupac), out.ctypes.data)
    return {f: int(v) for f, v in zip(TALLY_FIELDS, out)}, secs
