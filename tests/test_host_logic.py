"""CPU-only tests: the C-ABI library loads and exports every symbol include/ntgpu.h declares, fails loudly
without a GPU, and the multi-rank host logic (sharding + tallies reduction) works at world_size 2 (gloo)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "ntgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ntg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import needletail_b200 as nt
    lib = nt.load_library()
    names = header_functions()
    assert len(names) >= 39
    for n in names:
        assert hasattr(lib, n), f"libntgpu.so does not export {n}"
    assert sorted(lib._declared) == names          # the ctypes face covers the whole header
    assert lib.ntg_abi_version() == 3


def test_struct_layouts_match_header():
    import needletail_b200 as nt
    assert C.sizeof(nt._Record) == 80 and C.sizeof(nt._Tallies) == 128
    assert C.sizeof(nt._TallyConfig) == 88 and C.sizeof(nt._ParseError) == 264


def test_no_cpu_fallback():
    """Without a CUDA device the product must fail loudly, never compute on the CPU."""
    import needletail_b200 as nt
    lib = nt.load_library()
    cnt = C.c_int(-1)
    st = lib.ntg_device_count(C.byref(cnt))
    if st == 0 and cnt.value > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(nt.NtgError):
        nt.Context(0)
    with pytest.raises(nt.NtgError):
        nt.normalize_seq("ACGT")


def test_decompress_host_side(fixtures):
    import needletail_b200 as nt
    plain = fixtures["data/test.fa"]
    assert nt._decompress(fixtures["data/test.fa.gz"]) == plain       # tests/test_compressed.rs:21-33
    assert nt._decompress(fixtures["data/test.fa.bz2"]) == plain
    assert nt._decompress(fixtures["data/test.fa.xz"]) == plain
    import gzip
    two = gzip.compress(b">a\nAC\n") + gzip.compress(b">b\nGT\n")      # multi-member == MultiGzDecoder
    assert nt._decompress(two) == b">a\nAC\n>b\nGT\n"
    assert nt._decompress(b">plain") == b">plain"


def test_shard_records_partition():
    from needletail_b200.shard import shard_records
    for n in (0, 1, 7, 100, 100_000_000):
        for w in (1, 2, 3, 4, 8):
            parts = [shard_records(n, w, r) for r in range(w)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            for (a, c), (b, _) in zip(parts, parts[1:]):
                assert a + c == b
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


WORKER = r'''
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import torch.distributed as dist
import oracle_lib as O
from needletail_b200.shard import shard_records, allreduce_tallies
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
N, L, seed = 4001, 150, 0x5EED0004
first, cnt = shard_records(N, 2, rank)
mine = O.tally_fastx(O.gen_fastq(seed, first, cnt, L, 655).tobytes(), k=31, m=21)
mine.pop("err_kind")
tot = allreduce_tallies(mine)
whole = O.tally_fastx(O.gen_fastq(seed, 0, N, L, 655).tobytes(), k=31, m=21)
whole.pop("err_kind")
assert tot == whole, (tot, whole)
assert tot["n_records"] == N
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_rank_gloo_sharded_tallies(tmp_path):
    """N>1 path on CPU: each rank tallies its record shard (oracle as the compute stand-in), one
    all-reduce of the tallies vector, result equals the single-rank tallies of the whole input."""
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=300)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_lookback_state_monoid(tmp_path):
    """combine()/identity_state() of the fused kernel's carried scan state form a monoid and folding per-span aggregates
    equals the state of the whole span (host build of the very same header, 20 000 random strings)."""
    import shutil
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "test_state_monoid")
    subprocess.check_call([nvcc, "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "test_state_monoid.cu")], stderr=subprocess.DEVNULL)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "state monoid ok" in out.stdout, out.stdout + out.stderr


def test_walkers_match_oracle_on_host(tmp_path):
    """find_ws / walk / walk_fast / walk_clean of the fused kernel are __host__ __device__: the very same source, compiled
    for the CPU, must reproduce the oracle's tallies on random lines (clean, non-ACGT, whitespace, T runs), whole and cut
    into warmed-up fragments (tests/cpp/test_walkers.cu)."""
    import shutil
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "test_walkers")
    subprocess.check_call([nvcc, "-std=c++17", "--expt-relaxed-constexpr", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "test_walkers.cu")], stderr=subprocess.DEVNULL)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "walkers ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_decode_phred_vectors():
    """quality.rs:34-64 and test_python.py:152-168 (host-only function of the Python face)."""
    import needletail_b200 as nt
    want = (2, 27, 14, 27, 14, 33, 33, 37, 37, 37, 33, 37, 27)
    assert nt.decode_phred("#</</BBFFFBF<") == want
    assert nt.decode_phred("B[N[Naaeeeae[", base_64=True) == want
    assert nt.decode_phred("") == ()
    with pytest.raises(ValueError, match="character ' ' cannot be decoded with offset '33'"):
        nt.decode_phred("#</</BBFFFBF ")
    with pytest.raises(ValueError, match=r"character '\?' cannot be decoded with offset '64'"):
        nt.decode_phred("B[N[Naaeeeae?", base_64=True)


def test_record_class_mirrors_reference_python_module():
    """test_python.py:17-99 (RecordClassTestCase) minus normalize(), which runs on the device (tests/test_gpu_parity.py)."""
    from needletail_b200 import Record
    r = Record("test description", "AGCTGATCGA")
    assert (r.id, r.seq, r.qual) == ("test description", "AGCTGATCGA", None)
    assert (r.name, r.description) == ("test", "description")
    assert Record("test", "A").description is None and Record("test", "A").name == "test"
    q = Record("test description", "AGCTGATCGA", ";**9;;????")
    assert q.qual == ";**9;;????" and q.is_fastq() and not q.is_fasta() and r.is_fasta() and not r.is_fastq()
    with pytest.raises(ValueError):
        Record("x", "ACGT", "II")
    r1, r2 = Record("test", "AGCTGATCGA", ";**9;;????"), Record("test", "AGCTGATCGA", ";**9;;????")
    others = [Record("test2", "AGCTGATCGA", ";**9;;????"), Record("test", "TCGATCAGCT", ";**9;;????"),
              Record("test", "AGCTGATCGA", "????;**9;;"), Record("test", "AGCTGATCGA")]
    assert r1 == r2 and hash(r1) == hash(r2) and all(r1 != o for o in others) and all(hash(r1) != hash(o) for o in others)
    assert str(Record("test", "AGCTGATCGA")) == ">test\nAGCTGATCGA\n"
    assert str(Record("test", "AGCTGATCGA", ";**9;;????")) == "@test\nAGCTGATCGA\n+\n;**9;;????\n"
    assert repr(Record("test", "AGCTGATCGAAGCTGATCGAA")) == "Record(id=test, seq=AGCTGATCGAAGCTGA\u2026GAA, qual=None)"
    assert repr(Record("test", "AGCTGATCGAAGCTGATCGAA", ";**9;;????;**9;;????;")) == \
        "Record(id=test, seq=AGCTGATCGAAGCTGA\u2026GAA, qual=;**9;;????;**9;;\u2026??;)"
    assert repr(Record("test more", "ACGT")) == "Record(id=test\u2026, seq=ACGT, qual=None)"
    assert len(Record("test", "AGCTGATCGA")) == 10


def test_record_normalize_plumbing():
    """Record.normalize / normalize_seq hand the bytes to Context.normalize and put the str back (test_python.py:37-42);
    the context here is a stand-in answering with the oracle, the real device path is tests/test_gpu_parity.py."""
    import needletail_b200 as nt
    import oracle_lib as O

    class FakeCtx:
        def normalize(self, seqs, iupac=False):
            res = [O.normalize(s, iupac) for s in seqs]
            return [r[0] for r in res], [r[1] for r in res]

    r = nt.Record("test", "AGCTGYrtcga")
    r.normalize(iupac=True, ctx=FakeCtx())
    assert r.seq == "AGCTGYRTCGA"
    r.normalize(ctx=FakeCtx())
    assert r.seq == "AGCTGNNTCGA"


def test_record_from_table_row():
    """Records yielded by a reader are cut out of the input by the ten offsets of an ntg_record row (include/ntgpu.h)."""
    from types import SimpleNamespace
    import needletail_b200 as nt
    data = np.frombuffer(b"@id one\nAC\r\n+\nII\r\n>x\nAC\nGT\n", dtype=np.uint8)
    row = SimpleNamespace(start=0, all_e=17, id_b=1, id_e=7, seq_b=8, seq_e=10, qual_b=14, qual_e=16, num_bases=2, line=1)
    r = nt.Record._from_table(data, row, "fastq")
    assert (r.id, r.seq, r.qual, r.raw_seq, r.num_bases, r.line, r.byte) == ("id one", "AC", "II", b"AC", 2, 1, 0)
    assert r.all == b"@id one\nAC\r\n+\nII\r" and r.name == "id" and r.description == "one" and r.is_fastq()
    row = SimpleNamespace(start=18, all_e=26, id_b=19, id_e=20, seq_b=21, seq_e=26, qual_b=0, qual_e=0, num_bases=4, line=5)
    r = nt.Record._from_table(data, row, "fasta")
    assert (r.id, r.seq, r.qual, r.raw_seq, r.num_bases) == ("x", "ACGT", None, b"AC\nGT", 4) and r.is_fasta()
    assert str(r) == ">x\nACGT\n" and r == nt.Record("x", "ACGT")


def test_experiment_builds_compile(tmp_path):
    """The diagnostic build (-DNTG_STATS=1: per-CTA cycle accounting) stays buildable."""
    import shutil
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    flags = ["-DNTG_STATS=1"]
    out = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "--expt-relaxed-constexpr", *flags, "-cubin",
                          "-o", str(tmp_path / "all_on.cubin"), os.path.join(ROOT, "needletail_b200", "csrc", "ntgpu.cu")],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
