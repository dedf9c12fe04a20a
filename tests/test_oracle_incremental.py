"""The reference's readers are incremental (64 KiB buffer that is refilled, shifted and grown:
src/parser/fastq.rs:312-384, src/parser/fasta.rs:250-287); the oracle used for GPU parity — and the GPU itself — work on
the whole buffer.  These tests run a literal restatement of the incremental readers (oracle/ntref_incremental.hpp) with
buffer capacities from 3 bytes up and with short reads, and require results identical to the whole-buffer oracle: the
reference's output does not depend on its buffer capacity.  CPU only."""
import random

import numpy as np

import oracle_lib as O
from conftest import load_fixtures

CAPS = (3, 4, 5, 7, 16, 61, 64, 1000, 65536)


def same(a, b, what):
    assert a.format == b.format, what
    assert a.err_kind == b.err_kind, (what, a.err_kind, b.err_kind)
    assert (a.err_line, a.err_id) == (b.err_line, b.err_id), what
    assert a.line_ending == b.line_ending, what
    ta, tb = a.table.copy(), b.table.copy()
    assert ta.shape == tb.shape, what
    if len(ta):
        empty = ta[:, 3] == ta[:, 4]
        ta[empty, 3] = ta[empty, 4] = tb[empty, 3] = tb[empty, 4] = 0
    assert np.array_equal(ta, tb), what
    assert (a.final_line, a.final_byte) == (b.final_line, b.final_byte), what


def check(data, caps=CAPS):
    whole = O.parse_fastx(data)
    for cap in caps:
        for max_read in (0, 1, 7):
            same(O.parse_fastx(data, capacity=cap, max_read=max_read), whole, (data[:60], cap, max_read))


def test_vectors_any_capacity():
    vecs = [b"@test\nAGCT\n+test\n~~a!\n@test2\nTGCA\n+test\nWUI9", b"@test\r\nAGCT\r\n+test\r\n~~a!\r\n@test2\r\nTGCA\r\n+test\r\nWUI9",
            b"@test\nACGT\n+\nIII", b"@test\nAGCT\n+test\n~~a!\n@test2\nTGCA", b"@test\nAGCT\n+test\n~~a!\n\n",
            b"@test\nAGCT\n+test\n~~a!\n\n@TEST\nA\n+TEST\n~", b"@\n\n+\n\n@test2\nTGCA\n+test2\n~~~~\n", b"@a\nAC\n+\nII\n\n\n\n",
            b">test\nACGT\n>test2\nTGCA\n", b">test\nACGT\nACGT\n>test2\nTGCA\nTG", b">test\r\nACGT\r\nACGT\r\n>test2\r\nTGCA\r\nTG",
            b">test\nAGCT\n>test2", b">test\r\nAGCT\r\n>test2\r\n", b">\n\n>shine\nAGGAGGU", b">a\n", b">a\n\n", b">a\n>b\nAC", b">x\nAC>GT\n\n>y\n"]
    for v in vecs:
        check(v)


def test_random_bytes_any_capacity():
    rng = random.Random(4321)
    alphabet = b"ACGTN\n\n\r@>+ I~-"
    for it in range(400):
        n = rng.randrange(0, 300)
        body = bytes(rng.choice(alphabet) for _ in range(n))
        check((b"@" if it % 2 else b">") + body, caps=(3, 5, 16, 64, 65536))


def test_fixture_files_default_capacity():
    """every reference fixture, 64 KiB buffer (the reference's default, parser/utils.rs:8) and a 1000-byte one"""
    for name, data in sorted(load_fixtures().items()):
        if name.endswith((".toml", ".gz", ".bz2", ".xz", ".zst")):
            continue
        whole = O.parse_fastx(data)
        for cap in (1000, 65536):
            same(O.parse_fastx(data, capacity=cap), whole, (name, cap))
