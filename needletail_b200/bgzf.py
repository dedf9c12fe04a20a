"""BGZF writer (SAM spec 4.1): gzip members of at most 64 KiB with the member size in the extra field, so that the members
can be inflated independently (ntg_stream_feed_gz with threads > 1).  Host-side format helper: used by the tests and by
bench.py's compressed-input pipeline (BASELINE config C5); any gzip reader (flate2::MultiGzDecoder, zcat) reads the result."""
import struct
import zlib

BLOCK = 0xFF00          # uncompressed bytes per member (bgzip's choice: always fits a 64 KiB member)
EOF_MARKER = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def member(chunk: bytes, level: int = 1) -> bytes:
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    payload = co.compress(chunk) + co.flush()
    bsize = 12 + 6 + len(payload) + 8
    assert bsize <= 0x10000
    return (b"\x1f\x8b\x08\x04" + b"\x00\x00\x00\x00" + b"\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1)
            + payload + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))


def compress(data, level: int = 1, eof_marker: bool = True) -> bytes:
    mv = memoryview(data)
    out = [member(bytes(mv[o:o + BLOCK]), level) for o in range(0, len(mv), BLOCK)]
    if eof_marker:
        out.append(EOF_MARKER)
    return b"".join(out)
