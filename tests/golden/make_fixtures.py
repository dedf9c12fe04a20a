#!/usr/bin/env python3
"""Bundle the reference's *test data* (not source code) into tests/golden/ref_fixtures.tar.gz.

/root/reference does not exist on the GPU box, so the parity inputs the reference's own
tests use (tests/data/*, tests/specimen/{FASTA,FASTQ}/* — FASTA/FASTQ files and the
specimen index.toml) travel as one deterministic tarball.  Run in the build container:

    python tests/golden/make_fixtures.py

It also writes MANIFEST.json (path -> sha256, size) so tests can verify the bundle and,
when /root/reference is mounted, that the bundle still equals the reference's files.
"""
import gzip, hashlib, io, json, os, tarfile

REF = "/root/reference/tests"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "ref_fixtures.tar.gz")


def main():
    files = []
    for sub in ("data", "specimen/FASTA", "specimen/FASTQ"):
        d = os.path.join(REF, sub)
        for name in sorted(os.listdir(d)):
            p = os.path.join(d, name)
            if os.path.isfile(p):
                files.append((f"{sub}/{name}", p))
    manifest = {}
    raw = io.BytesIO()
    with tarfile.open(fileobj=raw, mode="w", format=tarfile.USTAR_FORMAT) as tf:
        for arc, p in files:
            data = open(p, "rb").read()
            ti = tarfile.TarInfo(arc)
            ti.size = len(data); ti.mtime = 0; ti.mode = 0o644; ti.uid = ti.gid = 0; ti.uname = ti.gname = ""
            tf.addfile(ti, io.BytesIO(data))
            manifest[arc] = {"sha256": hashlib.sha256(data).hexdigest(), "size": len(data)}
    with open(OUT, "wb") as f:
        with gzip.GzipFile(fileobj=f, mode="wb", mtime=0, compresslevel=9) as gz:
            gz.write(raw.getvalue())
    with open(os.path.join(HERE, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print(f"{len(files)} files -> {OUT} ({os.path.getsize(OUT)} bytes)")


if __name__ == "__main__":
    main()
