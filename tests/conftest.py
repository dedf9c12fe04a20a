import gzip, hashlib, io, json, os, sys, tarfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run under gpurun)")


def pytest_sessionstart(session):
    """A fresh checkout has no built artefacts (they are git-ignored): build the product library once if it is MISSING
    (same recipe as __graft_entry__.build(); an existing library is never rebuilt here — on the GPU box the one that
    travelled with the snapshot is what must be tested)."""
    so = os.environ.get("NTGPU_SO", os.path.join(ROOT, "needletail_b200", "libntgpu.so"))
    if not os.path.exists(so) and "NTGPU_SO" not in os.environ:
        from needletail_b200 import build as nt_build
        nt_build.build()


def _have_sm100_gpu():
    try:
        import ctypes as C
        import needletail_b200 as nt
        lib = nt.load_library()
        h = C.c_void_p()
        if lib.ntg_create(0, C.byref(h)) != 0:
            return False
        lib.ntg_destroy(h)
        return True
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests are skipped (not errors) on a box without an sm_100 device (round-1 ADVICE)."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if gpu_items and not _have_sm100_gpu():
        skip = pytest.mark.skip(reason="no sm_100 CUDA device: run under gpurun")
        for it in gpu_items:
            it.add_marker(skip)


_FIXTURES = None


def load_fixtures():
    """name -> bytes for every file of the reference's tests/data and tests/specimen (bundled)."""
    global _FIXTURES
    if _FIXTURES is None:
        out = {}
        with tarfile.open(os.path.join(GOLDEN, "ref_fixtures.tar.gz"), "r:gz") as tf:
            for m in tf.getmembers():
                out[m.name] = tf.extractfile(m).read()
        _FIXTURES = out
    return _FIXTURES


@pytest.fixture(scope="session")
def fixtures():
    return load_fixtures()


def parse_specimen_index(text):
    """Minimal reader for tests/specimen/*/index.toml: [[valid]]/[[invalid]] tables with
    filename = "..." and optional tags = [..] (ref: tests/format_specimens.rs:7-19)."""
    out = {"valid": [], "invalid": []}
    cur = None
    for line in text.splitlines():
        s = line.strip()
        if s in ("[[valid]]", "[[invalid]]"):
            cur = {"filename": None, "tags": []}
            out[s.strip("[]")].append(cur)
        elif cur is not None and s.startswith("filename"):
            cur["filename"] = s.split("=", 1)[1].strip().strip('"')
        elif cur is not None and s.startswith("tags"):
            cur["tags"] = [t.strip().strip('"') for t in s.split("=", 1)[1].strip().strip("[]").split(",") if t.strip()]
    return out
