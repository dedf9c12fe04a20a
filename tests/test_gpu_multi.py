"""Multi-GPU paths on real devices (run with gpurun --gpus 2): two ranks as two PROCESSES over NCCL — the in-stream all-reduce of
the tallies and the two spectrum reductions (dense all-reduce, hash-partitioned exchange).  Skipped with fewer than two GPUs."""
import collections
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, collections
sys.path.insert(0, os.environ["NT_ROOT"]); sys.path.insert(0, os.path.join(os.environ["NT_ROOT"], "tests"))
import numpy as np, torch, torch.distributed as dist
import needletail_b200 as nt
from needletail_b200 import shard
import oracle_lib as O
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
ctx = nt.Context(rank)
shard.init_nccl_comm(ctx)
L, total = 100, 40000
first, n = shard.shard_records(total, world, rank)
rb = 2 * L + 16
whole = O.gen_fastq(7, 0, total, L, 655).tobytes()
mine = whole[first * rb:(first + n) * rb]
d = ctx.device_alloc(len(mine) + 16); ctx.h2d(d, np.frombuffer(mine, dtype=np.uint8))
# (1) tallies: in-stream all-reduce == oracle over the whole input
exp = O.tally_fastx(whole, k=31, m=21)
ctx.tally_device_enqueue(d, len(mine), k=31, m=21, allreduce=True)
got = ctx.tally_device_collect()
assert not got["not_reduced"]
for key in ("n_records", "n_bases", "n_kmers", "n_not_rc", "kmer_sum_lo", "n_minimizers", "minimizer_sum"):
    assert got[key] == exp[key], (key, got[key], exp[key])
def oracle_spectrum(data, k):
    cnt = collections.Counter()
    for r in O.parse_fastx(data).records:
        _, km, _ = O.bit_kmers(O.normalize(r["raw_seq"], False)[0], k, True)
        cnt.update(int(v) for v in km)
    return cnt
# (2) dense spectrum: every rank ends with the job-wide histogram
want = oracle_spectrum(whole, 9)
sp = ctx.spectrum(9); sp.add_device(d, len(mine)); sp.reduce()
keys, counts = sp.items()
assert dict(zip(map(int, keys), map(int, counts))) == dict(want)
sp.close()
# (3) hash spectrum: every key ends on exactly one rank with its job-wide count
want = oracle_spectrum(whole, 25)
sp = ctx.spectrum(25, capacity=1 << 23); sp.add_device(d, len(mine)); sp.reduce()
keys, counts = sp.items()
mine_d = dict(zip(map(int, keys), map(int, counts)))
for key, c in mine_d.items():
    assert want[key] == c
sizes = torch.tensor([len(mine_d)], device="cuda"); dist.all_reduce(sizes)
assert int(sizes.item()) == len(want), (int(sizes.item()), len(want))
chk = torch.tensor([sum(mine_d.values())], device="cuda"); dist.all_reduce(chk)
assert int(chk.item()) == sum(want.values())
sp.close()
ctx.close()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_ranks_nccl(tmp_path):
    import needletail_b200 as nt
    import ctypes as C
    cnt = C.c_int(0)
    nt.load_library().ntg_device_count(C.byref(cnt))
    if cnt.value < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, NT_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout
