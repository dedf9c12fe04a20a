import sys, os
sys.path.insert(0, os.getcwd())
import needletail_b200 as nt
ctx = nt.Context(0)
import ctypes

for nrec in (5000, 20000000):
    L=150; nb = nrec*(2*L+16)
    d = ctx.device_alloc(nb); ctx.synth_fastq_device(d, 0x5EED0002, 0, nrec, L, 0)
    t = ctx.tally_device(d, nb, k=31, m=21)
    t = ctx.tally_device(d, nb, k=31, m=21)
    print(nrec, 'fallback', t['fallback'], 'ws_handover', t['ws_handover'], 'n_kmers', t['n_kmers'], t['n_kmers']==nrec*120, t['ws_cycles'])
    ctx.device_free(d)
