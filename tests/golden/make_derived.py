#!/usr/bin/env python3
"""Writes tests/golden/derived.json: oracle-computed tallies on the reference's data files
(SURVEY.md §8c "derived goldens").  Not reference-asserted; they pin the oracle against drift."""
import json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O
from conftest import load_fixtures

fx = load_fixtures()
out = {
    "28S_k4_AAAA": O.tally_fastx(fx["data/28S.fasta"], k=4, m=0, iupac=False, query=b"AAAA"),
    "28S_k31_iupac": O.tally_fastx(fx["data/28S.fasta"], k=31, m=21, iupac=True),
    "28S_k21_m11": O.tally_fastx(fx["data/28S.fasta"], k=21, m=11, iupac=False),
    "PRJNA_k31_m21": O.tally_fastx(fx["data/PRJNA271013_head.fq"], k=31, m=21, iupac=False),
    "PRJNA_k51": O.tally_fastx(fx["data/PRJNA271013_head.fq"], k=51, m=0, iupac=False),
}
json.dump(out, open(os.path.join(HERE, "derived.json"), "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1))
