timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
NT_MC_ONLY=0,2,3,4 bash tools/ab_run2.sh cur noearly stats cur
