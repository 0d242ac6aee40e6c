#!/bin/bash
# round 2, GPU call M (session 3): state check after the container was re-created: full GPU suite, bench line, small decks
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --maxfail=10 > $O/r2m_pytest.log 2>&1; echo "pytest rc=$?" > $O/r2m_steps.log
timeout 600 python bench.py > $O/r2m_bench.json 2> $O/r2m_bench.err; echo "bench rc=$?" >> $O/r2m_steps.log
timeout 300 python tools/small_deck_bench.py > $O/r2m_small.txt 2>&1; echo "small rc=$?" >> $O/r2m_steps.log
cat $O/r2m_steps.log; tail -3 $O/r2m_pytest.log; grep -E "^FAILED|^ERROR" $O/r2m_pytest.log | head; cat $O/r2m_bench.json; cat $O/r2m_small.txt
