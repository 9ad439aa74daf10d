#!/bin/bash
# Round-end check on one B200 (under gpurun): the whole GPU test suite, then smoke().
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/verify_tests.log 2>&1
tail -4 $O/verify_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/verify_smoke.log 2>&1
tail -2 $O/verify_smoke.log
