#!/bin/bash
# round 2, session 2, last GPU call: the whole GPU suite and smoke() on the final build.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -x -q -m gpu > $O/r3h_pytest.log 2>&1; echo rc=$? >> $O/r3h_pytest.log; tail -n 3 $O/r3h_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2 | tee $O/r3h_smoke.log
