#!/usr/bin/env bash
# compute-sanitizer over smoke() and a parity subset touching every round-2 kernel path (tables, staged gathers, Pupko v2, buckets, multi)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"; tail -3 gpurun_out/memcheck_smoke.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests -m gpu -x -q -k "small_trees or tables_on_small or priors or multi_device or ragged or epsilon or pupko_states" > gpurun_out/memcheck_tests.log 2>&1; echo "memcheck tests rc=$?"; tail -4 gpurun_out/memcheck_tests.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?"; grep -c "Race reported" gpurun_out/racecheck_smoke.log; tail -3 gpurun_out/racecheck_smoke.log
