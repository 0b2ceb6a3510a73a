#!/bin/bash
# compute-sanitizer passes over the small GPU tests (memcheck / racecheck / synccheck / initcheck); logs under gpurun_out/.
# Usage (on the GPU box): bash tools/r02_sanitizer.sh
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
SMALL='mhd_l16_with_rigid or mhd_l16_chunked or stress_free_boundaries or hydro_minc3 or anelastic_hydro or full_sphere_centre or double_curl or phase_field or conducting'
run() {  # name, timeout, tool options..., -- command
    local name=$1 tmo=$2; shift 2
    local log=gpurun_out/sanitizer_$name.log
    ( time timeout $tmo $CS "$@" ) > $log 2>&1
    echo "== $name: rc=$? $(grep -c 'ERROR SUMMARY' $log) summaries"
    grep -h "ERROR SUMMARY\|passed\|failed\|RACECHECK SUMMARY\|^real" $log | sort | uniq -c | tail -12
    grep -h -m 40 "=========     at \|========= Invalid\|========= Uninitialized\|========= Error\|========= Race\|========= Barrier\|========= Warning" $log | cut -c1-220 | sort | uniq -c | sort -rn | head -20
}
run memcheck_smoke 200 --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()"
run memcheck_tests 480 --tool memcheck --print-limit 20 python -m pytest tests/test_sht_gpu.py tests/test_rloop_gpu.py tests/test_diagnostics.py tests/test_transpose_gpu.py tests/test_lm_side_gpu.py -m gpu -q -x -k "not l96 and not bitwise" -p no:cacheprovider
run racecheck_tests 360 --tool racecheck --racecheck-report all --print-limit 20 python -m pytest tests/test_rloop_gpu.py tests/test_diagnostics.py tests/test_lm_side_gpu.py -m gpu -q -x -k "$SMALL or diag or derivatives" -p no:cacheprovider
run synccheck_tests 240 --tool synccheck --print-limit 20 python -m pytest tests/test_rloop_gpu.py tests/test_diagnostics.py -m gpu -q -x -k "$SMALL or diag" -p no:cacheprovider
run initcheck_tests 240 --tool initcheck --print-limit 20 python -m pytest tests/test_rloop_gpu.py tests/test_diagnostics.py -m gpu -q -x -k "$SMALL or diag" -p no:cacheprovider
