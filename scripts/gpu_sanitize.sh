#!/bin/bash
# compute-sanitizer (memcheck, then racecheck on the shared-memory stacks / film tiles) over smoke() and small parity cases of the final code.
T=${1:-r02_y}
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"; tail -3 gpurun_out/${T}_memcheck_smoke.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_textures.py tests/test_refit.py tests/test_prb_reparam.py -m gpu -q -x -k "textured_scene_parity or refit_matches or gpu_kernel" > gpurun_out/${T}_memcheck_tests.log 2>&1; echo "memcheck tests rc=$?"; tail -4 gpurun_out/${T}_memcheck_tests.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?"; tail -3 gpurun_out/${T}_racecheck_smoke.log
