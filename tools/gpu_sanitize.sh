#!/bin/bash
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${TAG}_sanitizer_memcheck.txt 2>&1; tail -2 gpurun_out/${TAG}_sanitizer_memcheck.txt
timeout 600 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${TAG}_sanitizer_racecheck.txt 2>&1; tail -2 gpurun_out/${TAG}_sanitizer_racecheck.txt
