#!/bin/bash
# compute-sanitizer memcheck over the C++ parity program (ram_permutation, sort_decommittment_requests, demux_log_queue) and
# over smoke() (ram_permutation + main_vm).  Run on a GPU box: bash tools/sanitize.sh > gpurun_out/sanitize.log 2>&1
set -u
cd "$(dirname "$0")/.."
make -s -C oracle
g++ -std=c++17 -O1 -I include -I oracle tests/cpp/host_mirror_test.cpp -o /tmp/host_mirror_test \
    era_zkevm_circuits_b200/libzkc_b200.so oracle/liborc.so -Wl,-rpath,$PWD/era_zkevm_circuits_b200 -Wl,-rpath,$PWD/oracle -ldl -lpthread
echo "== memcheck: C++ parity program"
timeout 150 compute-sanitizer --tool memcheck --error-exitcode 9 /tmp/host_mirror_test parity 2>&1 | tail -12
echo "exit: ${PIPESTATUS[0]}"
echo "== memcheck: smoke()"
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -12
echo "exit: ${PIPESTATUS[0]}"
for tool in racecheck initcheck synccheck; do
    echo "== $tool: C++ parity program"
    timeout 200 compute-sanitizer --tool $tool --error-exitcode 9 /tmp/host_mirror_test parity 2>&1 | tail -8
    echo "exit: ${PIPESTATUS[0]}"
done
echo "== racecheck: smoke()"
timeout 240 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8
echo "exit: ${PIPESTATUS[0]}"
