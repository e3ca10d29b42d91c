#!/bin/bash
# compute-sanitizer over every kernel family (scripts/sanitize_target.py): memcheck (global / shared accesses), racecheck
# (shared-memory hazards of the warp-synchronous buffers: PoolSink, the select kernel's histograms, the batched kernel's
# staging windows) and synccheck.  The device-side waits (peer records, pipelined hand-overs) are bounded by
# TKS_SPIN_TIMEOUT_MS / TKS_TAU_WAIT_US; under the sanitizer kernels run 10-100x slower, so the bounds are raised.
#   scripts/sanitize.sh [out_dir]       logs: <out_dir>/sanitize_{memcheck,racecheck,synccheck}.log
out=${1:-gpurun_out}
mkdir -p "$out"
export TKS_SPIN_TIMEOUT_MS=600000 TKS_TAU_WAIT_US=60000000
rc=0
for tool in memcheck racecheck synccheck; do
    timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python scripts/sanitize_target.py > "$out/sanitize_$tool.log" 2>&1
    code=$?
    echo "$tool: exit $code: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$out/sanitize_$tool.log" | tail -1)"
    [ $code -ne 0 ] && rc=1
done
exit $rc
