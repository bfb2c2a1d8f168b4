#!/bin/bash
# compute-sanitizer over the frame kernels (run on the GPU box: gpurun -- 'bash tools/sanitize.sh').
# memcheck: out-of-bounds / misaligned accesses; racecheck: shared-memory hazards between the phases of k_tile and
# k_front; synccheck: divergent barriers.  Logs under gpurun_out/; summaries are copied to profiles/.
mkdir -p gpurun_out
T=${TAG:-r02}
for tool in memcheck racecheck synccheck; do
  timeout ${SAN_TIMEOUT:-600} compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_frames.py > gpurun_out/${T}_sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize frames ok|Error|hazard" gpurun_out/${T}_sanitizer_$tool.log | head -8
done
