#!/bin/bash
# ncu capture of the density kernel (one launch).  usage: gpu_ncu_knn.sh NG TAG kernel|application "sweep settings" [extra ncu args...]
mkdir -p gpurun_out
NG=${1:-256}
TAG=${2:-knn}
MODE=${3:-kernel}
SET=${4:-}
shift 4
SECTIONS="--section SpeedOfLight --section SchedulerStats --section WarpStateStats --section SourceCounters --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section InstructionStats"
PROBE_REPS=1 timeout 1200 ncu $SECTIONS --replay-mode $MODE --clock-control none --import-source on -k regex:knn_ -s 0 -c 1 -f -o gpurun_out/${TAG} "$@" python scripts/gpu_knn_sweep.py $NG 64 $SET > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/${TAG}_ncu.log
tail -4 gpurun_out/${TAG}_ncu.log
ls -la gpurun_out/${TAG}.ncu-rep
