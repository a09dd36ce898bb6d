#!/bin/bash
# ncu capture with the standard sections.  usage: gpu_ncu.sh TAG KERNEL_REGEX SKIP COUNT kernel|application -- command...
mkdir -p gpurun_out
TAG=$1; REGEX=$2; SKIP=$3; COUNT=$4; MODE=$5; shift 6
SECTIONS="--section SpeedOfLight --section SchedulerStats --section WarpStateStats --section SourceCounters --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section InstructionStats"
timeout 1500 ncu $SECTIONS --metrics dram__bytes_read.sum,dram__bytes_write.sum --replay-mode $MODE --clock-control none --import-source on -k regex:$REGEX -s $SKIP -c $COUNT -f -o gpurun_out/${TAG} "$@" > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/${TAG}_ncu.log
tail -3 gpurun_out/${TAG}_ncu.log
ls -la gpurun_out/${TAG}.ncu-rep
