#!/bin/bash
# final 1-GPU call of the round: full -m gpu suite, the bench (both arms), then the ncu evidence of the final density kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/f_suite.log 2>&1
echo "suite exit $?" >> gpurun_out/f_suite.log; tail -5 gpurun_out/f_suite.log
timeout 900 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
echo "bench exit $?"; tail -c 600 gpurun_out/f_bench.json
timeout 600 python bench.py --impl reference > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err
echo "reference arm exit $?"; tail -c 400 gpurun_out/f_bench_ref.json
export PROBE_REPS=1
SEC="--section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section SchedulerStats --section WarpStateStats"
timeout 900 ncu $SEC --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum --replay-mode application --clock-control none -k regex:knn_hp -c 1 -f -o gpurun_out/r2_knn_hp_512 python scripts/gpu_knn_sweep.py 512 64 > gpurun_out/r2_knn_hp_512_ncu.log 2>&1
echo "ncu 512 exit $?"
bash scripts/gpu_ncu.sh r2_knn_hp_256 knn_hp 0 1 application -- python scripts/gpu_knn_sweep.py 256 64
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches_bench_512cube.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2_launches_bench.log 2>&1
echo "launch list exit $?"
