#!/bin/bash
# ncu evidence for profiles/: final density kernel (application replay: the persistent kernel does not survive kernel replay),
# FOF link kernels and the build kernels (kernel replay), and the launch list of the bench command
mkdir -p gpurun_out
export PROBE_REPS=1
# 1. density kernel, 512^3 (the bench workload): DRAM bytes + speed-of-light + memory + occupancy
SEC="--section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section SchedulerStats --section WarpStateStats"
timeout 1500 ncu $SEC --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum --replay-mode application --clock-control none -k regex:knn_hp -c 1 -f -o gpurun_out/r2_knn_hp_512 python scripts/gpu_knn_sweep.py 512 64 > gpurun_out/r2_knn_hp_512_ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/r2_knn_hp_512_ncu.log; tail -2 gpurun_out/r2_knn_hp_512_ncu.log
# 2. density kernel, 256^3: all sections with source counters
bash scripts/gpu_ncu.sh r2_knn_hp_256 knn_hp 0 1 application -- python scripts/gpu_knn_sweep.py 256 64
# 3. FOF link kernels (3D and 6D) and union-find, 256^3
bash scripts/gpu_ncu.sh r2_fof_256 "fof_link|fof_roots|uf_" 0 12 kernel -- python scripts/gpu_probe_fof_build.py 256
# 4. build kernels, 256^3 (first build only: ~ a few dozen launches)
bash scripts/gpu_ncu.sh r2_build_256 "v2_" 0 60 kernel -- python scripts/gpu_probe_fof_build.py 256
# 5. launch list of the bench command (never a bench value)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches_bench_512cube.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2_launches_bench.log 2>&1
echo "launch list exit $?"; tail -2 gpurun_out/r2_launches_bench.log
ls -la gpurun_out/*.ncu-rep
