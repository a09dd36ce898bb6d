#!/bin/bash
# compute-sanitizer passes over small GPU tests: racecheck on a broad subset (memcheck / synccheck were clean in the previous pass)
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 2400 $CS --tool racecheck --error-exitcode 7 --print-limit 40 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
  -k "golden or build_structure or duplicates or odd_k or attached_halo or fof_linked or checked_fof or density_kernel_variants" > gpurun_out/n_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/n_racecheck.log; tail -6 gpurun_out/n_racecheck.log
grep "Race reported\|hazard" gpurun_out/n_racecheck.log | sed 's/(.*in / in /' | sort | uniq -c | sort -rn | head -20
