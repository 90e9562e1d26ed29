#!/bin/bash
# Developer script (run under gpurun): the ncu evidence committed under profiles/ for one kernel version.
# usage: tools/capture_profiles.sh r02_v3   -> gpurun_out/<tag>_*.  Every ncu run follows a plain run of the same command.
v=${1:-r02_vX}
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-plugin > gpurun_out/plain_bench.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/${v}_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-plugin > gpurun_out/ncu_bench.log 2>&1
python tools/gpu_one.py lensesAndMirrors 2097152 1 4 > gpurun_out/plain_one.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 2 -c 1 -o gpurun_out/${v}_trace \
      python tools/gpu_one.py lensesAndMirrors 2097152 1 4 > gpurun_out/ncu_one.log 2>&1
python tools/gpu_one.py hugeArray 33554432 1 2 > gpurun_out/plain_huge.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv \
      --log-file gpurun_out/${v}_huge_launches.csv python tools/gpu_one.py hugeArray 33554432 1 1 > gpurun_out/ncu_huge.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wf_traverse4 -s 0 -c 2 -o gpurun_out/${v}_wf_traverse4 \
    python tools/gpu_one.py hugeArray 33554432 1 1 > gpurun_out/ncu_huge2.log 2>&1
python tools/gpu_binned.py lambertSource 2e7 1000 2 > gpurun_out/plain_binned.log 2>&1 && \
  ODW_RAYS_PER_LAUNCH=20000000 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 1 -c 1 -o gpurun_out/${v}_binned \
      python tools/gpu_binned.py lambertSource 2e7 1000 2 > gpurun_out/ncu_binned.log 2>&1
cat gpurun_out/plain_one.log gpurun_out/plain_huge.log gpurun_out/plain_binned.log
cut -c1-200 gpurun_out/plain_bench.log | tail -1
wc -l gpurun_out/${v}_launches.csv
