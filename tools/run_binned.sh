#!/bin/bash
# Developer script (under gpurun [--gpus N]): the device-binned + NCCL workload on N GPUs -> gpurun_out/binned_N<N>.jsonl
N=${1:-1}
OUT=gpurun_out/binned_N${N}.jsonl
: > $OUT
run() {
  if [ "$N" = "1" ]; then python bench.py --gpus 1 "$@" >> $OUT 2>> gpurun_out/binned_N${N}.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N "$@" >> $OUT 2>> gpurun_out/binned_N${N}.err; fi
}
run --workload binned --rays 1e9 --steps 3 --warmup 2 --no-cpu-baseline
if [ "$N" = "8" ]; then run --workload binned --rays 1.25e10 --steps 1 --warmup 1 --no-cpu-baseline; fi
if [ "$N" = "1" ]; then run --workload binned --rays 1e9 --steps 2 --warmup 1; fi
python tools/show_configs.py $OUT
