#!/bin/bash
# Developer script (run under gpurun [--gpus N]): the BASELINE.json configs at their stated sizes on N GPUs; every bench.py
# JSON line is appended to gpurun_out/configs_N<N>.jsonl (copied to profiles/configs_r02.jsonl afterwards).
# usage: tools/run_configs.sh N
N=${1:-1}
OUT=gpurun_out/configs_N${N}.jsonl
: > $OUT
run() {   # args: bench.py arguments
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 "$@" >> $OUT 2>> gpurun_out/configs_N${N}.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" >> $OUT 2>> gpurun_out/configs_N${N}.err
  fi
  tail -c 300 $OUT | tr '\n' ' '; echo
}
per() { python -c "print(int($1/$N))"; }
if [ "$N" = "1" ]; then
  run --scene minimal --rays 1e5 --steps 20 --warmup 5                                    # configs[0]: launch-latency case
  run --steps 10 --warmup 3                                                                # configs[1]: the headline line
  run --scene hugeArray --rays 1.25e9 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline     # one GPU's share of configs[3]
fi
if [ "$N" = "8" ]; then
  run --steps 5 --warmup 3 --no-plugin                                                     # configs[1] on 8 GPUs (weak), e2e through the C ABI
  run --scene hugeArray --rays 1.25e9 --steps 2 --warmup 1 --no-e2e                        # configs[3]: 1e10 rays over 8 GPUs
  run --workload binned --rays 1.25e10 --steps 1 --warmup 1                                # configs[4]: 1e11 rays, binned, NCCL-reduced
fi
run --scene lensesAndMirrorsSequential --rays $(per 1e9) --steps 2 --warmup 1 --no-e2e --no-cpu-baseline    # configs[2]: 1e9 rays at 1/2/4/8 GPUs
run --workload binned --rays 1e9 --steps 3 --warmup 2 --no-cpu-baseline                                    # scaling record of the binned path (weak)
wc -l $OUT
