#!/bin/bash
# round 2: scaling runs on one 8-GPU box (gpurun --gpus 8): weak and strong scaling of the train step (c4) and the
# C5 inference sweep at N = 1, 2, 4, 8.  One JSON line per run in gpurun_out/scale_*.json.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > $O/smi_multi.txt 2>&1
run() {  # name N args...
  name=$1; n=$2; shift 2
  if [ "$n" = 1 ]; then
    timeout 300 python bench.py --gpus 1 --no-cpu-baseline --no-gpu-incumbent "$@" > $O/scale_${name}_n$n.json 2> $O/scale_${name}_n$n.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --no-cpu-baseline --no-gpu-incumbent "$@" > $O/scale_${name}_n$n.json 2> $O/scale_${name}_n$n.err
  fi
  echo "$name N=$n rc=$? $(grep -o '"value": [0-9.e+]*, "unit": "ray-samples/s", "n_gpus": [0-9]*, "steps": [0-9]*, "warmup": [0-9]*, "ms_per_step": [0-9.]*' $O/scale_${name}_n$n.json | head -1)"
}
for n in 1 2 4 8; do run weak_train $n --steps 20 --warmup 5; done
for n in 1 2 4 8; do run strong_train $n --scaling strong --steps 20 --warmup 5; done
for n in 1 2 4 8; do run strong_c5 $n --workload c5 --scaling strong --steps 5 --warmup 3; done
for n in 8; do run weak_c5 $n --workload c5 --steps 5 --warmup 3; done
ls $O | grep scale_ | head -40
