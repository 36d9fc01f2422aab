#!/bin/bash
# round 2, final code: scaling runs on one 8-GPU box (gpurun --gpus 8).  Weak (4096 rays per GPU) and strong (4096
# rays in all = BASELINE configs[3] literally) scaling of the train step, the strong step again as ONE CUDA graph per
# rank (NCCL all-reduce captured), and the C5 inference sweep.  One JSON line per run in gpurun_out/scale2_*.json.
mkdir -p gpurun_out
O=gpurun_out
run() {  # name N args...
  name=$1; n=$2; shift 2
  if [ "$n" = 1 ]; then
    timeout 200 python bench.py --gpus 1 --no-cpu-baseline --no-gpu-incumbent "$@" > $O/scale2_${name}_n$n.json 2> $O/scale2_${name}_n$n.err
  else
    timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --no-cpu-baseline --no-gpu-incumbent "$@" > $O/scale2_${name}_n$n.json 2> $O/scale2_${name}_n$n.err
  fi
  echo "$name N=$n rc=$? $(grep -o '"value": [0-9.e+]*, "unit": "ray-samples/s", "n_gpus": [0-9]*, "steps": [0-9]*, "warmup": [0-9]*, "ms_per_step": [0-9.]*' $O/scale2_${name}_n$n.json | head -1) $(grep -o '"graphed": {"ms_per_step": [0-9.]*, "value": [0-9.e+]*, "e2e_ms_per_step": [0-9.]*' $O/scale2_${name}_n$n.json | head -1)"
}
for n in 1 8; do run weak_train $n --steps 20 --warmup 5; done
for n in 1 4 8; do run strong_train $n --scaling strong --steps 30 --warmup 5; done
run strong_train_graph 8 --scaling strong --steps 30 --warmup 5 --graph
for n in 1 8; do run strong_c5 $n --workload c5 --scaling strong --steps 5 --warmup 3; done
