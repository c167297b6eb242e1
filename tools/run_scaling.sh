#!/bin/bash
# Multi-GPU lines kept under profiles/ (run under `gpurun --gpus N`): bash tools/run_scaling.sh <N> <tag> <what...>
# what: c2 | both | c3 | c4small
set -u
N=$1; tag=$2; shift 2
out=gpurun_out
mkdir -p $out
run() {  # name, bench args...
  name=$1; shift
  if [ "$N" = 1 ]; then python bench.py --gpus 1 "$@" > $out/${tag}_${name}_n${N}.json 2> $out/${tag}_${name}_n${N}.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N "$@" > $out/${tag}_${name}_n${N}.json 2> $out/${tag}_${name}_n${N}.err; fi
  tail -1 $out/${tag}_${name}_n${N}.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name n=$N', round(d['value'],1), d['unit'], round(d['ms_per_step'],1), 'ms/step', d['scaling'], d['config']['global_batch'], 'step_roofline', round(d['step_roofline']['frac'],3))" || tail -5 $out/${tag}_${name}_n${N}.err
}
for w in "$@"; do
  case $w in
    c2) run c2 --steps 5 --warmup 3 --no-cpu-baseline --no-extras ;;
    both) run c2_both --steps 5 --warmup 3 --model both --no-cpu-baseline --no-extras ;;
    c3) run c3 --workload c3 --micro-batch 128 --steps 2 --warmup 3 --no-cpu-baseline --no-extras ;;
    c4small) run c4_b128pergpu --workload c4 --batch 128 --micro-batch 8 --steps 1 --warmup 3 --no-cpu-baseline --no-extras ;;
  esac
done
