#!/bin/bash
# Two-GPU call: the --workload lines (spectrogram channel shard, resample range shard, filt column shard; Welch real with
# the PSD all-reduce), each validating its own outputs in-run, then the headline bench (weak + strong legs).   usage: bash profiles/gpu_round_2gpu.sh <tag>
tag=${1:-r2m}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 20 --warmup 5 "${@:2}"; }
step() { local name=$1 to=$2; shift 2; local s=$(date +%s); timeout "$to" "$@"; echo "$name rc=$? $(( $(date +%s) - s ))s" | tee -a gpurun_out/${tag}_steps.log; }
port=29511
for w in resample spectrogram filt_columns welch_real; do
  step $w 150 bash -c "$(declare -f run); run $port --workload $w >> gpurun_out/${tag}_bench_workloads_2gpu.jsonl 2>> gpurun_out/${tag}_bench_workloads_2gpu.err"
  port=$((port + 1))
done
step filt_columns_td 150 bash -c "$(declare -f run); run $port --workload filt_columns --filt-alg td >> gpurun_out/${tag}_bench_workloads_2gpu.jsonl 2>> gpurun_out/${tag}_bench_workloads_2gpu.err"
cat gpurun_out/${tag}_bench_workloads_2gpu.jsonl | cut -c1-400
tail -5 gpurun_out/${tag}_bench_workloads_2gpu.err
step bench2 300 bash -c "$(declare -f run); run 29520 > gpurun_out/${tag}_bench_2gpu.json 2> gpurun_out/${tag}_bench_2gpu.err"
tail -c 1500 gpurun_out/${tag}_bench_2gpu.json
