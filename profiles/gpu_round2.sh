#!/bin/bash
# second call of the round: full parity suite on the current default kernels, resample kernel A/B, bench.   usage: bash profiles/gpu_round2.sh <tag>
tag=${1:-r2j}
mkdir -p gpurun_out
step() { local name=$1 to=$2; shift 2; local s=$(date +%s); timeout "$to" "$@"; echo "$name rc=$? $(( $(date +%s) - s ))s" | tee -a gpurun_out/${tag}_steps.log; }
step tests 420 bash -c "python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_gpu_tests.log 2>&1"
tail -5 gpurun_out/${tag}_gpu_tests.log
step resample_ab 240 bash -c "(python profiles/resample_ab.py; DSPB200_RS_MP2=0 python profiles/resample_ab.py) > gpurun_out/${tag}_resample_ab.jsonl 2> gpurun_out/${tag}_resample_ab.err"
cat gpurun_out/${tag}_resample_ab.jsonl
step ncu_resample 150 bash profiles/ncu_capture.sh resample resample_mp2_kernel ${tag}_resample
step bench 300 bash -c "python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err"
step workloads 240 bash -c "(for w in welch_real spectrogram resample filt_columns; do python bench.py --workload \$w --steps 20 --warmup 5; done; python bench.py --workload filt_columns --filt-alg td --steps 20 --warmup 5) > gpurun_out/${tag}_bench_workloads_1gpu.jsonl 2> gpurun_out/${tag}_bench_workloads.err"
