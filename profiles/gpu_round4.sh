#!/bin/bash
# parity suite on the DSP_RS_V3 resample kernel, A/B against the previous build (libdspb200_rsv0.so), ncu summary, timings.   usage: bash profiles/gpu_round4.sh <tag>
tag=${1:-r2n}
mkdir -p gpurun_out
step() { local name=$1 to=$2; shift 2; local s=$(date +%s); timeout "$to" "$@"; echo "$name rc=$? $(( $(date +%s) - s ))s" | tee -a gpurun_out/${tag}_steps.log; }
step resample_ab 180 bash -c "(python profiles/resample_ab.py; DSPB200_LIB=\$PWD/dsp.jl_b200/libdspb200_rsv0.so python profiles/resample_ab.py) > gpurun_out/${tag}_resample_ab.jsonl 2> gpurun_out/${tag}_resample_ab.err"
cat gpurun_out/${tag}_resample_ab.jsonl
step tests 420 bash -c "python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_gpu_tests.log 2>&1"
tail -5 gpurun_out/${tag}_gpu_tests.log
step ncu_resample 150 bash profiles/ncu_capture.sh resample resample_mp2_kernel ${tag}_resample
step time_kernels 180 bash -c "python profiles/time_kernels.py 10 > gpurun_out/${tag}_time.jsonl 2> gpurun_out/${tag}_time.err"
step workloads 240 bash -c "(for w in resample; do python bench.py --workload \$w --steps 20 --warmup 5; done) > gpurun_out/${tag}_bench_workloads_1gpu.jsonl 2> gpurun_out/${tag}_bench_workloads.err"
