#!/bin/bash
# third call of the round: parity suite on the register-tiled FIR kernel + 128-bit tap loads in the resample kernel, A/B of both, ncu of the FIR kernel, bench.
# usage: bash profiles/gpu_round3.sh <tag>
tag=${1:-r2k}
mkdir -p gpurun_out
step() { local name=$1 to=$2; shift 2; local s=$(date +%s); timeout "$to" "$@"; echo "$name rc=$? $(( $(date +%s) - s ))s" | tee -a gpurun_out/${tag}_steps.log; }
step tests 420 bash -c "python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_gpu_tests.log 2>&1"
tail -5 gpurun_out/${tag}_gpu_tests.log
step fir_ab 180 bash -c "(python profiles/fir_ab.py; DSPB200_FIR_TILE=0 python profiles/fir_ab.py) > gpurun_out/${tag}_fir_ab.jsonl 2> gpurun_out/${tag}_fir_ab.err"
cat gpurun_out/${tag}_fir_ab.jsonl
step resample_ab 180 bash -c "(python profiles/resample_ab.py; DSPB200_LIB=\$PWD/dsp.jl_b200/libdspb200_rshq0.so python profiles/resample_ab.py) > gpurun_out/${tag}_resample_ab.jsonl 2> gpurun_out/${tag}_resample_ab.err"
cat gpurun_out/${tag}_resample_ab.jsonl
step ncu_fir 120 bash profiles/ncu_capture.sh fir fir_tile_kernel ${tag}_fir
step bench 300 bash -c "python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err"
step workloads 240 bash -c "(for w in welch_real spectrogram resample filt_columns; do python bench.py --workload \$w --steps 20 --warmup 5; done; python bench.py --workload filt_columns --filt-alg td --steps 20 --warmup 5) > gpurun_out/${tag}_bench_workloads_1gpu.jsonl 2> gpurun_out/${tag}_bench_workloads.err"
