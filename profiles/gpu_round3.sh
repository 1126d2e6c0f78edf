#!/bin/bash
# One gpurun call: parity suite, smoke, FIR kernel A/B (register-tiled vs plain), bench line + launch list, kernel timings, workloads,
# ncu summary of the FIR kernel, strong-scaling probe.   usage: bash profiles/gpu_round3.sh <tag>
tag=${1:-r2l}
mkdir -p gpurun_out
step() { local name=$1 to=$2; shift 2; local s=$(date +%s); timeout "$to" "$@"; echo "$name rc=$? $(( $(date +%s) - s ))s" | tee -a gpurun_out/${tag}_steps.log; }
step tests 420 bash -c "python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_gpu_tests.log 2>&1"
tail -5 gpurun_out/${tag}_gpu_tests.log
step smoke 120 bash -c "python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/${tag}_smoke.log 2>&1"
step fir_ab 180 bash -c "(python profiles/fir_ab.py; DSPB200_FIR_TILE=0 python profiles/fir_ab.py) > gpurun_out/${tag}_fir_ab.jsonl 2> gpurun_out/${tag}_fir_ab.err"
cat gpurun_out/${tag}_fir_ab.jsonl
step bench 300 bash -c "python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err"
tail -c 400 gpurun_out/${tag}_bench_1gpu.json
step launches 180 bash -c "ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extra --no-check > gpurun_out/${tag}_bench_under_ncu.log 2>&1"
step time_kernels 180 bash -c "python profiles/time_kernels.py 10 > gpurun_out/${tag}_time.jsonl 2> gpurun_out/${tag}_time.err"
step workloads 240 bash -c "(for w in welch_real spectrogram resample filt_columns; do python bench.py --workload \$w --steps 20 --warmup 5; done; python bench.py --workload filt_columns --filt-alg td --steps 20 --warmup 5) > gpurun_out/${tag}_bench_workloads_1gpu.jsonl 2> gpurun_out/${tag}_bench_workloads.err"
step ncu_fir 120 bash profiles/ncu_capture.sh fir fir_tile_kernel ${tag}_fir
step strong_probe 150 bash -c "python profiles/strong_probe.py > gpurun_out/${tag}_strong_probe.jsonl 2> gpurun_out/${tag}_strong_probe.err"
