#!/bin/bash
# final record of the round: parity suite, smoke, bench line + launch list, kernel timings, workload lines, ncu summary of the resample kernel.
tag=${1:-r2p}
mkdir -p gpurun_out
step() { local name=$1 to=$2; shift 2; local s=$(date +%s); timeout "$to" "$@"; echo "$name rc=$? $(( $(date +%s) - s ))s" | tee -a gpurun_out/${tag}_steps.log; }
step tests 300 bash -c "python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_gpu_tests.log 2>&1"
tail -3 gpurun_out/${tag}_gpu_tests.log
step smoke 60 bash -c "python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/${tag}_smoke.log 2>&1"
step bench 120 bash -c "python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err"
step launches 90 bash -c "ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extra --no-check > gpurun_out/${tag}_bench_under_ncu.log 2>&1"
step time_kernels 60 bash -c "python profiles/time_kernels.py 10 > gpurun_out/${tag}_time.jsonl 2> gpurun_out/${tag}_time.err"
step workloads 120 bash -c "(for w in welch_real spectrogram resample filt_columns; do python bench.py --workload \$w --steps 20 --warmup 5; done; python bench.py --workload filt_columns --filt-alg td --steps 20 --warmup 5) > gpurun_out/${tag}_bench_workloads_1gpu.jsonl 2> gpurun_out/${tag}_bench_workloads.err"
step ncu_resample 60 bash profiles/ncu_capture.sh resample resample_mp2_kernel ${tag}_resample
