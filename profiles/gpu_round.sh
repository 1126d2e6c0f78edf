#!/bin/bash
# One gpurun call of the round: parity suite, smoke, bench line, launch list, kernel timings, ncu summaries of the kernels
# that had none, strong-scaling probe, widened rows.  Every step has its own timeout and log under gpurun_out/; later
# (optional) steps are skipped once the call has used its time budget.   usage: bash profiles/gpu_round.sh <tag> [budget_s]
tag=${1:-r2x}; budget=${2:-780}
mkdir -p gpurun_out
t0=$(date +%s)
left() { echo $(( budget - ($(date +%s) - t0) )); }
step() { # name timeout cmd...
  local name=$1 to=$2; shift 2
  if [ "$(left)" -lt 20 ]; then echo "skip $name (budget)" | tee -a gpurun_out/${tag}_steps.log; return; fi
  local s=$(date +%s)
  timeout "$to" "$@"
  echo "$name rc=$? $(( $(date +%s) - s ))s" | tee -a gpurun_out/${tag}_steps.log
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
step tests 420 bash -c "python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_gpu_tests.log 2>&1"
tail -5 gpurun_out/${tag}_gpu_tests.log
step smoke 120 bash -c "python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/${tag}_smoke.log 2>&1"
step bench 300 bash -c "python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err"
tail -c 600 gpurun_out/${tag}_bench_1gpu.json
step launches 180 bash -c "ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extra --no-check > gpurun_out/${tag}_bench_under_ncu.log 2>&1"
step time_kernels 180 bash -c "python profiles/time_kernels.py 10 > gpurun_out/${tag}_time.jsonl 2> gpurun_out/${tag}_time.err"
step workloads 240 bash -c "(for w in welch_real spectrogram resample filt_columns; do python bench.py --workload \$w --steps 20 --warmup 5; done; python bench.py --workload filt_columns --filt-alg td --steps 20 --warmup 5) > gpurun_out/${tag}_bench_workloads_1gpu.jsonl 2> gpurun_out/${tag}_bench_workloads.err"
step win_ab 240 bash -c "python profiles/win_exact_ab.py > gpurun_out/${tag}_win_ab.jsonl 2> gpurun_out/${tag}_win_ab.err; DSPB200_LIB=\$PWD/dsp.jl_b200/libdspb200_winexact.so python profiles/win_exact_ab.py >> gpurun_out/${tag}_win_ab.jsonl 2>> gpurun_out/${tag}_win_ab.err"
step ncu_resample 150 bash profiles/ncu_capture.sh resample resample_mp_kernel ${tag}_resample
step ncu_fir 120 bash profiles/ncu_capture.sh fir fir_td_kernel ${tag}_fir
step ncu_welch_r 150 bash profiles/ncu_capture.sh welch_r welch_fused_kernel ${tag}_welch_r
step strong_probe 150 bash -c "python profiles/strong_probe.py > gpurun_out/${tag}_strong_probe.jsonl 2> gpurun_out/${tag}_strong_probe.err"
step widened 240 bash -c "python profiles/time_widened.py > gpurun_out/${tag}_time_widened.jsonl 2> gpurun_out/${tag}_time_widened.err"
step bench_ref 240 bash -c "python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err"
cat gpurun_out/${tag}_steps.log
