#!/bin/bash
# One GPU session: parity tests, smoke, both bench arms, ncu launch list of the bench command, and ncu --set full
# captures of the three bench kernels at bench shard size.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
R=${ROUND:-r2}
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_$R.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$R.log
tail -n 3 gpurun_out/pytest_gpu_$R.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$R.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke_$R.log
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$R.json 2> gpurun_out/bench_ref_$R.err; echo "ref rc=$?"
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; rc=$?; echo "bench rc=$rc"
if [ $rc -eq 0 ]; then
  timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --headline-only > gpurun_out/plain_headline.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"window_|spectral_" -c 80 --csv \
      --log-file gpurun_out/launches_$R.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --headline-only > gpurun_out/ncu_launches_$R.log 2>&1
  echo "ncu launches rc=$?"
fi
NSUB=125 python tools/ncu_target.py c3_full > gpurun_out/plain_full.log 2>&1 && \
  NSUB=125 ncu --set full --clock-control none --import-source on -k regex:window_stats -s 1 -c 1 -f -o gpurun_out/prof_stats_$R python tools/ncu_target.py c3_full > gpurun_out/ncu_stats_$R.log 2>&1
NSUB=125 python tools/ncu_target.py c3_spec > gpurun_out/plain_spec.log 2>&1 && \
  NSUB=125 ncu --set full --clock-control none --import-source on -k regex:spectral_ -s 1 -c 1 -f -o gpurun_out/prof_spec_$R python tools/ncu_target.py c3_spec > gpurun_out/ncu_spec_$R.log 2>&1
NSUB=125 python tools/ncu_order_target.py c3 > gpurun_out/plain_order.log 2>&1 && \
  NSUB=125 ncu --set full --clock-control none --import-source on -k regex:window_order -s 1 -c 1 -f -o gpurun_out/prof_order_$R python tools/ncu_order_target.py c3 > gpurun_out/ncu_order_$R.log 2>&1
echo done
