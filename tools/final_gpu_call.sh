set -x
python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1; echo smoke rc=$?; tail -3 gpurun_out/r2b_smoke.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; echo bench rc=$?; tail -c 1500 gpurun_out/r2b_bench_n1.json
python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2b_bench_plain.json && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2b_ncu_l.log 2>&1; echo launchlist rc=$?
python tools/profile_step.py thai2_1080p bvh 12 | cut -c1-300 && ncu --set full --clock-control none --import-source on -k regex:trace_shade_persistent -s 11 -c 1 -f -o gpurun_out/r2b_prof_default python tools/profile_step.py thai2_1080p bvh 12 > gpurun_out/r2b_ncu_step.log 2>&1; echo ncu rc=$?
