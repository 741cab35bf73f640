mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/bench_pipeline.py 1000000 2 2>&1 | tail -1 | tee gpurun_out/r02_pipeline.json
timeout 300 python bench.py --workload dimer --primers 100000 --steps 1 --warmup 0 2>/dev/null | tail -1 | tee gpurun_out/r02_dimer_1gpu.json
for kk in k_prefilter k_hist k_cscan; do timeout 300 ncu --set full --clock-control none --import-source on -k regex:^${kk}\$ -s 2 -c 1 -o gpurun_out/r02e_$kk python tools/profile_calls.py 1000000 1 > gpurun_out/ncu_$kk.log 2>&1; tail -1 gpurun_out/ncu_$kk.log; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02e_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02e_bench_under_ncu.log 2>&1
ls -la gpurun_out | tail -8
