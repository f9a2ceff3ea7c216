# development helper: the measurement pass of round 2 (default bench line, launch list, ncu captures) -> gpurun_out/
set -x
R=r02
python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench_default.json 2> gpurun_out/${R}_bench_default.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --only-main --e2e-frames 2 --e2e-threads 1 > gpurun_out/${R}_ncu_l.log 2>&1
for w in c2 c4 bg; do ncu --set full --clock-control none --import-source on -k regex:ncr_composite -c 1 -s 6 -o gpurun_out/${R}_composite_$w -f python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --only-main --e2e-frames 2 --e2e-threads 1 > gpurun_out/${R}_ncu_$w.log 2>&1; done
ncu --set full --clock-control none --import-source on -k regex:ncr_bin_fine -c 1 -s 6 -o gpurun_out/${R}_bin_fine_c2 -f python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --only-main --e2e-frames 2 --e2e-threads 1 > gpurun_out/${R}_ncu_fine.log 2>&1
ncu --set full --clock-control none -k regex:ncr_yuv420p -c 1 -s 2 -o gpurun_out/${R}_yuv_c5 -f python bench.py --workload c5 --present yuv420p --steps 3 --warmup 3 --no-cpu-baseline --only-main --e2e-frames 2 --e2e-threads 1 > gpurun_out/${R}_ncu_yuv.log 2>&1
ls -la gpurun_out | tail -12
