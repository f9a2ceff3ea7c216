# development helper: the full measurement pass of a round (bench on every workload, launch list, ncu captures) -> gpurun_out/
set -x
python bench.py > gpurun_out/s2_bench_default.json 2> gpurun_out/s2_bench_default.err
for w in c1 c3 c4 c5 bg c2x c3p; do python bench.py --workload $w --steps 30 --warmup 3 > gpurun_out/s2_bench_$w.json 2> gpurun_out/s2_bench_$w.err; done
python bench.py --workload c4 --steps 30 --warmup 3 --no-cpu-baseline --video 1200 --present yuv420p > gpurun_out/s2_bench_c4_video_yuv.json 2>/dev/null
python bench.py --workload c5 --steps 30 --warmup 3 --no-cpu-baseline --video 600 > gpurun_out/s2_bench_c5_video.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01s2_launches_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-frames 2 --e2e-threads 1 > gpurun_out/s2_ncu_l.log 2>&1
for w in c2 bg; do ncu --set full --clock-control none --import-source on -k regex:ncr_composite -c 1 -s 4 -o gpurun_out/r01s2_composite_$w -f python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --e2e-frames 2 --e2e-threads 1 > gpurun_out/s2_ncu_$w.log 2>&1; done
ncu --set full --clock-control none -k regex:ncr_yuv420p -c 1 -o gpurun_out/r01s2_yuv_c4 -f python bench.py --workload c4 --present yuv420p --steps 3 --warmup 3 --no-cpu-baseline --e2e-frames 2 --e2e-threads 1 > gpurun_out/s2_ncu_yuv.log 2>&1
ls -la gpurun_out | tail -20
