# development helper: GPU tests (unless SKIP_TESTS=1) + quick A/B timing of the main build and any variants under lib/variants/
# ELIDE="0 1": NCR_ELIDE=0 (every flush writes the f64 canvas: comparable with round 1) / 1 (present-only flush, default)
# PREFETCH="0 1 auto": composite variant forced / chosen by the library
[ -n "$SKIP_TESTS" ] || python -m pytest tests -m gpu -x -q 2>&1 | tail -${TEST_TAIL:-3}
for v in main $(ls libnativecpurenderer_b200/lib/variants 2>/dev/null); do
  if [ $v = main ]; then unset NCR_LIBRARY; else export NCR_LIBRARY=$PWD/libnativecpurenderer_b200/lib/variants/$v/libNativeCPURenderer.so; fi
  for w in ${WORKLOADS:-c2 c4}; do for el in ${ELIDE:-1}; do for pf in ${PREFETCH:-auto}; do
    if [ $pf = auto ]; then unset NCR_PREFETCH; else export NCR_PREFETCH=$pf; fi
    NCR_ELIDE=$el python bench.py --steps ${STEPS:-30} --warmup 3 --workload $w --no-cpu-baseline --e2e-frames ${E2E:-4} ${BENCH_ARGS} 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v elide=$el prefetch=$pf', d['metric'], round(d['value'],1), 'fps | e2e', round(d['e2e']['value'],1), '| ms', {k:round(x,4) for k,x in d['kernel_ms'].items()}, '| parity', d['parity']['match'], d['parity']['outputs_checked'])"
  done; done; done
done
