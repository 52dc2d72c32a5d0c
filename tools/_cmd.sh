BENCH_ARGS=" " bash tools/gpu_iter2.sh
for sl in 2 4 5; do python bench.py --steps 100 --warmup 10 --slots $sl --batches 6 --skip-e2e --skip-cpu-baseline --skip-configs --no-kernel-timing 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('slots $sl', d['value'], d['ms_per_step'], d['batch_latency_ms']['mean'])"; done
python bench.py --stream 40 --warmup 2 2> gpurun_out/stream.err | tee gpurun_out/stream40.json | cut -c1-900; tail -3 gpurun_out/stream.err
