#!/bin/bash
# end-of-round validation on one GPU: full GPU test suite, smoke, both bench arms, launch list of the bench command
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/final_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_reference_line_c.json 2> gpurun_out/bench_ref.err; tail -c 400 gpurun_out/r2_bench_reference_line_c.json
timeout 600 python bench.py > gpurun_out/r2_bench_line_c.json 2> gpurun_out/bench_c.err; tail -c 300 gpurun_out/bench_c.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_line_c.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'prefill', d['detail']['prefill_ms'], 'dec', d['detail']['decode_ms_per_token'], 'train', d['detail'].get('train',{}).get('ms_per_step'), 'beam5', d['detail'].get('beam5'))
PY
if [ "${SKIP_NCU:-0}" != "1" ]; then
MERLIN_BENCH_PROFILE=1 MERLIN_BENCH_TRAIN=0 MERLIN_BENCH_CPU=0 MERLIN_BENCH_TORCH_GPU=0 MERLIN_BENCH_BEAM=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2b_bench_launches.csv python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_bench.log 2>&1
fi
[ "${SKIP_NCU:-0}" != "1" ] && tail -2 gpurun_out/ncu_bench.log | cut -c1-300 || true
