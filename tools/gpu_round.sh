#!/bin/bash
# One GPU-box session: GPU test suite, smoke, benches, then the two ncu passes (launch list + full capture).
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee gpurun_out/smoke.log
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 3000 gpurun_out/bench_c2.json
python bench.py --workload headline --steps 10 > gpurun_out/bench_headline.json 2> gpurun_out/bench_headline.err
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
if [ "${1:-}" = "ncu" ]; then
  python bench.py --steps 2 --warmup 1 > gpurun_out/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 1 > gpurun_out/ncu1.log 2>&1
  python bench.py --steps 2 --warmup 1 > gpurun_out/plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:fa_ -s 8 -c 4 -f -o gpurun_out/prof \
      python bench.py --steps 2 --warmup 1 > gpurun_out/ncu2.log 2>&1
  tail -3 gpurun_out/ncu1.log gpurun_out/ncu2.log
fi
