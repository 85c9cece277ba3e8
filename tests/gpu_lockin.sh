# One GPU call that refreshes everything under profiles/ (run through gpurun from the repo root).
set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --steps 40 --warmup 5 > gpurun_out/r1_bench_c2.json 2> gpurun_out/r1_bench_c2.err; tail -c 300 gpurun_out/r1_bench_c2.json
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python tests/gpu_ab.py > gpurun_out/r1_ab.jsonl 2>&1; cat gpurun_out/r1_ab.jsonl
timeout 300 python bench.py --workload c3 --steps 20 --warmup 4 > gpurun_out/r1_bench_c3.json 2>/dev/null; cut -c1-150 gpurun_out/r1_bench_c3.json
timeout 300 python bench.py --workload c5 --steps 10 --warmup 3 > gpurun_out/r1_bench_c5.json 2>/dev/null; cut -c1-150 gpurun_out/r1_bench_c5.json
timeout 200 python bench.py --workload c2hash --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r1_bench_c2hash.json 2>/dev/null; cut -c1-150 gpurun_out/r1_bench_c2hash.json
timeout 100 python tests/gpu_hash_bench.py 2>&1 | tail -1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r1_launches_bench.log 2>&1; tail -1 gpurun_out/r1_launches.csv | cut -c1-120
MB_N=8192 timeout 500 ncu --set full --import-source on --clock-control none -k regex:"mlp_(fwd|dgrad|wgrad_kernel)" --launch-skip 4 -c 4 -o gpurun_out/r1_prof -f python tests/gpu_profile_target.py 2>&1 | tail -2
