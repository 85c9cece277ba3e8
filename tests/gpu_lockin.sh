#!/bin/bash
# One command that regenerates the round's GPU evidence under gpurun_out/ (run on the GPU box through gpurun):
#   tests + bench + launch list + ncu --set full of every own kernel at the C2 size.
# usage: bash tests/gpu_lockin.sh <tag>
tag=${1:-r2}
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/${tag}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_pytest.log)
tail -3 gpurun_out/${tag}_pytest.log
(timeout 300 python __graft_entry__.py --smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_smoke.log); tail -2 gpurun_out/${tag}_smoke.log
(timeout 900 python bench.py --steps 40 --warmup 5 > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench_c2.err)
(timeout 300 python tests/gpu_ab.py > gpurun_out/${tag}_ab.jsonl 2>&1)
(timeout 300 python bench.py --workload c5 --steps 8 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_bench_c5.json 2> gpurun_out/${tag}_bench_c5.err)
(timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.err)
(timeout 300 python bench.py --workload c1 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/${tag}_bench_c1.json 2> gpurun_out/${tag}_bench_c1.err)
# memcheck + racecheck over the kernels added last (deep hash heads, pose kernels) and the sky / occupancy paths
(timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_hashgrid.py tests/test_gpu_round2.py -q \
   -k "deep_heads or pose or sky or occupancy" > gpurun_out/${tag}_sanitizer.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/${tag}_sanitizer.log
 timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_hashgrid.py -q \
   -k "deep_heads and odd" >> gpurun_out/${tag}_sanitizer.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/${tag}_sanitizer.log
 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q \
   -k "mlp_backward and 128-2-640 or mlp_forward_layers and 256-4-1000" >> gpurun_out/${tag}_sanitizer.log 2>&1; echo "memcheck (tcgen05 kernels) rc=$?" >> gpurun_out/${tag}_sanitizer.log)
grep -E "rc=|ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/${tag}_sanitizer.log | tail -8
(timeout 200 python tests/gpu_probe_l2.py > gpurun_out/${tag}_probe_l2.json 2>&1; timeout 200 python tests/gpu_probe_store.py > gpurun_out/${tag}_probe_bulk_store.txt 2>&1; timeout 200 python tests/gpu_hash_hotspot.py > gpurun_out/${tag}_hash_hotspot.json 2>&1)
# launch list of a short bench run (every launch with its device time; shares, not absolutes)
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/${tag}_launches_raw.csv \
   python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_launches_bench.log 2>&1)
# full counters + source of every own kernel (one joint iteration, one map-only iteration, one render, one hash iteration), C2 size
(MB_N=8192 timeout 1500 ncu --set full --clock-control none --profile-from-start off \
   -k regex:'^(adam|sgd|mlp_|ogm_|pack|ray_|render|sample_|wgrad_|hash_|points_|loss_|pose_)' \
   -o gpurun_out/${tag}_prof python tests/gpu_profile_target.py > gpurun_out/${tag}_ncu.log 2>&1)
ncu -i gpurun_out/${tag}_prof.ncu-rep --page raw --csv 2>/dev/null | python profiles/summarize_ncu.py > gpurun_out/${tag}_ncu_summary.csv
# gpurun merges at most 64 MiB back: the compact summary is what gets committed, the raw report only if it is small
[ $(stat -c %s gpurun_out/${tag}_prof.ncu-rep) -gt 40000000 ] && rm -f gpurun_out/${tag}_prof.ncu-rep
cuobjdump -sass loner_b200/libloner_b200.so | grep -oE "UTCHMMA[.A-Z0-9]*|UTCBAR[.A-Z0-9]*|LDTM[.A-Z0-9x]*|UBLKCP[.A-Z0-9]*|UCGABAR[_A-Z]*|F2FP[.A-Z0-9_]*|HSET2[.A-Z0-9]*" | sort | uniq -c > gpurun_out/${tag}_sass_mnemonics.txt
ls -la gpurun_out | tail -20
