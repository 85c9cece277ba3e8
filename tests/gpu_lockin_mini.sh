#!/bin/bash
# Short form of tests/gpu_lockin.sh for a tight GPU budget: tests, smoke, A/B of the kernel variants, the C2 bench line
# and `ncu --set full` of the tcgen05 kernels only.   usage: bash tests/gpu_lockin_mini.sh <tag> [extra]
tag=${1:-r2f}
mkdir -p gpurun_out
(timeout 300 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_pytest.log)
grep -E "^FAILED|^ERROR|passed|failed|rc=" gpurun_out/${tag}_pytest.log | tail -12
(timeout 150 python tests/gpu_ab.py > gpurun_out/${tag}_ab.jsonl 2>&1); cat gpurun_out/${tag}_ab.jsonl
(timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench_c2.err)
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_bench_c2.json"))
    print(d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], {k: v["ms"] for k, v in d["roofline"]["kernels"].items()}, d["roofline"]["step"], d.get("e2e", {}).get("value"))
except Exception as ex:
    print("bench line unreadable:", ex)
PY
(timeout 120 python __graft_entry__.py --smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_smoke.log); tail -2 gpurun_out/${tag}_smoke.log
(MB_N=8192 timeout ${NCU_T:-400} ncu --set full --clock-control none --profile-from-start off \
   -k regex:"${NCU_K:-^(mlp_|wgrad_)}" -o gpurun_out/${tag}_prof python tests/gpu_profile_target.py > gpurun_out/${tag}_ncu.log 2>&1)
ncu -i gpurun_out/${tag}_prof.ncu-rep --page raw --csv 2>/dev/null | python profiles/summarize_ncu.py > gpurun_out/${tag}_ncu_summary.csv
[ -f gpurun_out/${tag}_prof.ncu-rep ] && [ $(stat -c %s gpurun_out/${tag}_prof.ncu-rep) -gt 30000000 ] && rm -f gpurun_out/${tag}_prof.ncu-rep
cut -d, -f1-8 gpurun_out/${tag}_ncu_summary.csv | head -12
if [ "$2" = "extra" ]; then
  (timeout 200 python bench.py --workload c5 --steps 8 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_bench_c5.json 2> gpurun_out/${tag}_bench_c5.err)
  (timeout 200 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.err)
  (timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/${tag}_launches_raw.csv \
     python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_launches_bench.log 2>&1)
  cuobjdump -sass loner_b200/libloner_b200.so | grep -oE "UTCHMMA[.A-Z0-9]*|UTCBAR[.A-Z0-9]*|LDTM[.A-Z0-9x]*|UBLKCP[.A-Z0-9]*|UCGABAR[_A-Z]*" | sort | uniq -c > gpurun_out/${tag}_sass_mnemonics.txt
fi
