#!/bin/bash
# Round-end confirmation on one B200 (run through gpurun): GPU parity suite, smoke, the bench lines of every BASELINE
# config that fits one GPU, the reference arm, the ncu launch list of the headline step and full captures of the dominant
# kernels.  Everything lands in gpurun_out/ (scratch); copy what should be judged into profiles/.
set -u
mkdir -p gpurun_out
O=gpurun_out
P=${1:-r2}
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee $O/${P}_pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee $O/${P}_smoke.txt
timeout 400 python -u bench.py 2>$O/${P}_bench_headline.err | tail -1 > $O/${P}_bench_headline.json
timeout 300 python -u bench.py --impl reference --steps 3 --warmup 1 2>$O/${P}_bench_ref.err | tail -1 > $O/${P}_bench_reference_arm.json
timeout 200 python -u bench.py --workload c1 --steps 10 2>/dev/null | tail -1 > $O/${P}_bench_c1.json
timeout 200 python -u bench.py --workload c1 --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > $O/${P}_bench_c1_reference_arm.json
for w in c2 c3 c4 c4share c5s4 c5s8 c5s16 c5s32 c5s64; do
  timeout 300 python -u bench.py --workload $w --no-e2e --no-cpu-baseline --steps 5 --warmup 3 2>/dev/null | tail -1 > $O/${P}_bench_$w.json
done
timeout 200 python -u bench.py --front xvectors --no-cpu-baseline --no-e2e --steps 5 --warmup 3 2>/dev/null | tail -1 > $O/${P}_bench_front_xvectors.json
timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${P}_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --parts 1 > $O/${P}_ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'loglik_mma|mstep_mma|forward_backward_la|project_tcgen05' -s 8 -c 4 -f -o $O/${P}_headline_kernels \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --parts 1 > $O/${P}_ncu_full.log 2>&1
for f in bench_headline bench_reference_arm bench_c1 bench_c1_reference_arm bench_c2 bench_c3 bench_c4 bench_c4share bench_c5s4 bench_c5s8 bench_c5s16 bench_c5s32 bench_c5s64 bench_front_xvectors; do
  python - <<PY
import json
try:
    d = json.load(open("$O/${P}_$f.json"))
    k = d.get("kernels", {})
    print("$f", round(d.get("ms_per_step", 0), 3), "%.4g" % d.get("value", 0), {n: round(v.get("ms_per_launch", v.get("ms_per_step", 0)), 4) for n, v in k.items()}, "e2e", (d.get("e2e") or {}).get("value"), "parity", (d.get("parity") or {}).get("ok"))
except Exception as e:
    print("$f", "unreadable:", e)
PY
done
