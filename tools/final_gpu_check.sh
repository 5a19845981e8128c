#!/bin/bash
# Round-end confirmation on one B200 (run through gpurun): GPU parity suite, smoke, the bench lines of every BASELINE
# config that fits one GPU, the ncu launch list of the headline step and one full capture of the forward-backward kernel.
# Everything lands in gpurun_out/ (scratch); copy what should be judged into profiles/.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee $O/pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee $O/smoke.txt
timeout 400 python -u bench.py 2>$O/bench_headline.err | tail -1 > $O/bench_headline.json
timeout 300 python -u bench.py --impl reference --steps 3 --warmup 1 2>$O/bench_ref.err | tail -1 > $O/bench_reference_arm.json
for w in c2 c3 c4; do
  timeout 200 python -u bench.py --workload $w --no-e2e --no-cpu-baseline --steps 5 --warmup 3 2>/dev/null | tail -1 > $O/bench_$w.json
done
timeout 200 python -u bench.py --front xvectors --no-cpu-baseline --steps 5 --warmup 3 2>/dev/null | tail -1 > $O/bench_front_xvectors.json
timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/ncu_launches.log 2>&1
timeout 250 ncu --set full --clock-control none --import-source on -k regex:forward_backward_la -s 12 -c 1 -f -o $O/fb_la \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/ncu_fb.log 2>&1
for f in bench_headline bench_reference_arm bench_c2 bench_c3 bench_c4 bench_front_xvectors; do
  python - <<PY
import json
try:
    d = json.load(open("$O/$f.json"))
    k = d.get("kernels", {})
    print("$f", round(d.get("ms_per_step", 0), 3), "%.4g" % d.get("value", 0), {n: round(v["ms_per_launch"], 4) for n, v in k.items()}, "e2e", (d.get("e2e") or {}).get("value"))
except Exception as e:
    print("$f", "unreadable:", e)
PY
done
