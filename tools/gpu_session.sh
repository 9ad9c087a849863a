#!/bin/bash
# One GPU-box session: parity tests, the bench lines, the ncu launch list and full captures of the top kernels.
# Usage (from the repo root on the box): bash tools/gpu_session.sh <tag>
tag=${1:-r1}
out=gpurun_out
python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1; tail -2 $out/pytest_gpu_$tag.log
python bench.py > $out/bench_${tag}_cfg2.json 2> $out/bench_${tag}_cfg2.err
for w in cfg3 cfg4 cfg5; do python bench.py --workload $w --steps 20 --warmup 3 --no-extra > $out/bench_${tag}_$w.json 2> $out/bench_${tag}_$w.err; done
python bench.py --impl reference --steps 5 --warmup 1 > $out/bench_${tag}_ref.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_${tag}_cfg2.csv python bench.py --steps 2 --warmup 1 --no-cpu > $out/ncu_launch_$tag.log 2>&1
export CHECK=0 LANES=0
WL=cfg2 MODE=gjk ncu --set full --clock-control none --import-source on -k regex:gjk_ -s 8 -c 2 -f -o $out/prof_${tag}_gjk_cfg2 python tests/tools/quick_time.py > $out/ncu_a_$tag.log 2>&1
WL=cfg3 MODE=epa ncu --set full --clock-control none --import-source on -k "regex:gjk_.*group|epa3" -s 12 -c 3 -f -o $out/prof_${tag}_cfg3 python tests/tools/quick_time.py > $out/ncu_b_$tag.log 2>&1
SEEDS=${SOAK_SEEDS:-10} python tests/tools/soak.py > $out/soak_$tag.log 2>&1; tail -1 $out/soak_$tag.log
python - <<PY
import json
for w in ["cfg2","cfg3","cfg4","cfg5"]:
    try:
        d=json.load(open("$out/bench_${tag}_%s.json" % w))
        print(w, "ms/step %.3f" % d["ms_per_step"], "value %.3e" % d["value"], "contacts/s %.3e" % d["contacts_per_s"], d["roofline"]["kernel_ms"], "frac %.3f" % d["roofline"]["frac"], "e2e %.3e" % d["e2e"]["value"], "cpu %.3e" % d["cpu_baseline"]["value"], d.get("extra",{}).get("contacts_per_s"))
    except Exception as e: print(w, "ERR", e)
PY
