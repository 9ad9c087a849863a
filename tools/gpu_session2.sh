#!/bin/bash
# Round-2 GPU session pieces (run from the repo root on the box): bash tools/gpu_session2.sh <tag> <piece> ...
tag=$1; shift
out=gpurun_out
for piece in "$@"; do case $piece in
  tests)   python -m pytest tests -x -q -m gpu > $out/${tag}_pytest_gpu.txt 2>&1; tail -3 $out/${tag}_pytest_gpu.txt ;;
  bench)   python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_cfg2.json 2> $out/${tag}_bench_cfg2.err; tail -3 $out/${tag}_bench_cfg2.err; wc -c $out/${tag}_bench_cfg2.json
           python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_ref.json 2>/dev/null ;;
  benchw)  for w in cfg3 cfg4 cfg5; do python bench.py --workload $w --steps 20 --warmup 3 --no-extra > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.err; tail -2 $out/${tag}_bench_$w.err; done
           python - <<PY
import json
for w in ["cfg2","cfg3","cfg4","cfg5"]:
    try:
        d=json.load(open("$out/${tag}_bench_%s.json" % w))
        print(w, "ms/step %.3f" % d["ms_per_step"], "value %.3e" % d["value"], d["roofline"]["kernel_ms"], "frac %.3f" % d["roofline"]["frac"], "e2e %.3e" % d["e2e"]["value"], "resident %.3e" % d["e2e"]["resident_store"]["value"], "cpu %.3e" % d.get("cpu_baseline",{}).get("value",0))
    except Exception as e: print(w, "ERR", e)
PY
           ;;
  bench4)  python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 20 --warmup 5 > $out/${tag}_bench_4gpu.json 2> $out/${tag}_bench_4gpu.err; tail -3 $out/${tag}_bench_4gpu.err; wc -c $out/${tag}_bench_4gpu.json ;;
  bench8)  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 > $out/${tag}_bench_8gpu.json 2> $out/${tag}_bench_8gpu.err; tail -3 $out/${tag}_bench_8gpu.err; wc -c $out/${tag}_bench_8gpu.json ;;
  bench2)  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $out/${tag}_bench_2gpu.json 2> $out/${tag}_bench_2gpu.err; tail -5 $out/${tag}_bench_2gpu.err; wc -c $out/${tag}_bench_2gpu.json ;;
  chunks)  python tools/chunk_sweep.py 2>&1 | tee $out/${tag}_chunk_sweep.txt ;;
  pairs)   g++ -O2 -fopenmp tools/pair_threads.cpp -Iinclude -Lcollision_detect_gjk_epa_b200 -lgjkepa_b200 -Wl,-rpath,$PWD/collision_detect_gjk_epa_b200 -o build/pair_threads
           for t in 1 4 16 32 64; do build/pair_threads $t 20000 8; done 2>&1 | tee $out/${tag}_pair_threads.txt
           build/pair_threads 32 20000 32 2>&1 | tee -a $out/${tag}_pair_threads.txt ;;
  ab)      bash tools/ab_kernels.sh $AB_LIBS 2>&1 | tee $out/${tag}_ab.txt ;;
  cfg5big) export CHECK=0 LANES=0
           WL=cfg5 N=4000000 MODE=epa python tests/tools/quick_time.py 2>&1 | grep quick | tee $out/${tag}_cfg5big.txt
           WL=cfg5big N=4000000 MODE=epa python tests/tools/quick_time.py 2>&1 | grep quick | tee -a $out/${tag}_cfg5big.txt
           WL=cfg5big N=4000000 MODE=epa python tests/tools/quick_time.py > $out/${tag}_plain_cfg5big.log 2>&1 &&
           WL=cfg5big N=4000000 MODE=epa ncu --set full --clock-control none --import-source on -k "regex:epa3|gjk_" -s 9 -c 3 -f -o $out/${tag}_prof_cfg5big python tests/tools/quick_time.py > $out/${tag}_ncu_d.log 2>&1
           python tools/ncu_summary.py $out/${tag}_prof_cfg5big.ncu-rep > $out/${tag}_prof_cfg5big.txt 2>&1
           for k in gjk_init_kernel gjk_walk_kernel epa3_kernel; do echo "=== top source lines by stall samples: $k" >> $out/${tag}_prof_cfg5big.txt; python tools/ncu_source_top.py $out/${tag}_prof_cfg5big.ncu-rep $k 14 --by-samples >> $out/${tag}_prof_cfg5big.txt 2>&1; done
           rm -f $out/${tag}_prof_cfg5big.ncu-rep ;;
  resident) python tools/resident_kernel_ms.py 2>&1 | tee $out/${tag}_resident_kernel_ms.txt ;;
  ncu)     # launch list of the bench command, then full captures of the hot kernels (each only after the same command ran
           # clean); the reports are summarised HERE (raw metrics, opcode mix, top source lines) because gpurun brings back
           # at most 64 MiB, and only the small ones travel
           python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > $out/${tag}_plain_bench.log 2>&1 &&
           ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > $out/${tag}_ncu_launch.log 2>&1
           export CHECK=0 LANES=0
           summarise() {   # <rep> <out.txt> <kernel regex> ...
               rep=$1; txt=$2; shift 2
               python tools/ncu_summary.py $rep > $txt 2>&1
               for k in "$@"; do echo "=== opcode mix: $k" >> $txt; python tools/ncu_opmix.py $rep "$k" 25 >> $txt 2>&1
                                 echo "=== top source lines by stall samples: $k" >> $txt; python tools/ncu_source_top.py $rep "$k" 30 --by-samples >> $txt 2>&1; done
           }
           WL=cfg2 MODE=gjk python tests/tools/quick_time.py > $out/${tag}_plain_cfg2.log 2>&1 &&
           WL=cfg2 MODE=gjk ncu --set full --clock-control none --import-source on -k regex:gjk_ -s 8 -c 2 -f -o $out/${tag}_prof_gjk_cfg2 python tests/tools/quick_time.py > $out/${tag}_ncu_a.log 2>&1
           summarise $out/${tag}_prof_gjk_cfg2.ncu-rep $out/${tag}_prof_gjk_cfg2.txt gjk_init_kernel gjk_walk_kernel
           WL=cfg4 MODE=epa python tests/tools/quick_time.py > $out/${tag}_plain_cfg4.log 2>&1 &&
           WL=cfg4 MODE=epa ncu --set full --clock-control none --import-source on -k "regex:epa|route|gjk_" -s 20 -c 10 -f -o $out/${tag}_prof_cfg4 python tests/tools/quick_time.py > $out/${tag}_ncu_b.log 2>&1
           summarise $out/${tag}_prof_cfg4.ncu-rep $out/${tag}_prof_cfg4.txt "epa3_kernel.*64EE" "epa3_kernel.*128EE" "epa_kernel"
           rm -f $out/${tag}_prof_cfg4.ncu-rep
           WL=cfg5 MODE=epa python tests/tools/quick_time.py > $out/${tag}_plain_cfg5.log 2>&1 &&
           WL=cfg5 MODE=epa ncu --set full --clock-control none --import-source on -k "regex:epa3|gjk_|hull_prep" -s 12 -c 4 -f -o $out/${tag}_prof_cfg5 python tests/tools/quick_time.py > $out/${tag}_ncu_c.log 2>&1
           summarise $out/${tag}_prof_cfg5.ncu-rep $out/${tag}_prof_cfg5.txt gjk_init_kernel gjk_walk_kernel epa3_kernel
           rm -f $out/${tag}_prof_cfg5.ncu-rep
           ls -la $out/${tag}_prof_*.ncu-rep ;;
esac; done
