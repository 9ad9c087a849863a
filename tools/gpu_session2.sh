#!/bin/bash
# Round-2 GPU session pieces (run from the repo root on the box): bash tools/gpu_session2.sh <tag> <piece> ...
tag=$1; shift
out=gpurun_out
for piece in "$@"; do case $piece in
  tests)   python -m pytest tests -x -q -m gpu > $out/${tag}_pytest_gpu.txt 2>&1; tail -3 $out/${tag}_pytest_gpu.txt ;;
  bench)   python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_cfg2.json 2> $out/${tag}_bench_cfg2.err; tail -3 $out/${tag}_bench_cfg2.err; wc -c $out/${tag}_bench_cfg2.json
           python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_ref.json 2>/dev/null ;;
  bench2)  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $out/${tag}_bench_2gpu.json 2> $out/${tag}_bench_2gpu.err; tail -5 $out/${tag}_bench_2gpu.err; wc -c $out/${tag}_bench_2gpu.json ;;
  chunks)  python tools/chunk_sweep.py 2>&1 | tee $out/${tag}_chunk_sweep.txt ;;
  pairs)   g++ -O2 -fopenmp tools/pair_threads.cpp -Iinclude -Lcollision_detect_gjk_epa_b200 -lgjkepa_b200 -Wl,-rpath,$PWD/collision_detect_gjk_epa_b200 -o build/pair_threads
           for t in 1 4 16 32 64; do build/pair_threads $t 20000 8; done 2>&1 | tee $out/${tag}_pair_threads.txt
           build/pair_threads 32 20000 32 2>&1 | tee -a $out/${tag}_pair_threads.txt ;;
  ab)      bash tools/ab_kernels.sh $AB_LIBS 2>&1 | tee $out/${tag}_ab.txt ;;
esac; done
