#!/bin/bash
# per-kernel times of the parked pre-split generation (profiles/attic) at a given CTA size: bash profiles/vanilla_v2_threads_probe.sh 1024
T=${1:-1024}
cp profiles/attic/drk_vanilla_v2_presplit.cu.txt deeprank-gnn-2_b200/csrc/drk_vanilla.cu
cd deeprank-gnn-2_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -I../../include -DDRK_VANILLA_THREADS=$T -c drk_vanilla.cu -o drk_vanilla.o || exit 1
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libdrk_b200.so drk_*.o || exit 1
cd ../..
python bench.py --config c4-vanilla --no-cpu-baseline --steps 4 --warmup 3 > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_vanilla_v2_t$T.csv python bench.py --config c4-vanilla --no-cpu-baseline --steps 4 --warmup 3 > /dev/null 2>&1
python profiles/summarize_launches.py gpurun_out/launches_vanilla_v2_t$T.csv | head -5
