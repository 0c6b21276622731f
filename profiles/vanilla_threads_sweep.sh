#!/bin/bash
cd deeprank-gnn-2_b200/csrc
for T in 1024 768 512; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -I../../include -DDRK_VANILLA_THREADS=$T -c drk_vanilla.cu -o drk_vanilla.o || exit 1
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libdrk_b200.so drk_*.o || exit 1
  cd ../..
  echo "== threads $T"
  timeout 300 python -m pytest tests/test_gpu_vanilla_fused.py -x -q 2>&1 | tail -2
  python bench.py --config c4-vanilla --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms_per_step', d['ms_per_step'])"
  cd deeprank-gnn-2_b200/csrc
done
