#!/bin/bash
# phase clocks of k_vanilla_fwd: rebuilds drk_vanilla.o with the probes on the GPU box, runs the probe, restores nothing (scratch copy)
cd deeprank-gnn-2_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -I../../include -DDRK_VANILLA_PROBE -c drk_vanilla.cu -o drk_vanilla.o || exit 1
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libdrk_b200.so drk_*.o || exit 1
cd ../..
python profiles/vanilla_phase_probe.py
