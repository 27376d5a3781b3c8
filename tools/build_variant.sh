#!/bin/bash
# usage: tools/build_variant.sh <name> "<nvcc -D flags>"  -- libmvtools_b200_<name>.so = the default objects with search_tma.cu rebuilt with the flags
set -e
cd "$(dirname "$0")/../vapoursynth-mvtools_b200"
mkdir -p obj_$1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-fvisibility=hidden,-Wall -cudart static $2 -Xptxas -v -c csrc/search_tma.cu -o obj_$1/search_tma.o 2> obj_$1/search_tma.ptxas.log
grep -E "spill|registers" obj_$1/search_tma.ptxas.log | paste - - | awk '{print $1,$2,$3,$4,$5,$6,$7,$8,$9,$10,$11,$12,$13,$14,$15,$16,$17,$18}' | tail -6
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o libmvtools_b200_$1.so obj/engine.o obj/super.o obj/search.o obj/search_fast.o obj_$1/search_tma.o obj/degrain.o obj/compensate.o obj/flow.o obj/host_logic.o
