#!/bin/bash
# Dev tool: build libsa_b200_<tag>.so with extra nvcc flags for sa_tile.cu only (A/B of the tile kernel's compile-time knobs);
# run with SA_B200_LIB=sequence-alignment-gpu_b200/libsa_b200_<tag>.so.   tools/build_variant.sh tg4 -DSA_TILE_TG=4 -DSA_TILE_TG_REQ=1
set -e
cd "$(dirname "$0")/../sequence-alignment-gpu_b200/csrc"
tag=$1; shift
nvcc "$@" -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-O3,-Wall --cudart static -DSA_TILE_PART=1 -c -o /tmp/sa_tile_$tag.o sa_tile.cu
nvcc "$@" -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-O3,-Wall --cudart static -DSA_TILE_PART=2 -c -o /tmp/sa_tile2_$tag.o sa_tile.cu
nvcc -gencode arch=compute_100a,code=sm_100a -shared --cudart static -o ../libsa_b200_$tag.so sa_api.o /tmp/sa_tile_$tag.o /tmp/sa_tile2_$tag.o sa_shim.o sa_frontend.o sa_utilities.o
echo built ../libsa_b200_$tag.so
