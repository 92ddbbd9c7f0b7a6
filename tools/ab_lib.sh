#!/bin/bash
# Dev tool: run a command once per library variant (libsa_b200_<tag>.so swapped in as libsa_b200.so)
set -e
cd "$(dirname "$0")/.."
P=sequence-alignment-gpu_b200
cp $P/libsa_b200.so /tmp/libsa_b200_orig.so
for tag in "$@"; do
  if [ "$tag" == "--" ]; then break; fi
done
