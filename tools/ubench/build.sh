#!/bin/sh
# builds the stand-alone tuning harnesses next to their sources (sm_100a; cross-compiles without a GPU)
set -e
cd "$(dirname "$0")"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -I ../../spectrogram_generator_b200/csrc"
nvcc $FLAGS -o duo_bench duo_bench.cu
nvcc $FLAGS -o pipes pipes.cu
nvcc $FLAGS -o mean_bench mean_bench.cu
