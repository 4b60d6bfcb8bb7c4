#!/bin/sh
# Bounds check of the kernel sources without a GPU (compute-sanitizer is closed on the GPU pool):
# builds the CPU SIMT emulator harness with AddressSanitizer and runs the emulator tests under
# it.  Shared memory is a heap vector and the global buffers are NumPy heap arrays, so an
# out-of-range index in any kernel (loads past a run, exchange-buffer overruns, row stores) aborts.
set -e
cd "$(dirname "$0")/.."
OUT=${TMPDIR:-/tmp}/libb2s_emu_asan.so
g++ -std=c++20 -O1 -g -fsanitize=address -fno-omit-frame-pointer -DB2S_EMU -shared -fPIC -pthread \
    -I tests/emu -I spectrogram_generator_b200/csrc tests/emu/emu_stft.cpp -o "$OUT"
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 B2S_EMU_LIB="$OUT" \
    python -m pytest tests/test_emulator.py -x -q -p no:cacheprovider "$@"
