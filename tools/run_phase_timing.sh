#!/bin/bash
# Builds a copy of the library whose fast forward interpreter records clock64 stamps (DFOL_PROG_TIMING) and prints the
# per-phase cycle counts of one question block (tools/time_program_phases.py).  Usage: tools/run_phase_timing.sh [batch]
set -e
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -Iinclude -Idfol_vqa_b200/csrc \
     -DDFOL_PROG_TIMING -c dfol_vqa_b200/csrc/program_fwd_fast.cu -o /tmp/pf_timing.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o /tmp/libdfol_timing.so \
     $(ls dfol_vqa_b200/build/*.o | grep -v program_fwd_fast.o) /tmp/pf_timing.o -lcudart
python tools/time_program_phases.py /tmp/libdfol_timing.so "$@"
