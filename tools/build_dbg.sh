#!/bin/bash
# development aid: debug-clock build of the engine into tools/libnm_dbg.so (use with NM_B200_LIB=tools/libnm_dbg.so python tools/probe.py ...)
set -e
cd "$(dirname "$0")/.."
B=/tmp/nm_dbg_build; mkdir -p $B
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -O2 -w -DNM_DEBUG_CLOCKS $NM_DBG_EXTRA -I include -I neuralmelting_b200/csrc"
for t in 1025 1024 512 256 0; do nvcc $F -DNM_TU=$t -c neuralmelting_b200/csrc/nm_engine.cu -o $B/e$t.o & done
nvcc $F -c neuralmelting_b200/csrc/nm_rdf.cu -o $B/rdf.o &
nvcc $F -c neuralmelting_b200/csrc/nm_peak.cu -o $B/peak.o &
nvcc $F -c neuralmelting_b200/csrc/nm_format.cpp -o $B/fmt.o &
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o tools/libnm_dbg.so $B/*.o
