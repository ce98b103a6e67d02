#!/bin/bash
# usage: scripts/probe.sh <basis> <n1> <flux> [min_blocks]   -> registers, spills, SASS opcode histogram
cd "$(dirname "$0")/../dflo_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I. -DDFLO_STAGE_MIN_BLOCKS=${4:-3} -DPB=$1 -DPN=$2 -DPF=$3 \
   -Xptxas -v -c ../../scripts/probe_kernel.cu -o /tmp/probe.o 2>&1 | grep -E "error|Used|spill"
cuobjdump -sass /tmp/probe.o | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//; s/^@!?U?P[0-9T] //' | awk '{print $1}' | sed 's/;//' \
   | sort | uniq -c | sort -rn | head -${5:-25}
echo "total static instructions: $(cuobjdump -sass /tmp/probe.o | grep -cE '^\s+/\*[0-9a-f]{4}\*/')"
