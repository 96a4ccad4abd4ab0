#!/bin/bash
# usage: scripts/dev/ptxas_k1.sh  -> one line per instantiation: CPL METRIC VIS registers spills
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xptxas -v -ccbin /usr/bin/g++ -c "$(dirname "$0")/ptxas_k1.cu" -o /dev/null 2>&1 \
 | grep -A3 "Compiling entry function" | grep -E "Compiling|spill|registers" | paste - - - \
 | sed -E 's/.*search_layer0_kernelILi([0-9]+)ELi([0-9]+)ELi([0-9]+)ELb([01]).*sm_100a.\s+(.*)ptxas info\s+: Used ([0-9]+) registers.*/CPL=\1 METRIC=\2 VIS=\3 EXCH=\4 regs=\6 \5/'
