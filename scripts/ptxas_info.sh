#!/bin/bash
# Registers / spills / stack per kernel of epsm_kernels.cu (compile only, no GPU needed):  scripts/ptxas_info.sh [extra nvcc flags]
cd "$(dirname "$0")/../epsm_mitsuba3_b200/csrc"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -ccbin /usr/bin/g++ \
  -Xcompiler -fPIC,-ffp-contract=off,-O2 --expt-relaxed-constexpr -Xptxas -v "$@" -c ${SRC:-epsm_kernels.cu} -o /tmp/ptxas_info.o 2>&1 |
  awk '/Compiling entry function/ {name=$0; sub(/.*function ./,"",name); sub(/. for.*/,"",name)}
       /bytes stack frame/ {stack=$0; sub(/^ */,"",stack)}
       /Used [0-9]+ registers/ {regs=$0; sub(/.*Used /,"",regs); printf "%-72s %s | %s\n", substr(name,1,72), regs, stack}' | grep -v cub
