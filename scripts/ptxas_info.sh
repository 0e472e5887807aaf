#!/bin/bash
# registers / spills / stack of the k_run instantiations the benchmark launches (ptxas -v); no GPU needed
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -Xptxas -v -I include \
  -o /tmp/sfl_ptxas.so network-distributed-q-learning_b200/csrc/sfl_api.cu 2>&1 | \
  awk '/Compiling entry function/ {name=$0} /Used [0-9]+ registers/ {print name; print prev; print $0} {prev=$0}' | \
  grep -A2 -E "k_runILi32ELi0ELb0ELb0ELb0ELb0|k_runILi32ELi0ELb0ELb0ELb0ELb1|k_runILi16ELi0ELb1ELb0ELb1ELb0|k_runILi16ELi0ELb0ELb0ELb1ELb1|k_runILi32ELi0ELb0ELb1ELb0ELb0|k_runILi32ELi2ELb0ELb0ELb0ELb0" | \
  sed -e 's/ptxas info    : //' -e "s/Compiling entry function '_ZN.*k_run/k_run/" -e "s/' for 'sm_100a'//"
