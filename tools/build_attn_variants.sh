#!/bin/bash
# Builds libsfb200_p<mask>.so variants that differ only in ATT_POLY_MASK (share of softmax exponentials on the FMA pipe).
set -e
cd "$(dirname "$0")/.."
python -m self_forcing_b200.build > /dev/null
B=self_forcing_b200/build
for m in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --extended-lambda -Xcompiler -fPIC -Xcompiler -fvisibility=default \
       -DATT_POLY_MASK=${m}ull -c self_forcing_b200/csrc/attention_tcgen05.cu -o /tmp/att_var_$m.o
  objs=$(ls $B/*.o | grep -v attention_tcgen05.o)
  nvcc -shared -o self_forcing_b200/libsfb200_p$m.so $objs /tmp/att_var_$m.o -gencode arch=compute_100a,code=sm_100a
  echo built self_forcing_b200/libsfb200_p$m.so
done
