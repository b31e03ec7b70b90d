#!/bin/bash
# usage: tools/sasshist.sh <file.cu> <kernel-name-substring> [extra nvcc flags]
# Compiles for sm_100a and prints the static SASS opcode histogram and register use of one kernel.
set -e
src=$1; pat=$2; shift 2
out=/tmp/sasshist_$$.cubin
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xptxas -v -cubin -o $out "$@" $src 2> /tmp/sasshist_$$.log || { cat /tmp/sasshist_$$.log; exit 1; }
grep -A3 "Compiling entry function.*$pat" /tmp/sasshist_$$.log | grep -E "registers|spill" | head -4
fn=$(cuobjdump -sass $out | grep "Function :" | grep "$pat" | head -1 | awk '{print $3}')
echo "kernel: $fn"
# main path only: stop at the first unpredicated EXIT (the BRA.DIV slow-path duplicates follow it)
cuobjdump -sass -fun "$fn" $out | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | sed -E 's/^\s+\/\*[0-9a-f]{4,5}\*\/\s+//' | awk '{print} /^EXIT/ {exit}' | sed -E 's/^@!?U?P[0-9T] +//' | awk '{print $1}' | sed 's/\..*//; s/;//' | sort | uniq -c | sort -rn | awk '{t+=$1; print} END {print t, "TOTAL"}' | head -${TOP:-28}
rm -f $out /tmp/sasshist_$$.log
