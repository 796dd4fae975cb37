#!/bin/bash
# torchrun --no-python profiles/ncu_rank0.sh <ncu-output-base> <script> [args...]: rank 0 under ncu, the others plain
out=$1; shift
if [ "${LOCAL_RANK:-0}" = "0" ]; then
  exec ncu --set full --import-source on --clock-control none ${NCU_EXTRA} -k regex:"k_window_reduce" -s 2 -c 1 -f -o "$out" python "$@"
else
  exec python "$@"
fi
