#!/bin/bash
# torchrun --no-python profiles/ncu_rank0.sh <ncu-output-base> <script> [args...]: rank 0 under ncu, the others plain.
# K4 reads CUDA-IPC peer memory: multi-pass replay (memory save / restore) fails on it ("UnknownError" with
# --set full), so the capture asks for counters that fit single passes: NCU_METRICS overrides the list.
out=$1; shift
METRICS=${NCU_METRICS:-gpu__time_duration.sum,nvlrx__bytes.sum,nvltx__bytes.sum,nvlrx__bytes_data_user.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_ltcfabric.sum}
if [ "${LOCAL_RANK:-0}" = "0" ]; then
  exec ncu --metrics "$METRICS" --clock-control none --cache-control none --replay-mode kernel -k regex:"k_window_reduce" -s 2 -c 2 --csv --log-file "$out.csv" python "$@"
else
  exec python "$@"
fi
