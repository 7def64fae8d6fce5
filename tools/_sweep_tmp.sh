for cfg in "2368 592 64 128" "2368 592 96 128" "1184 592 64 128" "4736 592 64 128" "2368 296 64 64" "2368 1184 128 128"; do
  set -- $cfg
  DP_SLICE_T4=$1 DP_SLICE_T8=$2 DP_SLICE_B1=$3 DP_SLICE_B4=$4 python tools/shard_balance.py > gpurun_out/r3b_bal_$1_$2_$3_$4.txt 2>&1
  echo "T4=$1 T8=$2 B1=$3 B4=$4: $(tail -1 gpurun_out/r3b_bal_$1_$2_$3_$4.txt)"
done
