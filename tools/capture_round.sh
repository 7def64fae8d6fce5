#!/bin/bash
# One GPU call's worth of evidence for profiles/: tests, the bench line, the ncu launch list of
# the bench command, and `--set full` captures of the dominant kernels (C2 refine; the HBM-regime
# score / filter launches, a light DRAM-bytes pass over the 0.6 s HBM-regime refine launch).  usage: bash tools/capture_round.sh TAG   (on the GPU box)
T=${1:-cap}
O=gpurun_out
if [ -z "$SKIP_PYTEST" ]; then python -m pytest tests -m gpu -q > $O/${T}_pytest.txt 2>&1; tail -3 $O/${T}_pytest.txt; fi
python bench.py --steps 3 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err; echo BENCH $?
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${T}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${T}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${T}_ncu_list.log 2>&1; echo LIST $?
export DP_SCENE_CACHE=/tmp/sc
python tools/profile_case.py --seeds 1048576 --full-res --reps 1 > $O/${T}_c2_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:refine_lane -c 1 -f -o $O/${T}_c2_refine \
  python tools/profile_case.py --seeds 1048576 --full-res --reps 1 > $O/${T}_c2_ncu.log 2>&1; echo C2 $?
# (gpurun brings back at most 64 MiB: export the pages that are read and drop the reports)
ncu -i $O/${T}_c2_refine.ncu-rep --page raw --csv > $O/${T}_c2_raw.csv 2>/dev/null
ncu -i $O/${T}_c2_refine.ncu-rep --page source --csv > $O/${T}_c2_src.csv 2>/dev/null
rm -f $O/${T}_c2_refine.ncu-rep
cat $O/${T}_c2_plain.log
python tools/scale_cases.py c4score --patches 1500000 --reps 1 > $O/${T}_hbm_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:score_lane -c 2 -f -o $O/${T}_hbm \
  python tools/scale_cases.py c4score --patches 1500000 --reps 1 > $O/${T}_hbm_ncu.log 2>&1 && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
  --clock-control none -k regex:refine_lane -c 1 -f -o $O/${T}_hbm_refine \
  python tools/scale_cases.py c4score --patches 1500000 --reps 1 > $O/${T}_hbm_refine_ncu.log 2>&1; echo HBM $?
ncu -i $O/${T}_hbm.ncu-rep --page raw --csv > $O/${T}_hbm_raw.csv 2>/dev/null
ncu -i $O/${T}_hbm_refine.ncu-rep --page raw --csv > $O/${T}_hbm_refine_raw.csv 2>/dev/null
rm -f $O/${T}_hbm.ncu-rep $O/${T}_hbm_refine.ncu-rep
tail -1 $O/${T}_hbm_plain.log
