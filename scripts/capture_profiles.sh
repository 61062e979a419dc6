#!/bin/bash
# Round profile capture on the GPU box (run under gpurun): launch list of a short bench run (after the same command has
# exited 0 without ncu), one `ncu --set full` capture per front-end kernel, and the post-processing kernels.  The text
# summaries (profiles/summarize_ncu.py, by_region.py) are made on the box; only the slide kernel's .ncu-rep is kept
# (gpurun brings back at most 64 MiB).
#   bash scripts/capture_profiles.sh r02
R=${1:-r02}; O=gpurun_out
BENCH="python bench.py --clips 64 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-detect --no-stress --parity-clips 1"
$BENCH > $O/${R}_plain.log 2>&1 || { echo "plain bench failed"; tail -5 $O/${R}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_launches.csv $BENCH > $O/${R}_ncu.log 2>&1
python scripts/fe_run_once.py 64 2 > $O/${R}_once.log 2>&1 || { echo "fe_run_once failed"; exit 1; }
FR=$(grep -o "frames [0-9]*" $O/${R}_once.log | cut -d" " -f2)
for k in slide_ws_kernel anchor_tc_kernel refine_groups_kernel minmax_kernel tile_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o /tmp/prof_${R}_$k python scripts/fe_run_once.py 64 2 > $O/${R}_full_$k.log 2>&1
  python profiles/summarize_ncu.py /tmp/prof_${R}_$k.ncu-rep "ncu --set full of scripts/fe_run_once.py 64 (64 x 60 s clips, 2-7 calls/s); frames = $FR per launch" > $O/${R}_${k}_ncu.txt
done
python profiles/by_region.py /tmp/prof_${R}_slide_ws_kernel.ncu-rep > $O/${R}_slide_ws_by_phase.txt 2>&1
python profiles/by_line.py /tmp/prof_${R}_refine_groups_kernel.ncu-rep 1.5 > $O/${R}_refine_groups_by_line.txt 2>&1
cp /tmp/prof_${R}_slide_ws_kernel.ncu-rep $O/
NBM_STRESS=1 python scripts/fe_run_once.py 32 2 > $O/${R}_once_stress.log 2>&1 && \
ncu --set full --clock-control none -k regex:slide_ws_kernel -s 1 -c 1 -o /tmp/prof_${R}_stress_slide env NBM_STRESS=1 python scripts/fe_run_once.py 32 2 > $O/${R}_full_stress.log 2>&1
FS=$(grep -o "frames [0-9]*" $O/${R}_once_stress.log | cut -d" " -f2)
python profiles/summarize_ncu.py /tmp/prof_${R}_stress_slide.ncu-rep "ncu --set full, n_fft 4410 / hop 44 (NBM_STRESS=1 scripts/fe_run_once.py 32); frames = $FS per launch" > $O/${R}_stress_slide_ws_kernel_ncu.txt
python scripts/bench_postproc.py > $O/${R}_postproc.json 2>&1
ncu --set full --clock-control none -k regex:"nms_|final_detections|rpn_decode|proposal_|roi_pool|merge_" -c 40 -o /tmp/prof_${R}_postproc python scripts/bench_postproc.py > $O/${R}_full_postproc.log 2>&1
python profiles/summarize_ncu.py /tmp/prof_${R}_postproc.ncu-rep "ncu --set full of scripts/bench_postproc.py, first 40 launches of the post-processing kernels" > $O/${R}_postproc_kernels_ncu.txt
cat $O/${R}_once.log $O/${R}_once_stress.log; du -sh $O
