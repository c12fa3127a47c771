#!/bin/bash
# (r03z: the `full` step ran 13 of the visit's 18 GPU-minutes because profile_scene.py did all of its passes under ncu; it now does one warm-up
# step and one captured step when SGB_PROFILE_FAST is set.)
# End-of-round GPU visit: parity tests, the bench line, the reference arm, per-kernel rooflines, the ncu launch list of the bench
# command and one `ncu --set full` capture of the 8-scene training step.  usage (under gpurun): bash tools/gpu_final.sh <tag> [steps...]
TAG=${1:-r03z}; shift
STEPS=${@:-"tests bench benchref kernels launches full"}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
for s in $STEPS; do
case $s in
tests)    timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.txt 2>&1; tail -3 $O/${TAG}_pytest_gpu.txt ;;
bench)    timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -c 300 $O/${TAG}_bench.json ;;
benchref) timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $O/${TAG}_bench_ref.json 2>&1 ;;
kernels)  timeout 900 python tools/bench_kernels.py --out $O/${TAG}_kernels.json > $O/${TAG}_kernels.log 2>&1; tail -2 $O/${TAG}_kernels.log | cut -c1-200 ;;
launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/${TAG}_launches.csv \
            python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra > $O/${TAG}_launches.log 2>&1
          python tools/summarize_launches.py $O/${TAG}_launches.csv > $O/${TAG}_launches_bench.md 2>&1; head -12 $O/${TAG}_launches_bench.md ;;
full)     SGB_PROFILE_FAST=1 timeout 600 ncu --set full --clock-control none \
            -k regex:'ec2_tc1_kernel|ec2_bwd_tc_kernel|segment_pool_staged_kernel|knn_sweep_kernel|forward_max_kernel|gram1_pt_kernel|centralize_kernel|export_labels_kernel|gcn_agg|segment_pool_bwd|group_nearby_kernel|gemm_tf32x3|cls_head|bwd_sparse' \
            --launch-skip 110 -c 110 -f -o /tmp/${TAG}_full_train python tools/profile_scene.py 150000 train 4.0 8 > $O/${TAG}_full_train.log 2>&1
          python tools/ncu_summary.py /tmp/${TAG}_full_train.ncu-rep --json $O/${TAG}_train_8x150k_ncu.json --points 150000 > $O/${TAG}_ncu_full_train_8x150k.txt 2>&1
          tail -3 $O/${TAG}_full_train.log ;;
esac
done
ls -la $O | tail -12
