#!/bin/bash
# One GPU-box visit: parity tests, the bench line, per-kernel rooflines, host profile, ncu captures.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [steps...]   steps default: all
TAG=${1:-rXX}; shift
STEPS=${@:-"tests bench kernels host full launches"}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
for s in $STEPS; do
case $s in
tests)    timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.txt 2>&1; tail -3 $O/${TAG}_pytest_gpu.txt ;;
bench)    timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -c 600 $O/${TAG}_bench.json ;;
benchref) timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $O/${TAG}_bench_ref.json 2>&1 ;;
kernels)  timeout 900 python tools/bench_kernels.py --out $O/${TAG}_kernels.json > $O/${TAG}_kernels.log 2>&1; tail -2 $O/${TAG}_kernels.log ;;
host)     timeout 600 python tools/host_profile.py > $O/${TAG}_host_profile.txt 2>&1 ;;
scene)    timeout 600 python tools/profile_scene.py 150000 train > $O/${TAG}_entrypoints_train_150k.txt 2>&1
          timeout 600 python tools/profile_scene.py 150000 ins_infer > $O/${TAG}_entrypoints_ins_infer_150k.txt 2>&1 ;;
full)     # ncu --set full captures; only the text summaries travel back (gpurun_out/ is capped at 64 MiB)
          timeout 900 ncu --set full --clock-control none \
            -k regex:'segment_pool_staged_kernel|ec2_tc_kernel|ec2_bwd_tc_kernel|kpconv_tc_fwd_kernel|knn_sweep_kernel|forward_max_kernel|gram1_pt_kernel|centralize_kernel|ind_max_pool_fwd' \
            -c 16 -f -o /tmp/${TAG}_full_kernels python tools/bench_kernels.py --points 500000 --reps 1 --warm 0 --only pool,centralize,knn,edgeconv,kpconv,kppool > $O/${TAG}_full_kernels.log 2>&1
          python tools/ncu_summary.py /tmp/${TAG}_full_kernels.ncu-rep > $O/${TAG}_ncu_full_kernels_500k.txt 2>&1
          timeout 900 ncu --set full --clock-control none \
            -k regex:'ec2_tc_kernel|ec2_bwd_tc_kernel|segment_pool_staged_kernel|group_nearby_kernel|unlabeled_union_kernel|bwd_sparse_kernel|gram1_pt_kernel|knn_sweep_kernel' \
            -c 18 -f -o /tmp/${TAG}_full_train python tools/profile_scene.py 150000 train > $O/${TAG}_full_train.log 2>&1
          python tools/ncu_summary.py /tmp/${TAG}_full_train.ncu-rep > $O/${TAG}_ncu_full_train_150k.txt 2>&1
          for f in /tmp/${TAG}_full_train.ncu-rep; do [ $(stat -c %s $f) -lt 40000000 ] && cp $f $O/; done ;;
gridrn)   timeout 600 python tools/bench_kernels.py --only pool,centralize,grid,neighbors,kppool --out $O/${TAG}_kernels_gather.json > $O/${TAG}_kernels_gather.log 2>&1; tail -3 $O/${TAG}_kernels_gather.log
          timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_gridrn.csv \
            python tools/bench_kernels.py --points 150000 --reps 1 --warm 1 --only grid,neighbors > $O/${TAG}_launches_gridrn.log 2>&1
          timeout 900 ncu --set full --clock-control none --import-source on \
            -k regex:'segment_pool_fwd_kernel|rn_search|rn_fill|centralize_kernel|ind_max_pool_fwd|gs_insert|gs_reduce_points|gs_sort_small' \
            -c 12 -f -o $O/${TAG}_full_gather python tools/bench_kernels.py --points 500000 --reps 1 --warm 0 --only pool,centralize,grid,neighbors,kppool > $O/${TAG}_full_gather.log 2>&1 ;;
launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $O/${TAG}_launches.csv \
            python bench.py --steps 1 --warmup 1 --scenes 1 --points 150000 --no-cpu-baseline --streams 1 > $O/${TAG}_launches.log 2>&1
          python tools/summarize_launches.py $O/${TAG}_launches.csv > $O/${TAG}_launches_train_150k.md 2>&1 ;;
esac
done
ls -la $O | tail -20
