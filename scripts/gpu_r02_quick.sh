#!/bin/bash
# lane sweep on a short leash, simulation probe, clustering at N = 65,536 (clustered pivot pass on / off)
TAG=${1:-r02o}
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py -k "every_lane_count or cull_changes_nothing" -x -q -m gpu > gpurun_out/${TAG}_lanes.log 2>&1
echo "lane sweep exit $?"; tail -2 gpurun_out/${TAG}_lanes.log
timeout 300 python scripts/gpu_sim_probe.py seeded 2>&1 | grep -v "^{" > gpurun_out/${TAG}_probe.txt
cat gpurun_out/${TAG}_probe.txt | sed 's/"checksum.*//'
timeout 300 python scripts/gpu_cluster_big.py > gpurun_out/${TAG}_cluster_big.txt 2>&1
UPC_B200_NO_PIVOT_CLUSTER=1 timeout 300 python scripts/gpu_cluster_big.py > gpurun_out/${TAG}_cluster_big_one_cta.txt 2>&1
grep pivot gpurun_out/${TAG}_cluster_big.txt gpurun_out/${TAG}_cluster_big_one_cta.txt
timeout 300 python -m pytest tests/test_gpu_seeded.py -k "65536 or c3_subsample" -x -q -m gpu > gpurun_out/${TAG}_big_tests.log 2>&1
tail -2 gpurun_out/${TAG}_big_tests.log
