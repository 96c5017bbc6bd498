#!/bin/bash
# One gpurun call: GPU parity tests, variant benches, launch list and full-set ncu captures.  Everything lands in gpurun_out/.
# usage: gpurun --timeout 1500 -- 'bash tools_gpu_session.sh [tag]'
TAG=${1:-s}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log
tail -n 3 $O/${TAG}_pytest_gpu.log
B="--spp 64 --steps 2 --warmup 1 --no-cpu --e2e-steps 1"
for S in 2048 4096 7085; do
  RSB_SLOTS_PER_SM=$S timeout 300 python bench.py $B > $O/${TAG}_b_recip_$S.json 2> $O/${TAG}_b_recip_$S.err
done
RSB_LIBRARY=$PWD/build/lib_truediv.so timeout 300 python bench.py $B > $O/${TAG}_b_div_2048.json 2> $O/${TAG}_b_div_2048.err
timeout 600 python bench.py > $O/${TAG}_bench_full.json 2> $O/${TAG}_bench_full.err
for f in $O/${TAG}_b_*.json $O/${TAG}_bench_full.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d["roofline"]
    print(sys.argv[1].split("/")[-1], "value %.1f e2e %.1f ms/step %.1f trace_ms %.4f share %.3f frac %.3f waves %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r["kernel_ms"], r["kernel_share_of_step"], r["frac"], d["waves_per_step"]))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
export RSB_NO_GRAPH=1
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -s 800 -c 200 --csv \
  --log-file $O/${TAG}_launches.csv python bench.py --pixels 1024 --spp 16 --steps 1 --warmup 1 --no-cpu --e2e-steps 1 > $O/${TAG}_ncu_list.log 2>&1
python tools_kernel_summary.py $O/${TAG}_launches.csv > $O/${TAG}_launch_summary.txt 2>&1; cat $O/${TAG}_launch_summary.txt
for K in k_wf_trace k_wf_shade k_wf_finalize; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$K -s 300 -c 1 -f -o $O/${TAG}_full_$K \
    python bench.py --pixels 1024 --spp 16 --steps 1 --warmup 1 --no-cpu --e2e-steps 1 > $O/${TAG}_ncu_$K.log 2>&1
done
ls -la $O | tail -n 30
