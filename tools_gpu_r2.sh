#!/bin/bash
# round-2 GPU session helper: bash tools_gpu_r2.sh TAG step...
TAG=${1:-s}; shift
O=gpurun_out; mkdir -p $O
for S in "$@"; do
case $S in
test)
  timeout 1200 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log
  tail -n 15 $O/${TAG}_pytest_gpu.log ;;
testk:*)
  K=${S#testk:}; K="${K//_/ }"
  timeout 900 python -m pytest tests -m gpu -x -q -k "$K" > $O/${TAG}_pytest_k.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_k.log
  tail -n 25 $O/${TAG}_pytest_k.log ;;
sanitize)
  timeout 900 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -x -q -k "zoo or mesh_hits or edge or rgb_pipeline_on_device" > $O/${TAG}_sanitize.log 2>&1
  tail -n 12 $O/${TAG}_sanitize.log ;;
sweep|sweep:*)
  L=${S#sweep:}; [ "$L" = "sweep" ] && L=""
  if [ -n "$L" ]; then export RSB_LIBRARY=$PWD/build/$L.so; fi
  timeout 900 python tools_sweep.py --n ${SWEEP_N:-1e7} --mesh-subdiv ${SWEEP_SUBDIV:-8} > $O/${TAG}_sweep_$L.jsonl 2> $O/${TAG}_sweep_$L.err
  echo "== sweep $L"; python - $O/${TAG}_sweep_$L.jsonl <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l)
    if "rays" in d: print(d["scene"], d.get("order"), "rays %g"%d["rays"], "Mrays/s %.1f"%d["Mrays_per_s"], "frac %.3f"%d["roofline_frac"], "KB/ray %.2f"%(d["algorithmic_bytes_per_ray"]/1e3), "hit %.3f"%d["hit_fraction"])
PY
  tail -n 3 $O/${TAG}_sweep_$L.err
  unset RSB_LIBRARY ;;
ncumesh)
  timeout 900 ncu --set full --clock-control none -k regex:k_rq_mesh -s 1 -c 1 -f -o $O/${TAG}_full_rq_mesh \
    python tools_sweep.py --n 4e6 --mesh-subdiv 8 --no-spheres --order ${NCU_ORDER:-random} > $O/${TAG}_ncumesh.log 2>&1
  python tools_ncu_summary.py $O/${TAG}_full_rq_mesh.ncu-rep k_rq_mesh > $O/${TAG}_ncu_full_k_rq_mesh.txt 2>&1; rm -f $O/${TAG}_full_rq_mesh.ncu-rep
  cat $O/${TAG}_ncu_full_k_rq_mesh.txt ;;
ncuworld)
  timeout 900 ncu --set full --clock-control none -k regex:k_rq_world -s 1 -c 1 -f -o $O/${TAG}_full_rq_world \
    python tools_sweep.py --n 4e6 --order ${NCU_ORDER:-random} > $O/${TAG}_ncuworld.log 2>&1
  python tools_ncu_summary.py $O/${TAG}_full_rq_world.ncu-rep k_rq_world > $O/${TAG}_ncu_full_k_rq_world.txt 2>&1; rm -f $O/${TAG}_full_rq_world.ncu-rep
  cat $O/${TAG}_ncu_full_k_rq_world.txt ;;
list)
  timeout 900 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -c 200 --csv \
    --log-file $O/${TAG}_sweep_launches.csv python tools_sweep.py --n ${LIST_N:-4e6} --mesh-subdiv 8 --order ${NCU_ORDER:-random} > $O/${TAG}_list.log 2>&1; tail -n 2 $O/${TAG}_list.log
  python tools_kernel_summary.py $O/${TAG}_sweep_launches.csv > $O/${TAG}_sweep_launch_summary.txt 2>&1; cat $O/${TAG}_sweep_launch_summary.txt ;;
bench|bench:*)
  A=${S#bench:}; [ "$A" = "bench" ] && A="--steps 2 --warmup 1 --e2e-steps 1"
  A="${A//_/ }"
  timeout 1500 python bench.py $A > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; python tools_bench_show.py $O/${TAG}_bench.json; tail -n 5 $O/${TAG}_bench.err ;;
listmesh)
  RSB_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -s 600 -c 400 --csv \
    --log-file $O/${TAG}_rendermesh_launches.csv python tools_render_mesh.py --pixels 1024 --spp 4 > $O/${TAG}_listmesh.log 2>&1; tail -n 2 $O/${TAG}_listmesh.log
  python tools_kernel_summary.py $O/${TAG}_rendermesh_launches.csv > $O/${TAG}_rendermesh_launch_summary.txt 2>&1; cat $O/${TAG}_rendermesh_launch_summary.txt ;;
rendermesh|rendermesh:*)
  A=${S#rendermesh:}; [ "$A" = "rendermesh" ] && A="--pixels 1024 --spp 16"
  A="${A//_/ }"
  timeout 900 python tools_render_mesh.py $A > $O/${TAG}_rendermesh.json 2> $O/${TAG}_rendermesh.err; echo "rendermesh $A"; cut -c1-500 $O/${TAG}_rendermesh.json; tail -n 3 $O/${TAG}_rendermesh.err ;;
listprism)
  RSB_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -s 300 -c 600 --csv \
    --log-file $O/${TAG}_prism_launches.csv python tools_render_prism.py --pixels 256 --rays 128 --bins 128 > $O/${TAG}_listprism.log 2>&1; tail -n 2 $O/${TAG}_listprism.log
  python tools_kernel_summary.py $O/${TAG}_prism_launches.csv > $O/${TAG}_prism_launch_summary.txt 2>&1; cat $O/${TAG}_prism_launch_summary.txt ;;
prism|prism:*)
  A=${S#prism:}; [ "$A" = "prism" ] && A=""
  A="${A//_/ }"
  for CH in ${PRISM_CHUNKS:-8388608}; do
    RSB_CHUNK_ITEMS=$CH timeout 900 python tools_render_prism.py $A > $O/${TAG}_prism_$CH.json 2> $O/${TAG}_prism_$CH.err; echo "prism chunk $CH $A"; cut -c1-400 $O/${TAG}_prism_$CH.json; tail -n 2 $O/${TAG}_prism_$CH.err
  done ;;
full)
  timeout 1700 python bench.py > $O/${TAG}_bench_full.json 2> $O/${TAG}_bench_full.err; python tools_bench_show.py $O/${TAG}_bench_full.json; tail -n 3 $O/${TAG}_bench_full.err ;;
listbench)
  RSB_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -s 2000 -c 800 --csv \
    --log-file $O/${TAG}_bench_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-plugin --configs none --e2e-steps 1 > $O/${TAG}_listbench.log 2>&1; tail -n 2 $O/${TAG}_listbench.log
  python tools_kernel_summary.py $O/${TAG}_bench_launches.csv > $O/${TAG}_bench_launch_summary.txt 2>&1; cat $O/${TAG}_bench_launch_summary.txt ;;
ncuwf:*)
  K=${S#ncuwf:}
  RSB_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none -k regex:$K -s 300 -c 1 -f -o $O/${TAG}_full_$K \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-plugin --configs none --e2e-steps 1 > $O/${TAG}_ncu_$K.log 2>&1
  python tools_ncu_summary.py $O/${TAG}_full_$K.ncu-rep $K > $O/${TAG}_ncu_full_$K.txt 2>&1; rm -f $O/${TAG}_full_$K.ncu-rep
  head -n 32 $O/${TAG}_ncu_full_$K.txt ;;
*) echo "unknown step $S" ;;
esac
done
