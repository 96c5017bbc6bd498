#!/bin/bash
# One gpurun call: GPU parity tests, variant benches, launch list and full-set ncu captures.  Everything lands in gpurun_out/.
# usage: gpurun --timeout 1500 -- 'bash tools_gpu_session.sh TAG [steps...]'   steps: test bench variants full list ncu:<kernel>
TAG=${1:-s}; shift
STEPS="${@:-test variants list}"
O=gpurun_out
mkdir -p $O
B="--spp 64 --steps 2 --warmup 1 --no-cpu --e2e-steps 1"
show() { python - "$@" <<'PY'
import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d["roofline"]
        print(f.split("/")[-1], "value %.1f e2e %.1f ms/step %.1f trace_ms %.4f share %.3f frac %.3f waves %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r["kernel_ms"], r["kernel_share_of_step"], r["frac"], d["waves_per_step"]))
    except Exception as e:
        print(f, "ERR", e)
PY
}
for S in $STEPS; do
case $S in
test)
  timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log
  tail -n 4 $O/${TAG}_pytest_gpu.log ;;
variants)
  for V in ${RSB_VARIANT_SLOTS:-2048 7085}; do
    RSB_SLOTS_PER_SM=$V timeout 300 python bench.py $B > $O/${TAG}_b_$V.json 2> $O/${TAG}_b_$V.err; show $O/${TAG}_b_$V.json
  done ;;
lib:*)
  L=${S#lib:}
  RSB_LIBRARY=$PWD/build/$L.so timeout 300 python bench.py $B > $O/${TAG}_b_$L.json 2> $O/${TAG}_b_$L.err; show $O/${TAG}_b_$L.json ;;
passes)
  for V in ${RSB_VARIANT_PASSES:-1 2 4 8}; do
    timeout 400 python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 --passes $V > $O/${TAG}_b_passes$V.json 2> $O/${TAG}_b_passes$V.err; show $O/${TAG}_b_passes$V.json
  done ;;
fulllib:*)
  L=${S#fulllib:}
  if [ "$L" != "main" ]; then export RSB_LIBRARY=$PWD/build/$L.so; fi
  timeout 400 python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 > $O/${TAG}_bfull_$L.json 2> $O/${TAG}_bfull_$L.err; show $O/${TAG}_bfull_$L.json
  unset RSB_LIBRARY ;;
philox)
  timeout 300 python bench.py $B --rng philox > $O/${TAG}_b_philox.json 2> $O/${TAG}_b_philox.err; show $O/${TAG}_b_philox.json ;;
full)
  timeout 900 python bench.py > $O/${TAG}_bench_full.json 2> $O/${TAG}_bench_full.err; show $O/${TAG}_bench_full.json ;;
list|list:*)
  V=${S#list:}; [ "$V" = "list" ] && V=8192
  RSB_SLOTS_PER_SM=$V RSB_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none -s 800 -c 120 --csv \
    --log-file $O/${TAG}_launches_$V.csv python bench.py --pixels 1024 --spp 64 --steps 1 --warmup 1 --no-cpu --e2e-steps 1 > $O/${TAG}_ncu_list_$V.log 2>&1
  python tools_kernel_summary.py $O/${TAG}_launches_$V.csv > $O/${TAG}_launch_summary_$V.txt 2>&1; echo "slots/SM $V"; cat $O/${TAG}_launch_summary_$V.txt ;;
ncu:*)
  K=${S#ncu:}
  RSB_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none -k regex:$K -s 200 -c 1 -f -o $O/${TAG}_full_$K \
    python bench.py --pixels 1024 --spp 64 --steps 1 --warmup 1 --no-cpu --e2e-steps 1 > $O/${TAG}_ncu_$K.log 2>&1 ;;
rendermesh|rendermesh:*)
  L=${S#rendermesh:}; [ "$L" = "rendermesh" ] && L=""
  if [ -n "$L" ]; then export RSB_LIBRARY=$PWD/build/$L.so; fi
  timeout 900 python tools_render_mesh.py --pixels 1024 --spp 16 > $O/${TAG}_rendermesh_$L.json 2> $O/${TAG}_rendermesh_$L.err; echo "rendermesh $L"; cut -c1-400 $O/${TAG}_rendermesh_$L.json; tail -n 2 $O/${TAG}_rendermesh_$L.err
  unset RSB_LIBRARY ;;
sweeplib:*)
  L=${S#sweeplib:}
  RSB_LIBRARY=$PWD/build/$L.so timeout 900 python tools_sweep.py --n 1e7 --mesh-subdiv 8 > $O/${TAG}_sweep_$L.jsonl 2> $O/${TAG}_sweep_$L.err
  echo "== $L"; python - $O/${TAG}_sweep_$L.jsonl <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l)
    if "rays" in d: print(d["scene"], "rays %g"%d["rays"], "Mrays/s %.1f"%d["Mrays_per_s"], "frac %.3f"%d["roofline_frac"], "KB/ray %.2f"%(d["algorithmic_bytes_per_ray"]/1e3))
PY
  ;;
ncusweep)
  timeout 900 ncu --set full --clock-control none -k regex:k_hit_sweep -s 2 -c 1 -f -o $O/${TAG}_full_sweep_mesh \
    python tools_sweep.py --n 4e6 --mesh-subdiv 8 --no-spheres > $O/${TAG}_ncusweep.log 2>&1; tail -n 3 $O/${TAG}_ncusweep.log ;;
sweep)
  timeout 900 python tools_sweep.py --n 1e7 --mesh-subdiv 8 > $O/${TAG}_sweep.jsonl 2> $O/${TAG}_sweep.err; cut -c1-400 $O/${TAG}_sweep.jsonl ;;
*) echo "unknown step $S" ;;
esac
done
