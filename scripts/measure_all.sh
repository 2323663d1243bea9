#!/bin/bash
# Full measurement pass on the GPU box (run from the repo root): writes everything profiles/ cites into gpurun_out/.
#   gpurun --timeout 1500 -- 'bash scripts/measure_all.sh'
# then, here:  python scripts/profile_to_json.py ncu gpurun_out   (-> profiles/kernel_profile.json, read by bench.py)
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem,pcie.link.gen.current,pcie.link.width.current --format=csv > $O/gpu.txt
python -m pytest tests -m gpu -q 2>&1 | tail -2 > $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
python scripts/parity_report.py --full > $O/parity.json 2> $O/parity.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
python bench.py --steps 200 --warmup 5 > $O/bench_c2.json 2> $O/bench_c2.err
for w in c1 c3 c4 c5; do python bench.py --workload $w --steps 100 --warmup 5 --no-cpu-baseline --no-c5 > $O/bench_$w.json 2> $O/bench_$w.err; done
for w in c2 c3; do python scripts/kernel_times.py $w > $O/kernel_times_$w.jsonl 2>&1; done
python scripts/api_overhead.py > $O/api_overhead.txt 2>&1
QUICK="--steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-c5"
# launch list of the bench command (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-c5 > $O/ncu_bench.log 2>&1
# DRAM traffic of the loss kernel per launch, every workload (-> profiles/kernel_profile.json)
for w in c1 c2 c3 c4 c5; do
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:loss_kernel -s 3 -c 1 \
      --csv --log-file $O/traffic_$w.csv python bench.py --workload $w $QUICK > $O/ncu_traffic_$w.log 2>&1
done
# one full capture of the dominant kernel (+ its source page), of the render kernels
ncu --set full --clock-control none --import-source on -k regex:loss_kernel -s 3 -c 1 -f -o $O/prof_loss \
    python bench.py $QUICK > $O/ncu_full.log 2>&1
ncu -i $O/prof_loss.ncu-rep --page raw --csv > $O/prof_loss_raw.csv 2>/dev/null
ncu -i $O/prof_loss.ncu-rep --page source --csv > $O/prof_loss_source.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:render_bwd_kernel -s 2 -c 1 -f -o $O/prof_render_bwd \
    python scripts/kernel_times.py c2 > $O/ncu_full_rbwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:render_fwd_kernel -s 2 -c 1 -f -o $O/prof_render_fwd \
    python scripts/kernel_times.py c2 > $O/ncu_full_rfwd.log 2>&1
python - <<'PY'
import json
for w in ("c1", "c2", "c3", "c4", "c5"):
    try:
        d = json.load(open("gpurun_out/bench_%s.json" % w))
        print(w, round(d["value"], 1), round(d["ms_per_step"], 4), round(d["roofline"]["frac"], 3), (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(w, "failed", e)
PY
cat $O/pytest_gpu.log $O/smoke.log
