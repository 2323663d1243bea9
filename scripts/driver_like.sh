python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
S=$(date +%s); python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/drv_ref.json 2> gpurun_out/drv_ref.err; echo "reference arm wall $(( $(date +%s) - S )) s"
S=$(date +%s); python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/drv_ours.json 2> gpurun_out/drv_ours.err; echo "our arm wall $(( $(date +%s) - S )) s"
python - <<PY
import json
r=json.load(open("gpurun_out/drv_ref.json")); d=json.load(open("gpurun_out/drv_ours.json"))
print("ref", r["value"], r["ms_per_step"], r["cpu_baseline"]["kind"], r["config"]["same_config_as_gpu_arm"])
print("ours", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["traffic"], "e2e", d["e2e"]["value"])
print(d["e2e"]["python_api_pageable_tensors"])
print("ratio", d["value"]/r["value"], "e2e ratio", d["e2e"]["value"]/r["value"])
print(d["cpu_baseline"]); print(d["executed_work"]["ncu_profile_matches_this_build"])
PY
