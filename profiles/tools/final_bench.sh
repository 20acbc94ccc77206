python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_r3r_final.json 2>/dev/null
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_r3r_reference.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r3r_final.json")); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["achieved"], d["gpu_launches"], d["cpu_baseline"]["value"], d["clocks"], d["ccc_delta"])
r=json.load(open("gpurun_out/bench_r3r_reference.json")); print(r["value"], r["cpu_baseline"])
PY
