python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2w_pytest.log; cat gpurun_out/r2w_pytest.log
for i in 1 2; do
for v in 1 0; do
JMT_PDL_ALL=$v python bench.py --steps 40 --warmup 5 > gpurun_out/bench_r2w_pdl$v.json 2> gpurun_out/bench_r2w.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r2w_pdl$v.json"))
print("pdl_all=$v", d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"])
PY
done
done
