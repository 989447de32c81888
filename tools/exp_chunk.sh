cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_rate.py tests/test_blocks.py tests/test_bench_pipeline.py -x -q -m gpu > gpurun_out/r2c_tests.log 2>&1; tail -3 gpurun_out/r2c_tests.log
timeout 200 python bench.py --no-encoder --no-cpu --steps 20 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
python - <<'P'
import json
d = json.loads([l for l in open("gpurun_out/r2c_bench.json") if l.startswith("{")][-1])
pk = d["roofline"]["per_kernel"]
print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), {k: round(v["ms_per_picture_alone"] * 1000, 1) for k, v in pk.items()})
P
