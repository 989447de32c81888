cd $GRAFT_REPO_ROOT
run() { # name, env...
  name=$1; shift
  env "$@" timeout 200 python bench.py --no-encoder --no-cpu --verify --steps 20 --warmup 3 > gpurun_out/r2e_$name.json 2> gpurun_out/r2e_$name.err
  python - "$name" <<'P'
import json, sys
n = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/r2e_%s.json" % n) if l.startswith("{")][-1])
    print(n, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "verify", d.get("verify"))
except Exception as e:
    print(n, "failed", e, open("gpurun_out/r2e_%s.err" % n).read()[-600:])
P
}
run old SVTB200_BENCH_SLOTS=3 SVTB200_E2E_STREAM_LAG=1
run s4l2 SVTB200_BENCH_SLOTS=4 SVTB200_E2E_STREAM_LAG=2
run s5l3 A=1
run s6l4 SVTB200_BENCH_SLOTS=6 SVTB200_E2E_STREAM_LAG=4
run s5l3_mapped SVTB200_E2E_PACK=mapped
