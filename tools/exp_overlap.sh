cd $GRAFT_REPO_ROOT
run() { # name, env...
  name=$1; shift
  env "$@" timeout 200 python bench.py --no-encoder --no-cpu --steps 20 --warmup 3 > gpurun_out/r2o_$name.json 2> gpurun_out/r2o_$name.err
  python - "$name" <<'P'
import json, sys
n = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/r2o_%s.json" % n) if l.startswith("{")][-1])
    pk = d["roofline"]["per_kernel"]["me_sb_kernel"]
    print(n, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "me_in_region_ms", round(pk.get("ms_per_launch_in_timed_region", 0), 4), "me_alone", round(pk.get("ms_per_picture_alone", 0), 4))
except Exception as e:
    print(n, "failed", e)
P
}
run base A=1
run me3 SVTB200_ME_CTAS_PER_SM=3
run me2 SVTB200_ME_CTAS_PER_SM=2
run me3p SVTB200_ME_CTAS_PER_SM=3 SVTB200_COMPUTE2_PRIORITY=1
run nopersist_p SVTB200_ME_NO_PERSISTENT=1 SVTB200_COMPUTE2_PRIORITY=1
run base2 A=1
