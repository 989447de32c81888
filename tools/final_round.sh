# the end-of-round validation set on one GPU:  gpurun --timeout 1500 -- 'bash tools/final_round.sh r2z'
cd $GRAFT_REPO_ROOT
TAG=${1:-rX}
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 300 python bench.py --impl reference > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
timeout 400 python bench.py > gpurun_out/${TAG}_bench_full.json 2> gpurun_out/${TAG}_bench_full.err
timeout 600 bash tools/profile_round.sh $TAG > gpurun_out/${TAG}_profile.log 2>&1
python - $TAG <<'P'
import json, sys
t = sys.argv[1]
for f in ("bench_ref", "bench_full", "bench"):
    try:
        d = json.loads([l for l in open("gpurun_out/%s_%s.json" % (t, f)) if l.startswith("{")][-1])
        print(f, "value", d.get("value"), "e2e", (d.get("e2e") or {}).get("value"), "check", (d.get("check") or {}).get("identical"), "clocks", d.get("clocks"))
        enc = d.get("encoder") or {}
        for k, v in (enc.get("per_stream") or {}).items():
            if isinstance(v, dict) and "speedup" in v:
                print("  encoder", k, "ref", v["reference_fps_runs"], "b200", v["b200_fps_runs"], "identical", v["bitstream_identical"], "speedup", v["speedup"], (v.get("me_free_ceiling") or {}).get("fps"))
    except Exception as e:
        print(f, "failed", e)
P
