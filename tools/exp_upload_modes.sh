cd $GRAFT_REPO_ROOT
export ENCODER_FPS_APPS='SvtVp9EncApp_ref,SvtVp9EncApp_b200,SvtVp9EncApp_b200,SvtVp9EncApp_b200,SvtVp9EncApp_b200,SvtVp9EncApp_ref'
timeout 200 python tools/encoder_fps.py 3840 2160 100 8 16 > gpurun_out/r2z_enc4k.jsonl 2> gpurun_out/r2z_enc4k.err
cat gpurun_out/r2z_enc4k.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); g = d.get('gpu') or {}
    print(d['app'], d['w'], 'fps', d['fps'], 'ms', d['encode_ms'], 'wall', d['wall_s'], 'rc', d['rc'], d['md5'][:8] if d['md5'] else None, d['bytes'], d.get('first_difference_at'), d.get('bytes_of_first_run'), {k: g[k] for k in g if k.startswith('ms_') or k in ('uploads', 'fallbacks')})
"
timeout 300 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 30 --log-file gpurun_out/r2z_race_me.log python -m pytest tests/test_me_gpu.py -x -q -m gpu -k "matches_oracle or m3t1 or m5t0" > gpurun_out/r2z_race_me.out 2>&1
tail -3 gpurun_out/r2z_race_me.out; tail -15 gpurun_out/r2z_race_me.log
timeout 240 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 30 --log-file gpurun_out/r2z_race_chain.log python -m pytest tests/test_blocks.py tests/test_rate.py -x -q -m gpu -k "chain or rates_of" > gpurun_out/r2z_race_chain.out 2>&1
tail -3 gpurun_out/r2z_race_chain.out; tail -15 gpurun_out/r2z_race_chain.log
