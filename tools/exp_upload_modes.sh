cd $GRAFT_REPO_ROOT
export ENCODER_FPS_APPS='SvtVp9EncApp_ref,SvtVp9EncApp_b200,SvtVp9EncApp_b200,SvtVp9EncApp_ref,SvtVp9EncApp_b200'
timeout 120 python tools/encoder_fps.py 1920 1080 300 8 40 > gpurun_out/r2z_enc1080.jsonl 2> gpurun_out/r2z_enc1080.err
export ENCODER_FPS_APPS='SvtVp9EncApp_ref,SvtVp9EncApp_b200,SvtVp9EncApp_b200,SvtVp9EncApp_ref'
timeout 200 python tools/encoder_fps.py 3840 2160 100 8 16 > gpurun_out/r2z_enc4k.jsonl 2> gpurun_out/r2z_enc4k.err
cat gpurun_out/r2z_enc1080.jsonl gpurun_out/r2z_enc4k.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); g = d.get('gpu') or {}
    print(d['app'], d['w'], 'fps', d['fps'], 'ms', d['encode_ms'], 'rc', d['rc'], d['md5'][:8] if d['md5'] else None, d.get('first_difference_at'), {k: g[k] for k in g if k.startswith('ms_') or k in ('uploads', 'fallbacks')})
"
