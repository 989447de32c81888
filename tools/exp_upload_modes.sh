cd $GRAFT_REPO_ROOT
export ENCODER_FPS_APPS='SvtVp9EncApp_ref,SvtVp9EncApp_b200,SvtVp9EncApp_b200,SvtVp9EncApp_b200,SvtVp9EncApp_b200,SvtVp9EncApp_b200,SvtVp9EncApp_b200,SvtVp9EncApp_b200,SvtVp9EncApp_b200,SvtVp9EncApp_b200,SvtVp9EncApp_b200,SvtVp9EncApp_b200,SvtVp9EncApp_b200'
timeout 260 python tools/encoder_fps.py 3840 2160 60 8 16 > gpurun_out/r2z_enc4k_rep.jsonl 2> gpurun_out/r2z_enc4k_rep.err
python -c "
import sys, json
for l in open('gpurun_out/r2z_enc4k_rep.jsonl'):
    d = json.loads(l); g = d.get('gpu') or {}
    print(d['app'], 'fps', d['fps'], 'rc', d['rc'], d['md5'][:8] if d['md5'] else None, d['bytes'], d.get('first_difference_at'), g.get('fallbacks'), g.get('ms_upload'))
"
