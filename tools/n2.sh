run() { echo "== $*"; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline --no-configs 2>gpurun_out/n2.err | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1])
print('ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'match', d.get('frame_matches_1gpu'), 'stages', {k:round(v,3) for k,v in d['roofline']['stage_ms'].items()}, 'prepass', d['roofline'].get('light_hull_prepass'))"; }
run X=0
run CGE_BANDS=2
run CGE_BANDS=1
run CGE_VIS_CULL=1
run CGE_VIS_CULL=0 CGE_BANDS=2
