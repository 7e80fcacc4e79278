# development aid: bench.py on N GPUs under a list of environment variants, one summary line each.
# usage: N=2 [STEPS=8] [EXTRA="--scale 0.5"] tools/bench_variants.sh X=0 CGE_BANDS=2 "CGE_VIS_CULL=0 CGE_BANDS=2"
N=${N:-4}
run() { echo "== $*"; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps ${STEPS:-8} --warmup 3 --no-cpu-baseline --no-configs ${EXTRA:-} 2>>gpurun_out/n$N.err | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1])
r=d.get('roofline',{})
print('ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'match', d.get('frame_matches_1gpu'), 'stages', {k:round(v,3) for k,v in r.get('stage_ms',{}).items()}, 'prepass', r.get('light_hull_prepass'), 'frac', r.get('frac'), 'gather', d.get('nccl_gather',{}).get('ms_per_step'), 'dynamic', d.get('dynamic_tiles',{}).get('ms_per_step'))"; }
for v in "$@"; do run $v; done
