# 8 x B200 box: where do the ~2.5 ms of an 8-GPU step over a 1-GPU step go?  (a) the normal run, (b) the same run with the gradient
# all-reduces skipped (TVT_DDP_DRY_RUN=1: lockstep only through the start / end barriers), (c) eight independent single-GPU runs at
# the same time (no process group at all: node power / per-GPU variation only), (d) the per-kernel overlap trace at 8 ranks.
set -x
cd $GRAFT_REPO_ROOT
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$R --master-port 29611 bench.py --gpus 8 > gpurun_out/d8_default.json 2> gpurun_out/d8_default.err
TVT_DDP_DRY_RUN=1 $R --master-port 29612 bench.py --gpus 8 --steps 15 --warmup 3 --no-extras > gpurun_out/d8_dry.json 2> gpurun_out/d8_dry.err
for i in 0 1 2 3 4 5 6 7; do
  CUDA_VISIBLE_DEVICES=$i python bench.py --steps 15 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/d8_solo_$i.json 2> gpurun_out/d8_solo_$i.err &
done
wait
$R --master-port 29613 tools/ddp_trace.py > gpurun_out/d8_trace.txt 2> gpurun_out/d8_trace.err
python - <<PY
import json
def last(f):
    try: return json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: return None
for name in ("default", "dry"):
    d = last(f"gpurun_out/d8_{name}.json")
    print(name, None if d is None else (d["n_gpus"], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["clocks"]))
for i in range(8):
    d = last(f"gpurun_out/d8_solo_{i}.json")
    print("solo", i, None if d is None else (d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"]))
PY
