set -x
cd $GRAFT_REPO_ROOT
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
python bench.py --steps 10 --no-extras --no-cpu-baseline > gpurun_out/s8_n1.json 2> gpurun_out/s8_n1.err
$R --master-port 29601 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/s8_default.json 2> gpurun_out/s8_default.err
$R --master-port 29602 bench.py --gpus 8 --steps 10 --warmup 3 --no-extras --graph 0 > gpurun_out/s8_eager.json 2> gpurun_out/s8_eager.err
NCCL_MAX_CTAS=4 $R --master-port 29603 bench.py --gpus 8 --steps 10 --warmup 3 --no-extras > gpurun_out/s8_cta4.json 2> gpurun_out/s8_cta4.err
NCCL_MAX_CTAS=16 $R --master-port 29604 bench.py --gpus 8 --steps 10 --warmup 3 --no-extras > gpurun_out/s8_cta16.json 2> gpurun_out/s8_cta16.err
TVT_BUCKET_MB=128 $R --master-port 29605 bench.py --gpus 8 --steps 10 --warmup 3 --no-extras > gpurun_out/s8_b128.json 2> gpurun_out/s8_b128.err
for f in n1 default eager cta4 cta16 b128; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s8_$f.json").read().strip().splitlines()[-1])
    print("$f", d["n_gpus"], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d.get("ranks_hold_identical_parameters"), d["config"]["launch_mode"][:12])
except Exception as e:
    print("$f", "FAILED", e)
PY
done
