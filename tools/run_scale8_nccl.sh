# 8 x B200: does the gradient all-reduce's protocol / algorithm change what it costs the step?  (NCCL picks RING_LL by itself.)
set -x
cd $GRAFT_REPO_ROOT
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
A="bench.py --gpus 8 --steps 15 --warmup 3 --no-extras"
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING $R --master-port 29621 $A > gpurun_out/n8_auto.json 2> gpurun_out/n8_auto.err
NCCL_PROTO=Simple $R --master-port 29622 $A > gpurun_out/n8_simple.json 2> gpurun_out/n8_simple.err
NCCL_PROTO=LL128 $R --master-port 29623 $A > gpurun_out/n8_ll128.json 2> gpurun_out/n8_ll128.err
NCCL_ALGO=NVLS $R --master-port 29624 $A > gpurun_out/n8_nvls.json 2> gpurun_out/n8_nvls.err
NCCL_ALGO=Tree $R --master-port 29625 $A > gpurun_out/n8_tree.json 2> gpurun_out/n8_tree.err
python - <<PY
import json
for name in ("auto", "simple", "ll128", "nvls", "tree"):
    try:
        d = json.loads(open(f"gpurun_out/n8_{name}.json").read().strip().splitlines()[-1])
        print(name, d["n_gpus"], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["clocks"]["sm_mhz"], d.get("ranks_hold_identical_parameters"))
    except Exception as e:
        print(name, "FAILED", repr(e)[:200])
PY
grep -h -i "nvls\|AllReduce.*->\|Algo\|proto" gpurun_out/n8_auto.err | head -30
