# 8 x B200: static tile order vs dynamic tile claims in the CTA-pair GEMMs (TVT_GEMM_DYNAMIC), same box, back to back
set -x
cd $GRAFT_REPO_ROOT
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
A="bench.py --gpus 8 --steps 20 --warmup 3 --no-extras"
TVT_GEMM_DYNAMIC=1 $R --master-port 29631 $A > gpurun_out/y8_dyn.json 2> gpurun_out/y8_dyn.err
TVT_GEMM_DYNAMIC=0 $R --master-port 29632 $A > gpurun_out/y8_static.json 2> gpurun_out/y8_static.err
python - <<PY
import json
for name in ("dyn", "static"):
    try:
        d = json.loads(open(f"gpurun_out/y8_{name}.json").read().strip().splitlines()[-1])
        print(name, d["n_gpus"], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["clocks"]["sm_mhz"], d.get("ranks_hold_identical_parameters"))
    except Exception as e:
        print(name, "FAILED", repr(e)[:200])
PY
