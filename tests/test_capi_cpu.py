"""CPU: the C-ABI library loads, exports every symbol include/tvt.h declares, its ctypes mirrors have the
C compiler's struct layout, and argument validation answers TVT_EINVAL without touching a GPU."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tvt.h")


@pytest.fixture(scope="module")
def capi():
    import tvt_b200
    from tvt_b200 import build
    build.build()
    return tvt_b200.capi


def _declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"TVT_API\s+[\w\s\*]+?\b(tvt_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol(capi):
    lib = capi.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/tvt.h but not exported"
    assert set(capi.ENTRY_POINTS) | set(capi.PLAIN_SYMBOLS) == set(names)
    assert lib.tvt_version() >= 100


def test_ctypes_structs_match_c_layout(capi):
    structs = {st.__name__: st for st in capi.ENTRY_POINTS.values()}
    cnames = {name: "tvt_" + re.sub(r"(?<!^)(?=[A-Z])", "_", name[:-4]).lower() + "_args" for name in structs}
    # irregular names
    fix = {"LayerNormFwdArgs": "tvt_layernorm_fwd_args", "LayerNormBwdArgs": "tvt_layernorm_bwd_args",
           "SpatialPoolArgs": "tvt_spatial_pool_args", "SplitArgs": "tvt_split_args", "ColsumArgs": "tvt_colsum_args",
           "PosencArgs": "tvt_posenc_args", "L2NormArgs": "tvt_l2norm_args", "Split3Args": "tvt_split3_args"}
    cnames.update(fix)
    body = "\n".join(f'  printf("{py} %zu\\n", sizeof({c}));' for py, c in cnames.items())
    prog = f'#include <stdio.h>\n#include "tvt.h"\nint main(void) {{\n{body}\n  return 0;\n}}\n'
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "sz.c")
        open(src, "w").write(prog)
        exe = os.path.join(td, "sz")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    for line in out.strip().splitlines():
        name, size = line.split()
        assert ctypes.sizeof(structs[name]) == int(size), f"{name}: ctypes {ctypes.sizeof(structs[name])} != C {size}"


def test_validation_needs_no_gpu(capi):
    lib = capi.load()
    g = capi.GemmArgs()
    assert lib.tvt_gemm(ctypes.byref(g), None) == -1            # TVT_EINVAL: null operands
    assert b"null" in lib.tvt_last_error()
    g.a, g.b, g.m, g.n, g.k, g.lda, g.ldb, g.splits = 16, 16, 8, 15, 64, 64, 64, 1
    assert lib.tvt_gemm(ctypes.byref(g), None) == -1
    assert b"multiple of 8" in lib.tvt_last_error()
    a = capi.AttentionFwdArgs()
    a.q = a.k = a.v = a.o = 16
    a.batch, a.heads, a.sq, a.sk, a.head_dim = 1, 1, 4, 4, 12
    assert lib.tvt_attention_fwd(ctypes.byref(a), None) == -1
    assert b"head_dim" in lib.tvt_last_error()
    ln = capi.LayerNormFwdArgs()
    ln.x = ln.y = ln.gamma = ln.beta = 16
    ln.rows, ln.d = 4, 12
    assert lib.tvt_layernorm_fwd(ctypes.byref(ln), None) == -1
    # blocked LayerNorm output: y_seq must divide rows and y_pitch must hold y_seq rows
    ln.rows, ln.d, ln.y_seq, ln.y_pitch = 12, 64, 5, 1024
    assert lib.tvt_layernorm_fwd(ctypes.byref(ln), None) == -1 and b"y_seq" in lib.tvt_last_error()
    ln.y_seq, ln.y_pitch = 4, 128
    assert lib.tvt_layernorm_fwd(ctypes.byref(ln), None) == -1 and b"y_seq" in lib.tvt_last_error()
    # fused bias gradient: only for accumulating (atomic_out) wgrad-oriented GEMMs
    g = capi.GemmArgs()
    g.a, g.b, g.m, g.n, g.k, g.lda, g.ldb, g.splits = 16, 16, 512, 512, 4096, 512, 512, 1
    g.a_mn_major = g.b_mn_major = 1
    g.out_f32, g.ld_f32, g.a_rowsum = 16, 512, 16
    assert lib.tvt_gemm(ctypes.byref(g), None) == -1 and b"a_rowsum" in lib.tvt_last_error()
    assert lib.tvt_gemm_rowsum_supported(0, 512, 4096, 1) == 0 and lib.tvt_gemm_rowsum_supported(512, 512, 64, 2) == 0


def test_product_path_fails_loudly_without_cuda(capi):
    """No CPU fallback: CPU tensors are rejected by the tensor-level wrappers."""
    import torch
    from tvt_b200 import TvtError, ops
    x = torch.zeros(8, 64)
    with pytest.raises(TvtError, match="CUDA"):
        ops.layernorm_fwd(x, torch.ones(64), torch.zeros(64))
    from tvt_b200 import hostapi
    m = hostapi.SimpleTransformer(batch_size=2, seq_len=4, cls=1, dropout=0.0, input_dimension=64, nhead=2, nhid=32,
                                  nlayers=1, learning_rate=1e-3, momentum=0.0, weight_decay=0.0)
    with pytest.raises(TvtError, match="CUDA"):
        m.ptn(torch.zeros(2, 4, 1, 64))


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "data-efficient-video-transformers_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f"{f} imports the oracle"


def test_state_dict_keys_match_oracle():
    import torch
    from oracle import param
    from tvt_b200 import hostapi
    cfg = dict(batch_size=2, seq_len=4, cls=1, dropout=0.0, input_dimension=64, nhead=2, nhid=32, nlayers=2,
               learning_rate=1e-3, momentum=0.0, weight_decay=0.0, model="ptn")
    assert list(hostapi.SimpleTransformer(**cfg).state_dict().keys()) == list(param.SimpleTransformer(**cfg).state_dict().keys())
    kw = dict(in_dims=(64, 32, 16), d=32, nhead=2, nhid=64, nlayers=1, batch_size=2, frames=8, fusion="cross", pyramid=True)
    assert list(hostapi.FusionTransformer(**kw).state_dict().keys()) == list(param.FusionTransformer(**kw).state_dict().keys())
    assert list(hostapi.FrameStream(d=64, nhead=2, nhid=32, nlayers=1, seq_len=6).state_dict().keys()) == \
        list(param.FrameStream(d=64, nhead=2, nhid=32, nlayers=1, seq_len=6).state_dict().keys())
    assert list(hostapi.Reasoning(1, 9, 15, 32).state_dict().keys()) == list(param.Reasoning(1, 9, 15, 32).state_dict().keys())
    # same RNG stream => same initial weights as the reference-order construction
    torch.manual_seed(1130); a = hostapi.SimpleTransformer(**cfg)
    torch.manual_seed(1130); b = param.SimpleTransformer(**cfg)
    for (k, va), (_, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(va, vb), k


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the CPU oracle arm of the bench contract) on the CPU-runnable C1 config: stdout is exactly
    one JSON line with the contract's keys, and nothing of the product path is needed for it."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train clips/sec fwd+bwd" and d["unit"] == "clips/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
