"""The C-ABI shared library builds for sm_100a, loads without a GPU and exports every symbol that
include/memvul_b200.h declares; argument validation answers before any CUDA work."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported(native_lib):
    hdr = open(os.path.join(ROOT, "include", "memvul_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(memvul_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(native_lib, name), f"{name} declared in the header but not exported"
    from memvul_b200 import native
    assert sorted(native.EXPORTS) == declared


def test_library_is_sm100a_with_tcgen05_and_tma():
    from memvul_b200 import native
    native.build()
    sass = subprocess.run(["cuobjdump", "-sass", native.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in subprocess.run(["cuobjdump", "-lelf", native.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "STTM", "UTMALDG"):       # tcgen05.mma / tcgen05.ld / .st / TMA (B200_PROFILING.md)
        assert mnemonic in sass, mnemonic
    assert "HMMA.16" not in sass                                   # no legacy mma.sync path


def test_attention_register_pool_balances():
    """The three-CTA attention kernel re-partitions registers with setmaxnreg (attention_tcgen05_v3.cuh): the soft-max
    warpgroup rises to 120 and the other drops to 40 from the launch value.  `setmaxnreg.inc` blocks until the CTA's pool
    holds enough registers, so the kernel only terminates if ptxas really launches it at (120 + 40) / 2 = 80 registers per
    thread -- and three CTAs only fit if 6 warps x 80 x 32 stays inside a scheduler's 16,384 registers."""
    from memvul_b200 import native
    native.build()
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", native.LIB_PATH], capture_output=True, text=True).stdout
    lines = res.splitlines()
    regs = [int(re.search(r"REG:(\d+)", lines[i + 1]).group(1)) for i, l in enumerate(lines)
            if "attention_tcgen05_v3_kernel" in l and i + 1 < len(lines)]
    assert regs and all(r == 80 for r in regs), regs
    src = open(os.path.join(os.path.dirname(native.LIB_PATH), "csrc", "attention_tcgen05_v3.cuh")).read()
    m = re.search(r"REGS_LAUNCH = (\d+), REGS_SOFTMAX = (\d+), REGS_AUX = (\d+)", src)
    launch, soft, aux = (int(x) for x in m.groups())
    assert launch == 80 and soft + aux == 2 * launch
    assert (256 // 32) * 3 // 4 * launch * 32 <= 16384             # 6 warps per scheduler at the launch value
    sass = subprocess.run(["cuobjdump", "-sass", native.LIB_PATH], capture_output=True, text=True).stdout
    assert f"USETMAXREG.TRY_ALLOC.CTAPOOL UP0, {hex(soft)}" in sass and f"USETMAXREG.DEALLOC.CTAPOOL {hex(aux)}" in sass


def test_argument_validation_without_gpu(native_lib):
    L = native_lib
    assert L.memvul_abi_version() == 3
    assert L.memvul_gemm_f16(None, None, None, None, None, 128, 100, 64, 0, None) == -1
    assert b"N % 128" in L.memvul_last_error()
    assert L.memvul_gemm_f16(None, None, None, None, None, 128, 128, 60, 0, None) == -1
    assert L.memvul_attention_f16(None, None, None, None, 1, 128, 768, None) == -1
    assert L.memvul_pool_match(None, 0, None, None, None, None, None, None, None, 4, 0, 768, 512, 0,
                               None, None, None, None, None, None, None, None, 31, None) == -1
    assert b"non-empty bank" in L.memvul_last_error()
    assert L.memvul_pool_match(None, 0, None, None, None, None, None, None, None, 4, 3, 768, 512, 2,
                               None, None, None, None, None, None, None, None, 1, None) == -1
    assert L.memvul_bank_prepare(None, None, 0, 512, None, None) == -1
    assert L.memvul_attention_f32(None, None, None, None, 1, 128, 768, None) == -1
    assert L.memvul_split3_f16(None, None, 4, 128, 0, None) == -1
    assert L.memvul_launch_count() == 0


def test_workspace_size_formula(native_lib):
    import torch
    from memvul_b200 import native
    from memvul_b200.synthetic import BERT_TINY, EMB, synthetic_state_dict
    w = native.PackedBert(synthetic_state_dict(BERT_TINY), EMB, torch.device("cpu"))
    M, H, I = 4 * 128, 128, 512
    up = lambda x: (x + 1023) // 1024 * 1024
    tail = up(4 * H * 4) + up(4 * H * 2) + up(4 * H * 2) + up(4 * I * 2)        # [B,*] rows of the CLS-only last layer
    base = up(M * H * 2) + up(M * 3 * H * 2) + up(M * H * 2) + up(M * I * 2) + tail
    assert w.workspace_bytes(4, 128) == base and w.workspace_bytes(4, 128, native.ENC_PACKED | native.ENC_CLS_ONLY) == base
    assert w.workspace_bytes(4, 128, native.ENC_PACKED) == base + up(M * H * 4)     # packed residual stream to unpack
    assert w.hidden == 128 and w.layers == 2 and w.heads == 2 and w.intermediate == 512
    # accuracy mode: fp32 residual + split operand [M,3H] + fp32 qkv / ctx / FFN intermediate + split GELU output [M,3I]
    precise = up(M * H * 4) + up(M * 3 * H * 2) + up(M * 3 * H * 4) + up(M * H * 4) + up(M * I * 4) + up(M * 3 * I * 2)
    assert w.workspace_bytes(4, 128, native.ENC_PRECISE | native.ENC_PACKED | native.ENC_CLS_ONLY) == precise


def test_split_weight_layout():
    """[W_hi | W_hi | W_lo] against [A_hi | A_lo | A_hi]: the K-concatenated product equals the three significant partial
    products, and hi + lo restores the fp32 value to ~2^-22."""
    import torch
    from memvul_b200 import native
    torch.manual_seed(0)
    w = torch.randn(16, 64) * 0.05
    a = torch.randn(8, 64)
    w3 = native.split3_weight(w)
    assert w3.shape == (16, 192) and w3.dtype == torch.float16
    hi, hi2, lo = w3[:, :64].float(), w3[:, 64:128].float(), w3[:, 128:].float()
    assert torch.equal(hi, hi2)
    assert float(((hi + lo) - w).abs().max()) <= float(w.abs().max()) * 2.0 ** -21
    a_hi = a.half().float()
    a_lo = (a - a_hi).half().float()
    a3 = torch.cat([a_hi, a_lo, a_hi], 1)
    ref = a.double() @ w.double().T
    got = a3.double() @ w3.double().T
    plain = a_hi.double() @ hi.double().T
    assert float((got - ref).abs().max()) < 2e-6 < float((plain - ref).abs().max())


def test_no_cpu_fallback():
    """The product path must fail loudly off-GPU (no oracle / CPU fallback)."""
    import torch
    from memvul_b200 import native
    from memvul_b200.synthetic import BERT_TINY, build_memory_model, synthetic_ids
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    model, _ = build_memory_model(BERT_TINY)
    ids, mask, t = synthetic_ids(2, 8, vocab_size=1024)
    with pytest.raises(native.NativeError):
        model.forward_gold_instances({"tokens": {"token_ids": ids, "mask": mask, "type_ids": t}},
                                     [{"type": "golden", "instance": [{"label": "x"}]}] * 2)
    src = "".join(open(os.path.join(ROOT, "memvul_b200", f)).read() for f in os.listdir(os.path.join(ROOT, "memvul_b200"))
                  if f.endswith(".py"))
    assert "oracle" not in src.replace("the CPU oracle", "").replace("CPU oracle", ""), "product code must not reference oracle/"
