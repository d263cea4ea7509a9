"""CPU: the C-ABI shared library builds, loads, and exports every entry point include/fedcola_b200.h declares
(no compute calls — there is no GPU here); ctypes struct layouts match the header; host-side model plumbing."""
import copy
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "fedcola_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from fedcola_b200 import _lib
    L = _lib.lib()
    names = declared_functions()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert L.fc_abi_version() == 1
    assert int(L.fc_aggregate_tile_floats()) % 4 == 0 and int(L.fc_chunk_floats()) % 4 == 0


def test_struct_layout_handshake():
    from fedcola_b200 import runtime as R
    L = R._sigs()
    assert L.fc_sizeof_mat_desc() == ctypes.sizeof(R.MatDesc)
    assert L.fc_sizeof_step_args() == ctypes.sizeof(R.StepArgs)
    assert R.CHUNK_DT.itemsize == 16 and R.PREP_DT.itemsize == 56 and R.AUX_DT.itemsize == 40


def test_error_reporting_without_gpu():
    """Entry points validate arguments before touching the device and report through fc_last_error()."""
    from fedcola_b200 import _lib
    L = _lib.lib()
    rc = L.fc_gemm_bf16(0, 0, 0, None, ctypes.c_longlong(0), 0, None, ctypes.c_longlong(0), 0, 0, None, None,
                        ctypes.c_longlong(0), None, None, None, 0, None, None, 0, ctypes.c_float(1.0), 1, 0, None)
    assert rc == -1
    assert b"empty problem" in L.fc_last_error()
    with pytest.raises(RuntimeError):
        _lib.check(rc, "fc_gemm_bf16")


def test_product_refuses_cpu():
    from fedcola_b200.models import mome
    from fedcola_b200.harness import make_args
    args = make_args(vocab_size=512, seq_len=16)
    m = mome.create_model("mome_d64_l2", pretrained=False, num_classes=[100, None], modalities=["img", None], args=args,
                          tasks=["cls", None])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m([torch.zeros(2, 3, 224, 224), None])


def test_model_mirror_state_plumbing():
    from fedcola_b200.models import mome
    from fedcola_b200.harness import make_args
    args = make_args(vocab_size=512, seq_len=16, shared_param="attn", share_scope="all")
    kw = dict(pretrained=False, args=args, with_aux=True, aux_trained=False)
    m = mome.create_model("mome_d64_l2", num_classes=[None, 4], modalities=[None, "txt"], tasks=[None, "cls"], **kw)
    sd = m.state_dict()
    # share_scope == 'all': the None encoder aliases the text blocks (mome.py:824-827)
    assert sd["blockses.0.0.attn.qkv.weight"].data_ptr() == sd["blockses.1.0.attn.qkv.weight"].data_ptr()
    assert not any("blockses.0" in k for k in m.required_params())
    assert all("aux_weight" in k for k in m.aux_params())
    assert m.get_parameter("blockses.1.0.attn.qkv.aux_weight").requires_grad is False
    assert torch.equal(sd["blockses.1.0.attn.qkv.aux_weight"], sd["blockses.1.0.attn.qkv.weight"])   # build_aux
    assert float(sd["blockses.1.0.attn.qkv.cross_modal_scale"]) == 0.0
    # parameters are views of ONE arena; deepcopy / load_state_dict keep that property
    m2 = copy.deepcopy(m)
    lo, hi = m2.arena.data_ptr(), m2.arena.data_ptr() + m2.arena.numel() * 4
    assert all(lo <= p.data_ptr() < hi for p in m2.parameters())
    m2.load_state_dict({k: v + 1 for k, v in sd.items()})
    assert torch.equal(m2.state_dict()["norm.bias"], sd["norm.bias"] + 1)
    assert len(list(m.named_parameters())) == len(list(m2.named_parameters()))
    for p in m2.parameters():
        p.requires_grad = False
    m3 = copy.deepcopy(m2)
    assert not any(p.requires_grad for p in m3.parameters())


def test_install_as_src_registers_drop_in_modules():
    import sys
    import fedcola_b200
    names = fedcola_b200.install_as_src()
    try:
        from importlib import import_module
        for alg in ("fedavg", "fedprox", "fediot"):
            assert hasattr(import_module(f"src.server.{alg}server"), f"{alg.title()}Server")
            assert hasattr(import_module(f"src.client.{alg}client"), f"{alg.title()}Client")
            assert hasattr(import_module(f"src.algorithm.{alg}"), f"{alg.title()}Optimizer")
    finally:
        for n in names:
            sys.modules.pop(n, None)
