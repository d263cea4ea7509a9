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
        # timm.create_model('mome_*') — what FedavgServer._init_model calls (fedavgserver.py:151-155) — lands here
        import timm
        from fedcola_b200.harness import make_args
        from fedcola_b200.models import mome
        args = make_args(vocab_size=512, seq_len=16)
        kw = dict(pretrained=False, num_classes=[100, None], modalities=["img", None], args=args, tasks=["cls", None])
        for name in ("mome_small_patch16", "mome_tiny_patch16", "mome_small_patch16_224_in21k",
                     "mome_base_patch16_224_ours", "mome_toy_patch16_224"):      # mome.py:924-1033
            assert name in mome._REGISTRY
        m = timm.create_model("mome_toy_patch16_224", **kw)
        assert isinstance(m, mome.ModalityAgnosticTransformer) and m.embed_dim == 4
        with pytest.raises(RuntimeError, match="Unknown model"):
            timm.create_model("no_such_model_xyz")
    finally:
        for n in names:
            if n != "timm.create_model":
                sys.modules.pop(n, None)
        sys.modules.pop("timm", None)


def test_mp_flag_is_refused_loudly():
    """--mp (ProcessPoolExecutor clients, fedavgserver.py:560-562) has no equivalent here: raise, do not ignore."""
    from fedcola_b200.harness import make_args
    from fedcola_b200.server.fedavgserver import FedavgServer
    with pytest.raises(NotImplementedError, match="--mp"):
        FedavgServer(args=make_args(mp=True), writer=None, server_dataset=(None, {}), client_datasets=[],
                     model_str="mome_d64_l2")


def test_placement_rules():
    from fedcola_b200 import aggregation as agg
    from fedcola_b200.arena import MatSpec
    assert agg.place_clients([0] * 5, 2, "reference") == [0, 1, 0, 1, 0]          # cuda:(i % ngpu), :310-311
    # BASELINE configs[2]: 6 img + 6 txt + 4 img-txt ViT-S clients over 8 ranks
    sp = MatSpec(embed_dim=384, depth=12, num_heads=6, modalities=("img", "txt"), num_classes=(None, None),
                 tasks=("rtv", "rtv"), vocab_size=30522, max_text_len=64)
    f = {m: agg.train_flops_per_sample(sp, m) for m in ("img", "txt", "img+txt")}
    assert abs(f["img"] / 27.59e9 - 1) < 0.01 and abs(f["txt"] / 8.38e9 - 1) < 0.01 and abs(f["img+txt"] / 35.97e9 - 1) < 0.01
    costs = [f["img"]] * 6 + [f["txt"]] * 6 + [f["img+txt"]] * 4

    def spread(rule):
        slots = agg.place_clients(costs, 8, rule)
        load = [sum(c for c, s in zip(costs, slots) if s == r) for r in range(8)]
        return max(load) / (sum(load) / 8)
    assert spread("balanced") < spread("reference") and spread("balanced") < 1.35
    assert agg.place_clients(costs, 8, "balanced") == agg.place_clients(costs, 8, "balanced")   # deterministic
    with pytest.raises(ValueError):
        agg.place_clients(costs, 8, "nope")
