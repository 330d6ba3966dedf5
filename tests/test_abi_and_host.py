"""CPU-side checks: the C-ABI library loads and exports every symbol include/nasb200.h declares; the host mirror
reproduces the reference's state_dict keys / registry names; and there is NO CPU fallback."""
import os
import re

import numpy as np
import pytest
import torch

import nas_segm_b200
from golden_util import NETS, keys_shapes
from nas_segm_b200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "nasb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nasb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    syms = _header_symbols()
    assert len(syms) >= 36
    handle = lib.load()
    for s in syms:
        assert hasattr(handle, s), s
    assert set(syms) == set(lib.EXPORTS)  # the ctypes table covers the header, nothing more, nothing less
    assert lib.version().startswith("nasb200") and "sm_100a" in lib.version()


def test_library_is_sm100a_only_and_torch_free():
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs
    ldd = subprocess.run(["ldd", lib.LIB_PATH], capture_output=True, text=True).stdout
    names = [ln.split("=>")[0].split("(")[0].strip() for ln in ldd.splitlines()]  # library names only (no load addresses)
    assert names and not any("torch" in n or "c10" in n for n in names), names


def test_registry_names_match_reference(golden):
    from nas_segm_b200.nn.layer_factory import AGG_OPS, OPS
    from nas_segm_b200.rl.genotypes import AGG_OP_NAMES, OP_NAMES, OP_NAMES_WACV
    fx = golden("ops")
    ref_names = {str(fx[k][0]) for k in fx.files if k.endswith("/meta") and k.startswith("op")}
    assert set(OPS) == ref_names and len(OPS) == 16
    assert set(AGG_OPS) == {"psum", "cat"}
    assert len(OP_NAMES) == 11 and len(OP_NAMES_WACV) == 6 and AGG_OP_NAMES == ["psum", "cat"]
    assert all(n in OPS for n in OP_NAMES + OP_NAMES_WACV)


def test_registry_state_dict_keys(golden):
    from nas_segm_b200.nn.layer_factory import AGG_OPS, OPS
    fx = golden("ops")
    for tag in sorted({k.split("/")[0] for k in fx.files}):
        meta = [str(v) for v in fx[tag + "/meta"]]
        if tag.startswith("op"):
            m = OPS[meta[0]](int(meta[1]), int(meta[2]), int(meta[3]), True, int(meta[4]))
        else:
            m = AGG_OPS[meta[0]](int(meta[1]), int(meta[2]), int(meta[3]), True, repeats=1, larger=bool(int(meta[4])))
        mine = {(k, tuple(v.shape)) for k, v in m.state_dict().items()}
        assert mine == set(keys_shapes(fx, tag + "/")), tag


@pytest.mark.parametrize("tag", sorted(NETS))
def test_network_state_dict_keys_and_info(golden, tag):
    from nas_segm_b200.nn.encoders import mbv2
    from nas_segm_b200.nn.micro_decoders import MicroDecoder, TemplateDecoder
    fx = golden("net_" + tag)
    paper, cfg, ncls, agg, rep, aux = NETS[tag]
    if paper == "wacv":
        enc = mbv2(return_layers=[1, 2])
        dec = TemplateDecoder(list(enc.out_sizes), ncls, cfg, agg_size=agg, repeats=rep)
    else:
        enc = mbv2()
        sizes = enc.out_sizes
        dec = MicroDecoder(sizes, ncls, cfg, agg_size=agg, aux_cell=aux, repeats=rep)
        assert sizes == [agg] * 4  # the reference overwrites the caller's list (micro_decoders.py:184)
    mine = {("encoder." + k, tuple(v.shape)) for k, v in enc.state_dict().items()}
    mine |= {("decoder." + k, tuple(v.shape)) for k, v in dec.state_dict().items()}
    assert mine == set(keys_shapes(fx))
    assert dec.info == str(fx["info"])
    n_params = sum(p.numel() for p in enc.parameters()) + sum(p.numel() for p in dec.parameters())
    assert n_params == int(fx["n_params"])
    # BN freezing scans must still find every BatchNorm2d (trainer.py:124-127)
    n_bn = sum(isinstance(m, torch.nn.BatchNorm2d) for m in dec.modules())
    assert n_bn == sum(1 for k, _ in keys_shapes(fx) if k.startswith("decoder.") and k.endswith("running_mean"))


def test_no_cpu_fallback():
    from nas_segm_b200.helpers import miou_utils
    from nas_segm_b200.nn.layer_factory import OPS
    m = OPS["sep_conv_3x3"](8, 8, 1, True, 1)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 8, 4, 4))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            miou_utils.fast_cm(np.zeros(4, np.uint8), np.zeros(4, np.uint8), 2)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "nas-segm-pytorch_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "/root/reference" not in src, f


def test_dropin_names():
    nas_segm_b200.dropin()
    from engine.inference import validate  # noqa: F401
    from engine.trainer import populate_task0, train_segmenter, train_task0  # noqa: F401
    from helpers.miou_utils import compute_iu, compute_ius_accs, fast_cm  # noqa: F401
    from nn.encoders import create_encoder, mbv2  # noqa: F401
    from nn.layer_factory import AGG_OPS, OPS  # noqa: F401
    from nn.micro_decoders import MicroDecoder, TemplateDecoder  # noqa: F401
    import inspect
    sig = inspect.signature(train_task0.__wrapped__)
    assert list(sig.parameters) == ["Xy_train", "segmenter", "optim_dec", "epoch", "segm_crit", "kd_crit", "batch_size",
                                    "freeze_bn", "do_kd", "kd_coeff", "dec_grad_clip", "do_polyak", "avg_param",
                                    "polyak_decay", "aux_weight"]
    sig = inspect.signature(validate.__wrapped__)
    assert list(sig.parameters) == ["segmenter", "val_loader", "epoch", "epoch2", "num_classes", "print_every",
                                    "omit_classes"]
    for m in ("nn", "rl", "engine", "helpers", "nn.layer_factory", "nn.micro_decoders", "nn.encoders", "rl.genotypes",
              "helpers.utils", "helpers.miou_utils", "engine.trainer", "engine.inference"):
        import sys
        sys.modules.pop(m, None)


def test_try_except_convention():
    from nas_segm_b200.helpers.utils import try_except

    @try_except
    def boom(kind):
        raise kind("x")

    assert boom(RuntimeError) == 0
    with pytest.raises(ValueError):
        boom(ValueError)


def test_zero_arena_host_logic():
    """The per-iteration zero arena (lib.zero_arena): inactive outside begin()/end(); the first iteration of a model is
    served by torch.zeros while the demand is recorded; later iterations hand out disjoint, 16-byte aligned, zeroed views of
    ONE buffer that is never reallocated (captured graphs keep its address); begin() re-zeroes what the previous
    iteration wrote; demand beyond the cleared prefix falls back to torch.zeros."""
    import torch
    from nas_segm_b200 import lib
    ar = lib._ZeroArena()
    ar.CAPACITY = 1 << 16
    cpu = torch.device("cpu")
    t = ar.take((3, 5), torch.float32, cpu)
    assert t.shape == (3, 5) and float(t.abs().sum()) == 0 and not ar.active

    def iteration(extra=0):
        ar.begin(cpu)
        outs = [ar.take((7,), torch.float32, cpu), ar.take(10, torch.float64, cpu), ar.take((4, 3, 1, 1), torch.float32, cpu)]
        if extra:
            outs.append(ar.take(extra, torch.float32, cpu))
        for o in outs:
            assert float(o.abs().sum()) == 0
            o += 1.0  # what the kernels do: accumulate
        ar.end()
        return outs

    first = iteration()
    buf = ar.bufs[cpu]
    base, size = buf.data_ptr(), buf.numel()
    assert all(not (base <= o.data_ptr() < base + size) for o in first)      # demand was unknown: torch.zeros
    second = iteration()
    assert all(base <= o.data_ptr() < base + size and o.data_ptr() % 16 == 0 for o in second)
    spans = sorted((o.data_ptr(), o.data_ptr() + o.numel() * o.element_size()) for o in second)
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))               # disjoint
    assert second[1].dtype == torch.float64 and second[2].shape == (4, 3, 1, 1)
    third = iteration()
    assert ar.bufs[cpu] is buf and [o.data_ptr() for o in third] == [o.data_ptr() for o in second]  # same addresses, re-zeroed
    big = iteration(extra=1 << 20)                                           # more than CAPACITY: falls back, never reallocates
    assert ar.bufs[cpu] is buf and not (base <= big[-1].data_ptr() < base + size) and float(big[-1].sum()) == (1 << 20)
    again = iteration()
    assert all(base <= o.data_ptr() < base + size for o in again)
    assert float(ar.take(4, torch.float32, cpu).sum()) == 0 and not ar.active  # outside an iteration: plain zeros


def test_abi_host_side_contracts_without_a_gpu():
    """Entry points that are pure host functions (shape envelopes, workspace sizes) and the argument validation of the
    launchers: bad arguments come back as NASB_ERR_BAD_ARG (10002) / NASB_ERR_UNSUPPORTED (10001) before any CUDA call, so
    this runs on a box without a device (no compute calls)."""
    import ctypes as C
    h = lib.load()
    BAD, UNS = 10002, 10001
    # shape envelopes of the tensor-core paths (include/nasb200.h)
    assert h.nasb_pw_tc_supported(32, 32) == 1 and h.nasb_pw_tc_supported(448, 320) == 1 and h.nasb_pw_tc_supported(960, 320) == 1
    assert h.nasb_pw_tc_supported(7, 32) == 0 and h.nasb_pw_tc_supported(32, 12) == 0 and h.nasb_pw_tc_supported(32, 8192) == 0 and h.nasb_pw_tc_supported(4104, 32) == 0
    assert h.nasb_pw_tc_wgrad_supported(192, 32) == 1 and h.nasb_pw_tc_wgrad_supported(32, 960) == 1 and h.nasb_pw_tc_wgrad_supported(32, 260) == 0
    assert h.nasb_conv3_tc_supported(64, 19) == 1
    assert h.nasb_pack_conv3_elems(19, 64, 0) > 0
    assert h.nasb_bn_stats_workspace(64) == 64 * (2 * 8 + 2 * 4) and h.nasb_loss_workspace() > 0
    # argument validation
    t32 = lib.NasbTensor(0x1000, 1, 4, 4, 8, 8, lib.F32)
    t16 = lib.NasbTensor(0x1000, 1, 4, 4, 8, 8, lib.BF16)
    img = lib.NasbTensor(0x1000, 1, 8, 8, 3, 3, lib.F32_NCHW)
    r = C.byref
    assert h.nasb_stem_im2col(None, 3, 2, 1, 1, None, None) == UNS
    assert h.nasb_stem_im2col(r(img), 3, 2, 1, 1, r(t16), None) == UNS            # patch matrix must have 32 channels
    assert h.nasb_bn_finalize_affine_act(None, 16, r(t16), None, None, 1e-5, 0.1, None, None, None, None, None, None, None, 0,
                                         None, r(t16), None) == BAD
    assert h.nasb_dwconv_tile(r(t32), C.c_void_p(0x1000), 3, 1, 1, 1, 0, None, None, 0, r(t32), None, None) == UNS  # bf16 only
    assert h.nasb_dwconv_tile(None, None, 3, 1, 1, 1, 0, None, None, 0, None, None, None) == BAD
    assert h.nasb_dwconv_wgrad_tile(r(t16), r(t16), 7, 1, 1, 3, C.c_void_p(0x1000), None) == UNS                    # k in {3,5}
    assert h.nasb_pw_tc_fwd(r(t32), C.c_void_p(0x1000), 8, None, None, 0, None, r(t32), None, None) == BAD          # bf16 only
    assert h.nasb_pw_tc_wgrad(None, None, None, None) == BAD
    assert h.nasb_conv3_tc_fwd(r(t32), C.c_void_p(0x1000), 8, 1, 1, None, None, 0, r(t32), None, None) == UNS
    assert h.nasb_bn_act_bwd(None, None, None, 0, None, None, None, None, None, None, 1, None, None, None, None, None) == BAD
    assert h.nasb_confmat_labels(None, None, 16, 300, None, None) == BAD                                            # C <= 256
    assert h.nasb_affine_act(r(t16), None, None, 0, r(t32), None) == BAD                                            # dtype mismatch
