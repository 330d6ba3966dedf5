"""Per-entry-point microbenchmark (GPU box; compare variants with the NASB_* switches listed in DESIGN.md section 4): times individual C-ABI calls at the shapes that dominate the arch0 training
iteration (batch 8 @2048x1024, bf16) with CUDA events, rotating over enough buffer sets that the inputs never sit in L2.

    python tools/kbench.py [filter-substring ...]        # e.g.  python tools/kbench.py bn_act_bwd dwconv

Prints one line per case: ms, algorithmic GB/s (numel(inputs)+numel(outputs), the figure bench.py's roofline uses).
Used to compare kernel variants (environment switches) in one gpurun call; it is a tool, not a test.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch  # noqa: E402

import nas_segm_b200  # noqa: E402,F401
from nas_segm_b200 import lib  # noqa: E402
from nas_segm_b200.lib import call, desc, new_act, ptr, ref  # noqa: E402

DEV = torch.device("cuda", 0)
BF = torch.bfloat16
L2_BYTES = 256 << 20


def act(n, c, h, w, dtype=BF, fill="randn"):
    t = new_act(n, c, h, w, dtype, DEV)
    if fill == "randn":
        t.normal_()
    return t


def timeit(fn_sets, iters=None):
    """fn_sets: list of zero-arg callables (one per buffer set).  Returns average ms per call."""
    k = len(fn_sets)
    for f in fn_sets:
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn_sets[0]()
    e1.record()
    torch.cuda.synchronize()
    one = max(e0.elapsed_time(e1), 1e-3)
    iters = iters or int(min(60, max(6, 30.0 / one)))
    e0.record()
    for i in range(iters):
        fn_sets[i % k]()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def nsets(bytes_per_set):
    return int(max(2, min(8, L2_BYTES // max(bytes_per_set, 1) + 1)))


def fvec(c, lo=0.5, hi=1.5):
    return torch.empty(c, device=DEV).uniform_(lo, hi)


CASES = []


def case(name):
    def deco(f):
        CASES.append((name, f))
        return f
    return deco


def report(name, ms, nbytes):
    print("%-64s %8.4f ms %8.1f GB/s" % (name, ms, nbytes / (ms * 1e-3) / 1e9), flush=True)


BN_SHAPES = [(8, 144, 256, 512, 2), (8, 32, 512, 1024, 2), (8, 96, 512, 1024, 2), (8, 64, 256, 512, 1), (8, 192, 128, 256, 2),
             (8, 64, 128, 256, 1), (8, 32, 128, 256, 1), (8, 24, 256, 512, 0), (8, 64, 32, 64, 1)]


@case("bn_act_bwd")
def _():
    for n, c, h, w, a in BN_SHAPES:
        nb = 3 * n * c * h * w * 2
        sets = []
        for _ in range(nsets(nb)):
            dy, z, dz = act(n, c, h, w), act(n, c, h, w), act(n, c, h, w, fill=None)
            g, b, sc, sh, mu, rs = fvec(c), fvec(c, -0.5, 0.5), fvec(c), fvec(c, -0.5, 0.5), fvec(c, -0.2, 0.2), fvec(c)
            dg, db = torch.zeros(c, device=DEV), torch.zeros(c, device=DEV)
            ws = lib.workspace(DEV, 1 << 20)
            sets.append(lambda dy=dy, z=z, dz=dz, g=g, b=b, sc=sc, sh=sh, mu=mu, rs=rs, dg=dg, db=db, ws=ws: call(
                "nasb_bn_act_bwd", ref(desc(dy)), None, ref(desc(z)), a, ptr(g), ptr(b), ptr(sc), ptr(sh), ptr(mu), ptr(rs), 1,
                ptr(dg), ptr(db), ref(desc(dz)), ptr(ws)))
        report("bn_act_bwd[%dx%dx%dx%d act%d]" % (n, h, w, c, a), timeit(sets), nb)


@case("affine_act")
def _():
    for n, c, h, w, a in BN_SHAPES[:6]:
        nb = 2 * n * c * h * w * 2
        sets = []
        for _ in range(nsets(nb)):
            z, y = act(n, c, h, w), act(n, c, h, w, fill=None)
            sc, sh = fvec(c), fvec(c, -0.5, 0.5)
            sets.append(lambda z=z, y=y, sc=sc, sh=sh: call("nasb_affine_act", ref(desc(z)), ptr(sc), ptr(sh), a, ref(desc(y))))
        report("affine_act[%dx%dx%dx%d act%d]" % (n, h, w, c, a), timeit(sets), nb)


@case("bn_stats")
def _():
    for n, c, h, w, a in [(8, 144, 256, 512, 2), (8, 96, 256, 512, 2), (8, 32, 512, 1024, 2), (8, 192, 128, 256, 2), (8, 32, 128, 256, 1)]:
        nb = n * c * h * w * 2
        sets = []
        for _ in range(nsets(nb)):
            z = act(n, c, h, w)
            g, b = fvec(c), fvec(c)
            rm, rv, sm, sr, sc, sh = (torch.zeros(c, device=DEV) for _ in range(6))
            ws = lib.workspace(DEV, 1 << 20)
            sets.append(lambda z=z, g=g, b=b, rm=rm, rv=rv, sm=sm, sr=sr, sc=sc, sh=sh, ws=ws: call(
                "nasb_bn_stats", ref(desc(z)), ptr(g), ptr(b), 1e-5, 0.1, ptr(rm), ptr(rv), ptr(sm), ptr(sr), ptr(sc), ptr(sh), None, ptr(ws)))
        report("bn_stats[%dx%dx%dx%d]" % (n, h, w, c), timeit(sets), nb)


@case("bn_finalize_affine_act")
def _():
    for n, c, h, w, a in BN_SHAPES[:7]:
        nb = 2 * n * c * h * w * 2
        sets = []
        for _ in range(nsets(nb)):
            z, y = act(n, c, h, w), act(n, c, h, w, fill=None)
            P = n * h * w
            sums = torch.cat([torch.randn(c, dtype=torch.float64, device=DEV) * P * 0.1,
                              (torch.rand(c, dtype=torch.float64, device=DEV) + 1.0) * P])
            g, b = fvec(c), fvec(c, -0.5, 0.5)
            rm, rv, sm, sr, sc, sh = (torch.zeros(c, device=DEV) for _ in range(6))
            nbt = torch.zeros((), dtype=torch.int64, device=DEV)
            sets.append(lambda z=z, y=y, sums=sums, g=g, b=b, rm=rm, rv=rv, sm=sm, sr=sr, sc=sc, sh=sh, nbt=nbt, P=P: call(
                "nasb_bn_finalize_affine_act", ptr(sums), P, ref(desc(z)), ptr(g), ptr(b), 1e-5, 0.1, ptr(rm), ptr(rv), ptr(sm), ptr(sr),
                ptr(sc), ptr(sh), ptr(nbt), a, None, ref(desc(y))))
        report("bn_finalize_affine_act[%dx%dx%dx%d act%d]" % (n, h, w, c, a), timeit(sets), nb)


@case("pool3x3")
def _():
    for n, c, h, w, stride in [(8, 32, 128, 256, 1), (8, 48, 256, 512, 2), (8, 24, 256, 512, 1), (8, 64, 32, 64, 1)]:
        oh, ow = (h + 2 - 3) // stride + 1, (w + 2 - 3) // stride + 1
        x, o = act(n, c, h, w), act(n, c, oh, ow, fill=None)
        arg = torch.empty((n, oh, ow, c), dtype=torch.uint8, device=DEV)
        dy, dx = act(n, c, oh, ow), act(n, c, h, w, fill=None)
        nb = n * c * (h * w + oh * ow) * 2
        report("pool3x3_fwd[%dx%dx%dx%d s%d]" % (n, h, w, c, stride),
               timeit([lambda: call("nasb_pool3x3_fwd", ref(desc(x)), 0, stride, ref(desc(o)), ptr(arg))]), nb)
        report("pool3x3_bwd[%dx%dx%dx%d s%d]" % (n, h, w, c, stride),
               timeit([lambda: call("nasb_pool3x3_bwd", ref(desc(dy)), 0, stride, ptr(arg), ref(desc(dx)))]), nb)


# (n, c, h, w, ks, stride, dil, pad)
DW_SHAPES = [(8, 144, 256, 512, 3, 1, 1, 1), (8, 32, 512, 1024, 3, 1, 1, 1), (8, 24, 256, 512, 5, 1, 1, 2),
             (8, 32, 128, 256, 5, 1, 1, 2), (8, 32, 128, 256, 5, 1, 6, 12), (8, 192, 128, 256, 3, 1, 1, 1),
             (8, 96, 512, 1024, 3, 2, 1, 1), (8, 144, 256, 512, 3, 2, 1, 1), (8, 64, 128, 256, 5, 1, 1, 2),
             (8, 32, 256, 512, 3, 1, 1, 1), (8, 64, 32, 64, 5, 1, 1, 2),
             # dilated separable ops of the search space at the search loop's geometry
             (64, 48, 64, 64, 5, 1, 6, 12), (32, 48, 88, 88, 5, 1, 6, 12), (64, 48, 64, 64, 3, 1, 3, 3)]


def _ohw(h, w, ks, s, d, p):
    return (h + 2 * p - d * (ks - 1) - 1) // s + 1, (w + 2 * p - d * (ks - 1) - 1) // s + 1


@case("dwconv_tile")
def _():
    for n, c, h, w, ks, s, d, p in DW_SHAPES:
        oh, ow = _ohw(h, w, ks, s, d, p)
        nb = n * c * (h * w + oh * ow) * 2
        for stats in (0, 1):
            sets = []
            for _ in range(nsets(nb)):
                x, o = act(n, c, h, w), act(n, c, oh, ow, fill=None)
                wt = torch.randn(c, 1, ks, ks, device=DEV)
                st = torch.zeros(2 * c, dtype=torch.float64, device=DEV) if stats else None
                sets.append(lambda x=x, o=o, wt=wt, st=st: call("nasb_dwconv_tile", ref(desc(x)), ptr(wt), ks, s, d, p, 0, None, None, 0,
                                                                ref(desc(o)), ptr(st)))
            report("dwconv_tile[%dx%dx%dx%d k%d s%d d%d%s]" % (n, h, w, c, ks, s, d, " +stats" if stats else ""), timeit(sets), nb)


@case("dwconv_wgrad_tile")
def _():
    for n, c, h, w, ks, s, d, p in DW_SHAPES:
        oh, ow = _ohw(h, w, ks, s, d, p)
        nb = n * c * (h * w + oh * ow) * 2
        sets = []
        for _ in range(nsets(nb)):
            x, dz = act(n, c, h, w), act(n, c, oh, ow)
            dwt = torch.zeros(c, 1, ks, ks, device=DEV)
            sets.append(lambda x=x, dz=dz, dwt=dwt: call("nasb_dwconv_wgrad_tile", ref(desc(x)), ref(desc(dz)), ks, s, d, p, ptr(dwt)))
        report("dwconv_wgrad_tile[%dx%dx%dx%d k%d s%d d%d]" % (n, h, w, c, ks, s, d), timeit(sets), nb)


@case("dwconv_dgrad_strided_tile")
def _():
    for n, c, h, w, ks, s, d, p in [t for t in DW_SHAPES if t[5] == 2]:
        oh, ow = _ohw(h, w, ks, s, d, p)
        nb = n * c * (h * w + oh * ow) * 2
        sets = []
        for _ in range(nsets(nb)):
            dz, dx = act(n, c, oh, ow), act(n, c, h, w, fill=None)
            wt = torch.randn(c, 1, ks, ks, device=DEV)
            sets.append(lambda dz=dz, dx=dx, wt=wt: call("nasb_dwconv_dgrad_strided_tile", ref(desc(dz)), ptr(wt), ks, s, d, p, ref(desc(dx))))
        report("dwconv_dgrad_strided_tile[%dx%dx%dx%d k%d]" % (n, h, w, c, ks), timeit(sets), nb)


@case("sepconv")
def _():
    """The fused separable inference kernel (csrc/sep_tcgen05.cu) next to the two kernels it replaces (TMA depthwise tile
    kernel + tcgen05 pointwise kernel), eval-mode constants, batch 8 and batch 1."""
    for n, h, w, c, cout, ks in [(8, 256, 512, 144, 24, 3), (8, 128, 256, 192, 32, 3), (8, 512, 1024, 32, 16, 3), (8, 128, 256, 64, 64, 5),
                                 (8, 256, 512, 32, 32, 3), (1, 256, 512, 144, 24, 3), (1, 128, 256, 64, 64, 5)]:
        pad = (ks - 1) // 2
        nb = n * h * w * (c + cout) * 2
        fused, two = [], []
        for _ in range(nsets(nb + n * h * w * c * 4)):
            x, t, y = act(n, c, h, w), act(n, c, h, w, fill=None), act(n, cout, h, w, fill=None)
            dww = torch.randn(c, 1, ks, ks, device=DEV)
            ms, mb, os_, ob = fvec(c), fvec(c, -0.5, 0.5), fvec(cout), fvec(cout, -0.5, 0.5)
            wp = torch.randn(cout * ((c + 7) // 8 * 8), device=DEV).to(torch.bfloat16)
            fused.append(lambda x=x, y=y, dww=dww, ms=ms, mb=mb, os_=os_, ob=ob, wp=wp: call(
                "nasb_sepconv_tc_fwd", ref(desc(x)), ptr(dww), ks, 1, 1, pad, ptr(ms), ptr(mb), 2, ptr(wp), cout, ptr(os_), ptr(ob), 0,
                None, ref(desc(y))))

            def both(x=x, t=t, y=y, dww=dww, ms=ms, mb=mb, os_=os_, ob=ob, wp=wp):
                call("nasb_dwconv_tile", ref(desc(x)), ptr(dww), ks, 1, 1, pad, 0, ptr(ms), ptr(mb), 2, ref(desc(t)), None)
                call("nasb_pw_tc_fwd", ref(desc(t)), ptr(wp), cout, ptr(os_), ptr(ob), 0, None, ref(desc(y)), None)
            two.append(both)
        report("sepconv[%dx%dx%d %d->%d k%d fused]" % (n, h, w, c, cout, ks), timeit(fused), nb)
        report("sepconv[%dx%dx%d %d->%d k%d dw+pw]" % (n, h, w, c, cout, ks), timeit(two), nb)


@case("dgrad_gated")
def _():
    """Data gradients with the NasbGate epilogue (gate + the producer's BN-backward reductions) next to the plain ones."""
    import ctypes as C
    # depthwise 3x3: (n, c, h, w of dx, stride)
    for n, c, h, w, s in [(8, 96, 512, 1024, 2), (8, 144, 256, 512, 2), (8, 144, 256, 512, 1), (8, 32, 512, 1024, 1), (8, 192, 128, 256, 1)]:
        oh, ow = _ohw(h, w, 3, s, 1, 1)
        nb = n * c * (h * w + oh * ow) * 2
        for gated in (0, 1):
            sets = []
            for _ in range(nsets(nb + gated * n * c * h * w * 2)):
                dz, dx, z = act(n, c, oh, ow), act(n, c, h, w, fill=None), act(n, c, h, w)
                wt = torch.randn(c, 1, 3, 3, device=DEV)
                gs, gb, sums = fvec(c), fvec(c, -0.5, 0.5), torch.zeros(2 * c, dtype=torch.float64, device=DEV)
                gate = lib.NasbGate(C.pointer(desc(z)), ptr(gs), ptr(gb), 2, 0, ptr(sums))
                if gated:
                    sets.append(lambda dz=dz, dx=dx, wt=wt, gate=gate, keep=(z, gs, gb, sums): call(
                        "nasb_dwconv_dgrad_gated", ref(desc(dz)), ptr(wt), 3, s, 1, 1, C.byref(gate), ref(desc(dx))))
                elif s == 1:
                    sets.append(lambda dz=dz, dx=dx, wt=wt: call("nasb_dwconv_tile", ref(desc(dz)), ptr(wt), 3, 1, 1, 1, 1, None, None,
                                                                 0, ref(desc(dx)), None))
                else:
                    sets.append(lambda dz=dz, dx=dx, wt=wt: call("nasb_dwconv_dgrad_strided_tile", ref(desc(dz)), ptr(wt), 3, s, 1, 1,
                                                                 ref(desc(dx))))
            report("dw3x3_dgrad[%dx%dx%dx%d s%d %s]" % (n, h, w, c, s, "gated" if gated else "plain"), timeit(sets), nb)
    # pointwise: dz [P, cdz] -> dx [P, cdx]
    for n, h, w, cdz, cdx in [(8, 256, 512, 24, 144), (8, 512, 1024, 16, 32), (8, 256, 512, 64, 128), (8, 128, 256, 32, 192)]:
        nb = n * h * w * (cdz + cdx) * 2
        for gated in (0, 1):
            sets = []
            for _ in range(nsets(nb + gated * n * h * w * cdx * 2)):
                dz, dx, z = act(n, cdz, h, w), act(n, cdx, h, w, fill=None), act(n, cdx, h, w)
                wp = torch.randn(cdx * ((cdz + 7) // 8 * 8), device=DEV).to(torch.bfloat16)
                gs, gb, sums = fvec(cdx), fvec(cdx, -0.5, 0.5), torch.zeros(2 * cdx, dtype=torch.float64, device=DEV)
                gate = lib.NasbGate(C.pointer(desc(z)), ptr(gs), ptr(gb), 2, 0, ptr(sums))
                if gated:
                    sets.append(lambda dz=dz, dx=dx, wp=wp, gate=gate, keep=(z, gs, gb, sums): call(
                        "nasb_pw_tc_dgrad_gated", ref(desc(dz)), ptr(wp), cdx, C.byref(gate), ref(desc(dx))))
                else:
                    sets.append(lambda dz=dz, dx=dx, wp=wp: call("nasb_pw_tc_fwd", ref(desc(dz)), ptr(wp), cdx, None, None, 0, None,
                                                                 ref(desc(dx)), None))
            report("pw_dgrad[%dx%dx%d %d->%d %s]" % (n, h, w, cdz, cdx, "gated" if gated else "plain"), timeit(sets), nb)


# (n, h, w, cin, cout, stats)
PW_SHAPES = [(8, 512, 1024, 16, 96, 1), (8, 256, 512, 24, 144, 1), (8, 256, 512, 144, 24, 1), (8, 512, 1024, 32, 32, 1),
             (8, 512, 1024, 96, 16, 0), (8, 256, 512, 224, 64, 1), (8, 128, 256, 32, 192, 1), (8, 128, 256, 192, 32, 1),
             (8, 128, 256, 32, 32, 1), (8, 128, 256, 32, 64, 1), (8, 32, 64, 64, 64, 1), (8, 256, 512, 24, 24, 1),
             # MobileNet-v2's last stage at the search loop's task-1 geometry (batch 32 @350x350 -> 11x11): K-ring mode
             (32, 11, 11, 960, 160, 1), (32, 11, 11, 160, 960, 1), (32, 11, 11, 960, 320, 1)]


@case("pw_tc_fwd")
def _():
    for n, h, w, ci, co, stats in PW_SHAPES:
        nb = n * h * w * (ci + co) * 2
        sets = []
        for _ in range(nsets(nb)):
            x, o = act(n, ci, h, w), act(n, co, h, w, fill=None)
            wt = torch.randn(co, ci, 1, 1, device=DEV) * 0.1
            wp = torch.empty(co * ((ci + 7) // 8 * 8), dtype=BF, device=DEV)
            call("nasb_pack_weight_bf16", ptr(wt), co, ci, 0, ptr(wp))
            st = torch.zeros(2 * co, dtype=torch.float64, device=DEV) if stats else None
            sets.append(lambda x=x, o=o, wp=wp, st=st: call("nasb_pw_tc_fwd", ref(desc(x)), ptr(wp), co, None, None, 0, None, ref(desc(o)), ptr(st)))
        report("pw_tc_fwd[%dx%dx%d %d->%d%s]" % (n, h, w, ci, co, " +stats" if stats else ""), timeit(sets), nb)


@case("pw_tc_wgrad")
def _():
    for n, h, w, ci, co, _s in PW_SHAPES:
        nb = n * h * w * (ci + co) * 2
        sets = []
        for _ in range(nsets(nb)):
            x, dz = act(n, ci, h, w), act(n, co, h, w)
            dwt = torch.zeros(co, ci, device=DEV)
            sets.append(lambda x=x, dz=dz, dwt=dwt: call("nasb_pw_tc_wgrad", ref(desc(x)), ref(desc(dz)), ptr(dwt)))
        report("pw_tc_wgrad[%dx%dx%d %d->%d]" % (n, h, w, ci, co), timeit(sets), nb)


@case("stem")
def _():
    n, h, w = 8, 1024, 2048
    img = torch.randn(n, 3, h, w, device=DEV)
    wt = torch.randn(32, 3, 3, 3, device=DEV)
    o = act(n, 32, h // 2, w // 2, fill=None)
    dz = act(n, 32, h // 2, w // 2)
    dwt = torch.zeros(32, 3, 3, 3, device=DEV)
    nb = n * 3 * h * w * 4 + n * 32 * (h // 2) * (w // 2) * 2
    d = lib.desc_nchw_f32(img)
    report("stem_fwd", timeit([lambda: call("nasb_stem_fwd", ref(d), ptr(wt), 3, 2, 1, 1, None, None, 0, ref(desc(o)))]), nb)
    report("stem_wgrad", timeit([lambda: call("nasb_stem_wgrad", ref(d), ref(desc(dz)), 3, 2, 1, 1, ptr(dwt))]), nb)


@case("conv3_tc")
def _():
    for n, h, w, ci, co, odt in [(8, 256, 512, 64, 19, torch.float32), (8, 256, 512, 24, 64, BF), (32, 88, 88, 48, 48, BF), (64, 64, 64, 48, 21, torch.float32)]:
        cip = (ci + 7) // 8 * 8
        xb = torch.randn(n, h, w, cip, device=DEV).to(BF)
        x = xb[..., :ci].permute(0, 3, 1, 2)
        o = act(n, co, h, w, odt, fill=None)
        wt = torch.randn(co, ci, 3, 3, device=DEV) * 0.1
        ne = int(lib.load().nasb_pack_conv3_elems(co, ci, 0))
        wp = torch.empty(ne, dtype=BF, device=DEV)
        call("nasb_pack_conv3_bf16", ptr(wt), co, ci, 0, ptr(wp))
        nb = n * h * w * (ci * 2 + co * (4 if odt == torch.float32 else 2))
        report("conv3_tc_fwd[%dx%dx%d %d->%d]" % (n, h, w, ci, co),
               timeit([lambda: call("nasb_conv3_tc_fwd", ref(desc(x)), ptr(wp), co, 1, 1, None, None, 0, ref(desc(o)), None)]), nb)


@case("resize_bwd")
def _():
    for n, c, h, w, ih, iw in [(8, 64, 256, 512, 32, 64), (8, 64, 256, 512, 256, 512), (8, 64, 128, 256, 32, 64), (8, 32, 256, 512, 128, 256)]:
        dz, dx = act(n, c, h, w), act(n, c, ih, iw, fill=None)
        nb = n * c * (h * w + ih * iw) * 2
        report("resize_bwd[%dx%dx%dx%d -> %dx%d]" % (n, h, w, c, ih, iw), timeit([lambda: call("nasb_resize_bwd", ref(desc(dz)), None, ref(desc(dx)))]), nb)


def main():
    filt = sys.argv[1:]
    lib.load()
    print("# kbench on %s, env: %s" % (torch.cuda.get_device_name(0), {k: v for k, v in os.environ.items() if k.startswith("NASB_")}))
    for name, f in CASES:
        if filt and not any(s in name for s in filt):
            continue
        f()
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
