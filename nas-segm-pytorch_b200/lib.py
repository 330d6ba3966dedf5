"""ctypes binding of libnasb200.so (C ABI declared in include/nasb200.h).

Fails loudly: a missing library is an ImportError at first use, any non-zero return code is a RuntimeError
(the reference's engine wrappers turn RuntimeError into "reward 0", src/helpers/utils.py:172-187)."""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnasb200.so")

F32, BF16, F32_NCHW = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_RELU6 = 0, 1, 2
POOL_MAX, POOL_AVG = 0, 1


class NasbTensor(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32),
                ("cstride", C.c_int32), ("dtype", C.c_int32)]


class NasbGate(C.Structure):
    _fields_ = [("z", C.POINTER(NasbTensor)), ("scale", C.c_void_p), ("shift", C.c_void_p), ("act", C.c_int32),
                ("reserved", C.c_int32), ("sums", C.c_void_p)]


class NasbConvUnit(C.Structure):
    _fields_ = [("weight", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p), ("running_mean", C.c_void_p),
                ("running_var", C.c_void_p), ("bias", C.c_void_p), ("eps", C.c_float), ("c_out", C.c_int32), ("ks", C.c_int32),
                ("stride", C.c_int32), ("dil", C.c_int32), ("pad", C.c_int32), ("dw", C.c_int32), ("in_relu", C.c_int32),
                ("act", C.c_int32)]


class NasbAugSample(C.Structure):
    _fields_ = [("image", C.c_void_p), ("mask", C.c_void_p), ("h", C.c_int32), ("w", C.c_int32), ("rh", C.c_int32),
                ("rw", C.c_int32), ("scale", C.c_double), ("top", C.c_int32), ("left", C.c_int32), ("mirror", C.c_int32),
                ("reserved", C.c_int32)]


_TP = C.POINTER(NasbTensor)
_P, _I, _F, _L = C.c_void_p, C.c_int, C.c_float, C.c_longlong

# name -> argtypes (return type int unless listed in _RET)
_SIG = {
    "nasb_conv_fwd": [_TP, _TP, _P, _I, _I, _I, _I, _P, _P, _I, _P, _P, _I, _TP, _TP, _P],
    "nasb_conv_dgrad": [_TP, _P, _I, _I, _I, _I, _TP, _TP, _P],
    "nasb_conv_wgrad": [_TP, _TP, _P, _P, _I, _TP, _I, _I, _I, _I, _P, _P],
    "nasb_pack_conv3_elems": [_I, _I, _I],
    "nasb_pack_conv3_bf16": [_P, _I, _I, _I, _P, _P],
    "nasb_conv3_tc_supported": [_I, _I],
    "nasb_conv3_tc_fwd": [_TP, _P, _I, _I, _I, _P, _P, _I, _TP, _P, _P],
    "nasb_conv3_tc_wgrad": [_TP, _TP, _I, _I, _P, _P],
    "nasb_stem_fwd": [_TP, _P, _I, _I, _I, _I, _P, _P, _I, _TP, _P],
    "nasb_stem_wgrad": [_TP, _TP, _I, _I, _I, _I, _P, _P],
    "nasb_stem_im2col": [_TP, _I, _I, _I, _I, _TP, _P],
    "nasb_pack_weight_bf16": [_P, _I, _I, _I, _P, _P],
    "nasb_pw_tc_supported": [_I, _I],
    "nasb_pw_tc_wgrad_supported": [_I, _I],
    "nasb_pw_tc_wgrad": [_TP, _TP, _P, _P],
    "nasb_pw_tc_fwd": [_TP, _P, _I, _P, _P, _I, _TP, _TP, _P, _P],
    "nasb_bn_finalize": [_P, _L, _I, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P, _P, _P],
    "nasb_dwconv_fwd": [_TP, _P, _I, _I, _I, _I, _I, _P, _P, _I, _TP, _P],
    "nasb_dwconv_dgrad": [_TP, _P, _I, _I, _I, _I, _TP, _P],
    "nasb_dwconv_wgrad": [_TP, _I, _TP, _I, _I, _I, _I, _P, _P],
    "nasb_dwconv_tile": [_TP, _P, _I, _I, _I, _I, _I, _P, _P, _I, _TP, _P, _P],
    "nasb_dwconv_dgrad_strided_tile": [_TP, _P, _I, _I, _I, _I, _TP, _P],
    "nasb_dwconv_wgrad_tile": [_TP, _TP, _I, _I, _I, _I, _P, _P],
    "nasb_dwconv_dgrad_gated": [_TP, _P, _I, _I, _I, _I, _P, _TP, _P],
    "nasb_pw_tc_dgrad_gated": [_TP, _P, _I, _P, _TP, _P],
    "nasb_bn_bwd_from_sums": [_TP, _TP, _I, _P, _P, _P, _P, _P, _P, _P, _TP, _P, _P],
    "nasb_pw_bn_bwd_prepare": [_P, _I, _I, _P, _P, _P, _P, _L, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "nasb_pw_bn_bwd_scratch": [_I],
    "nasb_bn_fold": [_P, _P, _P, _P, _F, _I, _P, _P, _P],
    "nasb_bn_stats_workspace": [_I],
    "nasb_bn_stats": [_TP, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "nasb_affine_act": [_TP, _P, _P, _I, _TP, _P],
    "nasb_bn_finalize_affine_act": [_P, _L, _TP, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P, _P, _I, _TP, _TP, _P],
    "nasb_bn_act_bwd": [_TP, _TP, _TP, _I, _P, _P, _P, _P, _P, _P, _I, _P, _P, _TP, _P, _P],
    "nasb_pool3x3_fwd": [_TP, _I, _I, _TP, _P, _P],
    "nasb_pool3x3_bwd": [_TP, _I, _I, _P, _TP, _P],
    "nasb_resize_axpby": [_TP, _P, _TP, _P, _I, _TP, _P],
    "nasb_resize_bwd": [_TP, _P, _TP, _P],
    "nasb_axpby_bwd_params": [_TP, _TP, _TP, _P, _P, _P, _P],
    "nasb_scale_copy": [_TP, _P, _I, _TP, _P],
    "nasb_channel_tile": [_TP, _I, _F, _TP, _P],
    "nasb_channel_tile_bwd": [_TP, _I, _F, _TP, _P],
    "nasb_spatial_mean": [_TP, _P, _P],
    "nasb_spatial_bcast": [_TP, _F, _TP, _P],
    "nasb_spatial_sum": [_TP, _P, _P],
    "nasb_channel_sum": [_TP, _P, _P, _P],
    "nasb_relu_bwd": [_TP, _TP, _TP, _P],
    "nasb_loss_workspace": [],
    "nasb_ce_fwd": [_TP, _P, _I, _P, _P, _P],
    "nasb_ce_bwd": [_TP, _P, _I, _P, _P, _TP, _P],
    "nasb_mse_fwd": [_TP, _TP, _P, _P, _P],
    "nasb_mse_bwd": [_TP, _TP, _P, _TP, _P],
    "nasb_berhu_fwd": [_TP, _TP, _F, _P, _P, _P],
    "nasb_berhu_bwd": [_TP, _TP, _F, _P, _P, _TP, _P],
    "nasb_confmat_labels": [_P, _P, _L, _I, _P, _P],
    "nasb_confmat_logits": [_TP, _P, _I, _I, _I, _P, _P],
    "nasb_ius_accs": [_P, _I, _P, _P, _P, _P],
    "nasb_sumsq": [_P, _L, _P, _P],
    "nasb_mt_grad_sumsq": [_P, _I, _P, _I, _P],
    "nasb_mt_optim_step": [_P, _I, _P, _I, _P, _I, _P, _F, _P],
    "nasb_pack_elems": [_I, _I, _I],
    "nasb_mt_pack_bf16": [_P, _I, _P],
    "nasb_conv_unit_scratch": [_I, _I],
    "nasb_conv_unit_infer": [_TP, _P, _TP, _TP, _P, _L, _I, _P],
    "nasb_conv_units_prepare": [_P, _P, _P, _I, _I, _P],
    "nasb_sepconv_tc_supported": [_I, _I, _I, _I, _I, _I],
    "nasb_sepconv_tc_fwd": [_TP, _P, _I, _I, _I, _I, _P, _P, _I, _P, _I, _P, _P, _I, _TP, _TP, _P],
    "nasb_sep_unit_infer": [_TP, _P, _P, _L, _P, _P, _L, _TP, _TP, _I, _P],
    "nasb_augment_batch": [C.POINTER(NasbAugSample), _I, _I, _I, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double), _P, _P, _P],
    "nasb_version": [],
}
_RET = {"nasb_version": C.c_char_p, "nasb_bn_stats_workspace": _L, "nasb_loss_workspace": _L, "nasb_pack_conv3_elems": _L,
        "nasb_pack_elems": _L, "nasb_conv_unit_scratch": _L, "nasb_pw_bn_bwd_scratch": _L}
EXPORTS = tuple(sorted(_SIG))

_lib = None
launches = 0  # kernels launched through the C ABI by this process (bench.py reports the delta over its timed region)
# kernels (and async memsets) behind one call of each entry point; everything not listed launches exactly one
_KERNELS_PER_CALL = {"nasb_sep_unit_infer": 4, "nasb_bn_bwd_from_sums": 2, "nasb_pw_bn_bwd_prepare": 2, "nasb_conv_unit_infer": 3, "nasb_mt_grad_sumsq": 3, "nasb_mt_optim_step": 2, "nasb_bn_stats": 3, "nasb_bn_act_bwd": 3, "nasb_ce_fwd": 3, "nasb_mse_fwd": 3, "nasb_berhu_fwd": 4,
                     "nasb_spatial_mean": 2, "nasb_spatial_sum": 2}
_prof = None  # list of (key, bytes, ev0, ev1) while profiling


def load():
    """dlopen the library (once) and attach signatures.  No compute, no GPU needed."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libnasb200.so is not built: run `python nas-segm-pytorch_b200/build.py` "
                              "(there is no CPU / eager fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, args in _SIG.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = _RET.get(name, C.c_int)
        _lib = lib
    return _lib


def version():
    return load().nasb_version().decode()


def require_cuda(t):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError("nas-segm-pytorch_b200 kernels run on a CUDA device only (no CPU fallback); got %s"
                           % (t.device if isinstance(t, torch.Tensor) else type(t)))


def call(name, *args):
    """Invoke an entry point on torch's current stream; raise RuntimeError on failure."""
    global launches
    fn = getattr(_lib or load(), name)
    if _prof is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        rc = fn(*args, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        ev1.record()
        key, nbytes = _describe(name, args)
        _prof.append((key, nbytes, ev0, ev1))
    else:
        rc = fn(*args, torch.cuda.current_stream().cuda_stream)
    launches += _KERNELS_PER_CALL.get(name, 1)
    if rc != 0:
        raise RuntimeError("%s failed with code %d%s" % (name, rc, _explain(rc)))


def try_call(name, *args):
    """Like call(), but NASB_ERR_UNSUPPORTED is returned (False) instead of raised: used where a specialised kernel
    declines a shape and the general CUDA kernel takes over (still no CPU fallback)."""
    global launches
    fn = getattr(load(), name)
    if _prof is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        rc = fn(*args, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        ev1.record()
        if rc == 0:
            key, nbytes = _describe(name, args)
            _prof.append((key, nbytes, ev0, ev1))
    else:
        rc = fn(*args, torch.cuda.current_stream().cuda_stream)
    if rc == 10001:
        return False
    if rc != 0:
        raise RuntimeError("%s failed with code %d%s" % (name, rc, _explain(rc)))
    launches += _KERNELS_PER_CALL.get(name, 1)
    return True


def _describe(name, args):
    """(key, algorithmic bytes) of one call: every NasbTensor argument counted once (reads + writes), the unit-level
    definition of SURVEY 8(d): numel(inputs) + numel(outputs), parameters being negligible."""
    shapes, nbytes = [], 0
    for a in args:
        t = getattr(a, "_obj", None)
        if isinstance(t, NasbTensor):
            esz = 2 if t.dtype == BF16 else 4
            nbytes += t.n * t.h * t.w * t.c * esz
            shapes.append("%dx%dx%dx%d%s" % (t.n, t.h, t.w, t.c, "b" if t.dtype == BF16 else "f"))
    ints = [str(a) for a in args if isinstance(a, int) and not isinstance(a, bool) and -65536 < a < 65536][:4]  # not addresses
    return name + "[" + ",".join(shapes) + "|" + ",".join(ints) + "]", nbytes


def profile_begin():
    global _prof
    _prof = []


def profile_end():
    """-> {key: (calls, total_ms, algorithmic_bytes_per_call)} with CUDA-event durations on the launching stream."""
    global _prof
    rec, _prof = _prof, None
    torch.cuda.synchronize()
    out = {}
    for key, nbytes, e0, e1 in rec:
        c, ms, b = out.get(key, (0, 0.0, nbytes))
        out[key] = (c + 1, ms + e0.elapsed_time(e1), b)
    return out


def _explain(rc):
    if rc == 10001:
        return " (unsupported configuration)"
    if rc == 10002:
        return " (bad argument)"
    try:
        return " (cudaError %d)" % rc
    except Exception:
        return ""


_DT = {torch.float32: F32, torch.bfloat16: BF16}


def new_act(n, c, h, w, dtype, device):
    """Allocate an activation: logical NCHW view over dense NHWC memory."""
    t = torch.empty((n, h, w, c), dtype=dtype, device=device).permute(0, 3, 1, 2)
    t._nasb_d = NasbTensor(t.data_ptr(), n, h, w, c, c, _DT[dtype])
    return t


def zeros_act(n, c, h, w, dtype, device):
    return torch.zeros((n, h, w, c), dtype=dtype, device=device).permute(0, 3, 1, 2)


def is_nhwc(t):
    """True if logical-NCHW tensor t is addressable as NHWC with a constant pixel stride (channel slices allowed)."""
    if t.dim() != 4:
        return False
    n, c, h, w = t.shape
    sn, sc, sh, sw = t.stride()
    if c > 1 and sc != 1:
        return False
    cs = sw if w > 1 else (sh // max(w, 1) if h > 1 else (sn // max(h * w, 1) if n > 1 else c))
    if cs < c:
        return False
    if w > 1 and sw != cs:
        return False
    if h > 1 and sh != w * cs:
        return False
    if n > 1 and sn != h * w * cs:
        return False
    return True


def to_nhwc(t, dtype=None):
    """Boundary conversion of a user tensor to dense NHWC storage (torch copy; not on the hot path)."""
    require_cuda(t)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    if is_nhwc(t) and t.dtype in _DT:
        return t
    return t.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)


def desc(t):
    """NasbTensor for a logical-NCHW torch tensor with NHWC storage (memoised on the tensor object: a tensor's geometry
    and address never change, and an activation is described once as an output and again by each of its consumers)."""
    d = getattr(t, "_nasb_d", None)
    if d is not None:
        return d
    n, c, h, w = t.shape
    sn, sc, sh, sw = t.stride()
    if w > 1:
        cs = sw
    elif h > 1:
        cs = sh
    elif n > 1:
        cs = sn
    else:
        cs = c
    d = NasbTensor(t.data_ptr(), n, h, w, c, cs, _DT[t.dtype])
    t._nasb_d = d
    return d


def desc_nchw_f32(t):
    """Planar fp32 image (contiguous NCHW) for the encoder stem."""
    assert t.dtype == torch.float32 and t.is_contiguous()
    n, c, h, w = t.shape
    return NasbTensor(t.data_ptr(), n, h, w, c, c, F32_NCHW)


def ref(d):
    return C.byref(d) if d is not None else None


def ptr(t):
    """Device address for a void* argument (ctypes converts a plain int; None is NULL)."""
    return t.data_ptr() if t is not None else None


class _ZeroArena:
    """Zero-initialised scratch for one training iteration (gradient accumulators of the atomically-reduced kernels,
    fp64 statistics cells): one memset per iteration instead of ~300 two-microsecond fill kernels.  The engine's step
    functions bracket an iteration with begin()/end(); outside of that take() is plain torch.zeros.  A tensor taken in
    iteration i is zeroed again by begin() of iteration i+1, so only the engine loops switch it on: they drop every
    gradient with zero_grad(set_to_none=True) before the next backward pass, which is what makes handing arena views to
    autograd as parameter gradients safe.  Code that keeps .grad tensors alive across iterations (gradient accumulation,
    set_to_none=False) must not run inside begin() / end() -- outside of it take() is plain torch.zeros.

    The buffer of a device is allocated ONCE (CAPACITY bytes) and never replaced: a captured CUDA graph bakes in its
    address (the memset and every view), and a graph of an earlier model may still be replayed after a larger model has
    come along.  begin() clears only the prefix that earlier iterations were seen to need (the high-water mark, which
    only grows); a take() that does not fit into the cleared prefix -- the first iteration of a model, or demand beyond
    CAPACITY -- is served by torch.zeros."""

    CAPACITY = 64 << 20

    def __init__(self):
        self.bufs, self.high = {}, {}
        self.buf, self.off, self.cleared, self.active = None, 0, 0, False

    @staticmethod
    def _capturing(device):
        return device.type == "cuda" and torch.cuda.is_current_stream_capturing()

    def begin(self, device):
        device = torch.device(device)
        buf = self.bufs.get(device)
        if buf is None:
            if self._capturing(device):  # never allocate the long-lived buffer from a graph's private pool
                self.active = False
                return
            buf = self.bufs[device] = torch.empty(self.CAPACITY, dtype=torch.uint8, device=device)
        n = min((self.high.get(device, 0) * 5 // 4 + 4095) // 4096 * 4096, self.CAPACITY)
        if n:
            buf[:n].zero_()
        self.buf, self.off, self.cleared, self.active = buf, 0, n, True

    def end(self):
        self.active = False

    def take(self, shape, dtype, device):
        if isinstance(shape, int):
            shape = (shape,)
        n = 1
        for d in shape:
            n *= int(d)
        if self.active and self.buf is not None and self.buf.device == torch.device(device):
            nbytes = (n * torch.empty((), dtype=dtype).element_size() + 15) // 16 * 16
            start, self.off = self.off, self.off + nbytes
            if self.off > self.high.get(self.buf.device, 0):
                self.high[self.buf.device] = self.off
            if self.off <= self.cleared:
                return self.buf[start:start + nbytes].view(dtype)[:n].view(shape)
        return torch.zeros(shape, dtype=dtype, device=device)


zero_arena = _ZeroArena()


def zeros(shape, dtype, device):
    """Zero-filled tensor: from the per-iteration arena inside an engine step, torch.zeros otherwise."""
    return zero_arena.take(shape, dtype, device)


class _WgradStream:
    """Weight-gradient kernels of an engine iteration on a second stream.

    Nothing in the backward pass depends on a weight gradient until the optimiser step, and the ~100 weight-gradient kernels
    of the small decoder layers are latency-bound (0.2-1 TB/s): inside an engine iteration (begin() ... end()) each unit forks
    them onto one side stream right after its dz exists and the main stream carries on with the data gradient; join() -- called
    before the optimiser step and by end() -- makes the main stream wait.  Works the same under CUDA-graph capture (the fork
    and join become graph edges).  Tensors the side stream reads are recorded on it, so the allocator cannot hand their
    memory to a later main-stream allocation while the side kernel is still pending.  Outside an engine iteration the weight
    gradients stay on the caller's stream (a user reading .grad right after backward() must not need to know about this)."""

    def __init__(self):
        self.streams, self.active, self.forked = {}, False, None

    def begin(self, device):
        from . import config
        device = torch.device(device)
        self.active = bool(config().async_wgrad) and device.type == "cuda"
        self.forked = None

    def fork(self, tensors):
        """Context manager that routes launches to the side stream, or None when inactive."""
        if not self.active:
            return None
        main = torch.cuda.current_stream()
        side = self.streams.get(main.device)
        if side is None:
            side = self.streams[main.device] = torch.cuda.Stream(device=main.device)
        ev = torch.cuda.Event()
        ev.record(main)
        side.wait_event(ev)
        for t in tensors:
            if t is not None:
                t.record_stream(side)
        self.forked = side
        return torch.cuda.stream(side)

    def join(self):
        if self.forked is not None:
            torch.cuda.current_stream().wait_stream(self.forked)
            self.forked = None

    def end(self):
        self.join()
        self.active = False


wgrad_stream = _WgradStream()


class _BranchStreams:
    """Independent branches of the decoder graph on concurrent streams.

    Both inputs of every aggregation (TemplateDecoder: op1(feat1) / op2(feat2); MergeCell: two contextual cells; a cell
    layer's two ops; Adapt's two 1x1 convolutions) are independent sub-graphs of small, latency-bound kernels.  run2(fa, fb)
    executes fb on a side stream (one per nesting depth) between a fork and a join of the current stream -- plain stream
    semantics in eager mode, graph edges under CUDA-graph capture, and autograd replays the backward of each branch on the
    stream its forward ran on.  Tensors that cross streams are recorded on the consuming stream."""

    def __init__(self):
        self.pool, self.depth, self.ws, self.quiet = {}, 0, {}, False

    def side(self, device):
        key = (device.index, self.depth)
        s = self.pool.get(key)
        if s is None:
            s = self.pool[key] = torch.cuda.Stream(device=device)
            # reduction scratch of its own: two branches may run BatchNorm reductions at the same time
            self.ws[s.cuda_stream] = torch.empty(1 << 20, dtype=torch.uint8, device=device)
        return s

    def run2(self, fa, fb, b_inputs=()):
        from . import config
        if not config().branch_streams or not torch.cuda.is_available():
            return fa(), fb()
        main = torch.cuda.current_stream()
        capturing = torch.cuda.is_current_stream_capturing()
        if capturing and (main.device.index, self.depth) not in self.pool:
            return fa(), fb()  # never create streams / scratch inside a capture (the eager warm-up creates them)
        side = self.side(main.device)
        if torch.is_grad_enabled() and not self.quiet:
            # parameters live on the stream they were created on, their gradients now arrive from the branch streams: autograd's
            # "AccumulateGrad stream mismatch" note is about a synchronisation this design accepts
            fn = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
            if fn is not None:
                fn(False)
            self.quiet = True
        if not capturing and not torch.is_grad_enabled():
            # eager inference is bound by host dispatch, not by the GPU: a fork / join (two event round trips) costs more than
            # the overlap returns (arch0 480x360: 4.1 -> 4.9 ms).  The streams exist now, so a later capture can use them.
            self.depth += 1
            try:
                return fa(), fb()
            finally:
                self.depth -= 1
        self.depth += 1
        try:
            for t in b_inputs:
                if t is not None and t.is_cuda:
                    t.record_stream(side)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                b = fb()
            a = fa()
            main.wait_stream(side)
        finally:
            self.depth -= 1
        for t in (b if isinstance(b, (tuple, list)) else (b,)):
            if torch.is_tensor(t) and t.is_cuda:
                t.record_stream(main)
        return a, b


branch_streams = _BranchStreams()

_ws = {}


def workspace(device, nbytes=1 << 20):
    """Per-device scratch for the reductions (BN statistics, loss sums); calls on one stream are serialised.  The side
    streams of _BranchStreams have a scratch each."""
    if branch_streams.ws:
        own = branch_streams.ws.get(torch.cuda.current_stream().cuda_stream)
        if own is not None and own.numel() >= nbytes:
            return own
    key = (device.index if device.index is not None else torch.cuda.current_device())
    buf = _ws.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _ws[key] = buf
    return buf
