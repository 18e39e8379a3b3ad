"""Host wrappers of the tcgen05 implicit-GEMM convolution engine (csrc/conv3d_tc.cu, csrc/conv3d_tc_wgrad.cu).

Activations of this engine are bf16, channels-last: a (B, C, D, H, W) logical tensor is held as a contiguous
(B, D, H, W, C) bf16 tensor (`ndhwc`).  Weights stay fp32 in the nn.Module (reference layout
(Cout, Cin, 3, 3, 3)); packed bf16 copies are refreshed whenever the parameter version changes.
"""
import ctypes
import threading

import torch

from . import _lib


def np_for(cout):
    """MMA N (16 or 32) for a layer with `cout` output channels."""
    if cout <= 16:
        return 16
    if cout <= 32:
        return 32
    if cout <= 48:
        return 48
    if cout <= 64:
        return 64
    raise _lib.VxmError("tcgen05 conv engine: at most 64 output channels per launch (got %d)" % cout)


def to_ndhwc_bf16(x):
    """(B,C,D,H,W) fp32 -> (B,D,H,W,C) bf16 contiguous (test / boundary helper; torch copy kernels)."""
    return x.permute(0, 2, 3, 4, 1).contiguous().to(torch.bfloat16)


def from_ndhwc(x):
    return x.permute(0, 4, 1, 2, 3).float().contiguous()


def pack_weights(w, transposed=False):
    """fp32 (Cout, Cin, kd, 3, 3) -> packed bf16 operand for vxm_conv3d_tc_fwd (dgrad operator if transposed)."""
    lib = _lib.load()
    if w.dim() == 4:
        w = w.unsqueeze(2)
    w = w.contiguous()
    Cout, Cin, kd = w.shape[0], w.shape[1], w.shape[2]
    cin_eff, nout = (Cout, Cin) if transposed else (Cin, Cout)
    NP = np_for(nout)
    nbytes = int(lib.vxm_conv3d_tc_packed_bytes(cin_eff, NP, kd))
    out = torch.empty(nbytes // 2, dtype=torch.bfloat16, device=w.device)
    _lib.check(lib.vxm_conv3d_tc_pack(_lib.ptr(w), _lib.ptr(out), Cout, Cin, kd, NP, 1 if transposed else 0,
                                      _lib.stream_ptr()), "vxm_conv3d_tc_pack")
    return out, NP


def conv_fwd(xa, xb, wpk, NP, bias, cout, kd, up=False, planar=None, out_fp32_planar=False, slope=None, mask=None, split=None):
    """Launch the tensor-core convolution.  xa / xb: bf16 NDHWC tensors (xa at half resolution when `up`);
    planar: list of <= 4 fp32 (B,1,D,H,W) tensors used instead of xa/xb.  Returns bf16 NDHWC (B,D,H,W,cout)
    or, with out_fp32_planar, fp32 (B,cout,D,H,W)."""
    lib = _lib.load()
    if planar is not None:
        ref = planar[0]
        B, D, H, W = ref.shape[0], ref.shape[-3] if ref.dim() == 5 else 1, ref.shape[-2], ref.shape[-1]
        arr_p = (ctypes.c_void_p * 4)(*([p.data_ptr() for p in planar] + [0] * (4 - len(planar))))
        arr_s = (ctypes.c_longlong * 4)(*([p.stride(0) for p in planar] + [0] * (4 - len(planar))))
        xf, xs, npl, Ca, Cb = arr_p, arr_s, len(planar), 0, 0
        dev = ref.device
    else:
        full = xb if xb is not None else xa
        B, D, H, W = full.shape[0], full.shape[1], full.shape[2], full.shape[3]
        if xb is None and up:
            D, H, W = (D * 2 if kd == 3 else D), H * 2, W * 2
        Ca = 0 if xa is None else xa.shape[-1]
        Cb = 0 if xb is None else xb.shape[-1]
        xf, xs, npl = None, None, 0
        dev = full.device
    out2 = None
    if out_fp32_planar:
        out = torch.empty((B, cout, D, H, W), dtype=torch.float32, device=dev)
    elif split:
        out = torch.empty((B, D, H, W, split), dtype=torch.bfloat16, device=dev)
        out2 = torch.empty((B, D, H, W, cout - split), dtype=torch.bfloat16, device=dev)
    else:
        out = torch.empty((B, D, H, W, cout), dtype=torch.bfloat16, device=dev)
    s = -1.0 if slope is None else float(slope)
    _lib.check(lib.vxm_conv3d_tc_fwd(_lib.ptr(xa), _lib.ptr(xb), xf, xs, npl, _lib.ptr(wpk), _lib.ptr(bias), _lib.ptr(out),
                                     _lib.ptr(mask), B, D, H, W, Ca, Cb, 1 if up else 0, cout, NP, kd,
                                     1 if out_fp32_planar else 0, s, _lib.ptr(out2), int(split or 0), _lib.stream_ptr()),
               "vxm_conv3d_tc_fwd")
    return (out, out2) if split else out


def _planar_args(planar):
    if planar is None:
        return None, None, 0
    arr_p = (ctypes.c_void_p * 4)(*([p.data_ptr() for p in planar] + [0] * (4 - len(planar))))
    arr_s = (ctypes.c_longlong * 4)(*([p.stride(0) for p in planar] + [0] * (4 - len(planar))))
    return arr_p, arr_s, len(planar)


_wgrad_ws = {}
_ws_lock = threading.Lock()      # workspace tables are shared by the per-GPU threads of nn.DataParallel


def conv_wgrad(xa, xb, gz, cin, cout, kd, up=False, planar_x=None, planar_g=None, need_bias=True, out_w=None, out_b=None, batch=None):
    """fp32 grad_w (cout, cin, kd, 3, 3) and grad_b (cout) from the layer input (xa/xb or planar_x) and gz.
    `batch` (a WgradBatch): only the tcgen05 partial-sum kernels are launched now, the reduction into gw / gb happens at
    `batch.flush()` together with every other layer's."""
    lib = _lib.load()
    if batch is not None and wgrad_deferrable(xa, xb, gz, planar_x, planar_g):
        dev = gz.device
        accumulate = out_w is not None
        gw = out_w if accumulate else torch.empty((cout, cin, kd, 3, 3), dtype=torch.float32, device=dev)
        gb = out_b if accumulate else (torch.empty(cout, dtype=torch.float32, device=dev) if need_bias else None)
        batch.add(xa, xb, gz, gw, gb, cin, cout, kd, up, accumulate)
        return gw, gb
    ref = gz if gz is not None else planar_g[0]
    dev = ref.device
    if gz is not None:
        B, D, H, W, Cg = gz.shape
    else:
        B, D, H, W, Cg = ref.shape[0], (ref.shape[-3] if ref.dim() == 5 else 1), ref.shape[-2], ref.shape[-1], 8
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream, kd)
    work = _wgrad_ws.get(key)
    if work is None:
        with _ws_lock:
            work = _wgrad_ws.get(key)
            if work is None:
                work = torch.empty(int(lib.vxm_conv3d_tc_wgrad_workspace_bytes(kd)), dtype=torch.uint8, device=dev)
                _wgrad_ws[key] = work
    # out_w / out_b: existing (contiguous fp32) gradient buffers to ACCUMULATE into instead of fresh tensors
    accumulate = out_w is not None
    gw = out_w if accumulate else torch.empty((cout, cin, kd, 3, 3), dtype=torch.float32, device=dev)
    gb = out_b if accumulate else (torch.empty(cout, dtype=torch.float32, device=dev) if need_bias else None)
    xf, xs, npx = _planar_args(planar_x)
    gf, gs, npg = _planar_args(planar_g)
    Ca = 0 if xa is None else xa.shape[-1]
    Cb = 0 if xb is None else xb.shape[-1]
    _lib.check(lib.vxm_conv3d_tc_wgrad(_lib.ptr(xa), _lib.ptr(xb), xf, xs, npx, _lib.ptr(gz), gf, gs, npg, _lib.ptr(gw),
                                       _lib.ptr(gb), _lib.ptr(work), B, D, H, W, Ca, Cb, 1 if up else 0, cin, Cg, cout, kd,
                                       1 if accumulate else 0, _lib.stream_ptr()), "vxm_conv3d_tc_wgrad")
    return gw, gb


def planar_to_ndhwc8(planes):
    """<= 8 planar fp32 (B,1,[D,]H,W) volumes -> one bf16 (B,D,H,W,8) tensor (unused channels zero)."""
    lib = _lib.load()
    ref = planes[0]
    B = ref.shape[0]
    D, H, W = (ref.shape[-3] if ref.dim() == 5 else 1), ref.shape[-2], ref.shape[-1]
    n = len(planes)
    arr_p = (ctypes.c_void_p * 8)(*([p.data_ptr() for p in planes] + [0] * (8 - n)))
    arr_s = (ctypes.c_longlong * 8)(*([p.stride(0) for p in planes] + [0] * (8 - n)))
    out = torch.empty((B, D, H, W, 8), dtype=torch.bfloat16, device=ref.device)
    _lib.check(lib.vxm_planar_to_ndhwc8_bf16(arr_p, arr_s, n, _lib.ptr(out), B, D * H * W, _lib.stream_ptr()),
               "vxm_planar_to_ndhwc8_bf16")
    return out


def kdfold_enabled():
    """kd-folded execution of the two layers with 2 / 3 real channels on one side (VXM_B200_KDFOLD=0: A/B switch)."""
    import os
    return os.environ.get("VXM_B200_KDFOLD", "1") != "0"


def planar_fold_kd(planes, cout):
    """<= cout/3 planar fp32 (B,1,D,H,W) volumes -> bf16 (B,D,H,W,cout) with channel kd * n + p = plane p at slice d + kd - 1
    (zero outside the volume): the kd taps of a 3-D convolution folded into the channels (csrc/ndhwc_ops.cu)."""
    lib = _lib.load()
    ref = planes[0]
    B, D, H, W = ref.shape[0], ref.shape[-3], ref.shape[-2], ref.shape[-1]
    n = len(planes)
    arr_p = (ctypes.c_void_p * 8)(*([p.data_ptr() for p in planes] + [0] * (8 - n)))
    arr_s = (ctypes.c_longlong * 8)(*([p.stride(0) for p in planes] + [0] * (8 - n)))
    out = torch.empty((B, D, H, W, cout), dtype=torch.bfloat16, device=ref.device)
    _lib.check(lib.vxm_planar_fold_kd_bf16(arr_p, arr_s, n, _lib.ptr(out), B, D, H * W, cout, _lib.stream_ptr()), "vxm_planar_fold_kd_bf16")
    return out


def pack_weights_fold(w, transposed=False):
    """Packed kd-folded 2-D operand of the 3-D weight w (Cout, Cin, 3, 3, 3) (see vxm_conv3d_tcs_pack_desc_fold).
    Returns (tensor, (coutp, "s")).  Stand-alone helper (tests / tools); the engine packs through its _PackPlan."""
    lib = _lib.load()
    w = w.contiguous()
    Cout, Cin = w.shape[0], w.shape[1]
    real_in, nout = (Cout, Cin) if transposed else (Cin, Cout)
    coutp = 16 if nout <= 16 else 32
    nbytes = int(lib.vxm_conv3d_tcs_packed_bytes(3 * real_in, coutp, 1))
    out = torch.empty(nbytes // 2, dtype=torch.bfloat16, device=w.device)
    dsz = int(lib.vxm_conv3d_tcs_pack_desc_bytes())
    host = ctypes.create_string_buffer(dsz)
    cnt = lib.vxm_conv3d_tcs_pack_desc_fold(ctypes.cast(host, ctypes.c_void_p), _lib.ptr(w), _lib.ptr(out), Cout, Cin, coutp, 1 if transposed else 0, 0)
    if cnt <= 0:
        raise _lib.VxmError("vxm_conv3d_tcs_pack_desc_fold: %s" % _lib.last_error())
    descs = torch.frombuffer(bytearray(host.raw), dtype=torch.uint8).to(w.device)
    _lib.check(lib.vxm_conv3d_tcs_pack_multi(_lib.ptr(descs), 1, cnt, _lib.stream_ptr()), "vxm_conv3d_tcs_pack_multi")
    torch.cuda.current_stream(w.device).synchronize()       # `descs` must outlive the launch
    return out, (coutp, "s")


# ---- weights-stationary ("transposed") kernel: Cin in {8,16,32,48}, Cout <= 32 -------------------------------------

def _variant():
    """'s' = kw-stacked kernel with swizzled operands, 't' = kw-stacked with SWIZZLE_NONE operands, 'n' = one MMA per tap."""
    import os
    return os.environ.get("VXM_B200_TC_KERNEL", "auto")


def use_t_kernel(ca, cb, cout):
    """True when this convolution runs on a kw-stacked kernel (VXM_B200_TC_KERNEL = auto | s | t | n)."""
    mode = _variant()
    if mode == "n":
        return False
    lib = _lib.load()
    if mode in ("auto", "s") and lib.vxm_conv3d_tcs_supported(ca, cb, cout):
        return True
    return bool(lib.vxm_conv3d_tct_supported(ca, cb, cout))


def _use_s(ca, cb, cout):
    return _variant() in ("auto", "s") and bool(_lib.load().vxm_conv3d_tcs_supported(ca, cb, cout))


def _use_s2(xa, xb, coutp, full, up, kd):
    """The two-issuer variant (conv3d_tc_s2.cu) wherever the one-issuer kernel would pick 8-row tiles; VXM_B200_TCS2=0
    switches back to the single issuer (A/B)."""
    import os
    if os.environ.get("VXM_B200_TCS2", "1") != "1":      # default on: +5 % step throughput on B200 (profiles/r2_tcs2_ab.md)
        return False
    cin = (0 if xa is None else xa.shape[-1]) + (0 if xb is None else xb.shape[-1])
    H = full.shape[2] * (2 if (xb is None and up) else 1)
    if coutp == 48:          # the single-pass dgrad of the 48-channel concat layer (32 -> 32 + 16 channels)
        return cin == 32 and H > 4 and os.environ.get("VXM_B200_TCS2_48", "1") == "1"
    return coutp in (16, 32) and cin in (8, 16, 32, 48) and H > 4 and not (cin == 48 and coutp == 32)


def pack_weights_t(w, transposed=False, variant=None):
    """Packed weights for a kw-stacked kernel.  Returns (tensor, (coutp, variant))."""
    lib = _lib.load()
    if w.dim() == 4:
        w = w.unsqueeze(2)
    w = w.contiguous()
    Cout, Cin, kd = w.shape[0], w.shape[1], w.shape[2]
    cin_eff, nout = (Cout, Cin) if transposed else (Cin, Cout)
    coutp = 16 if nout <= 16 else (32 if nout <= 32 else (48 if nout <= 48 else 64))
    if variant is None:
        variant = "s" if _variant() in ("auto", "s") else "t"
    fb, fp = (lib.vxm_conv3d_tcs_packed_bytes, lib.vxm_conv3d_tcs_pack) if variant == "s" else \
             (lib.vxm_conv3d_tct_packed_bytes, lib.vxm_conv3d_tct_pack)
    nbytes = int(fb(cin_eff, coutp, kd))
    out = torch.empty(nbytes // 2, dtype=torch.bfloat16, device=w.device)
    _lib.check(fp(_lib.ptr(w), _lib.ptr(out), Cout, Cin, kd, coutp, 1 if transposed else 0, _lib.stream_ptr()),
               "vxm_conv3d_tc%s_pack" % variant)
    return out, (coutp, variant)


def conv_fwd_t(xa, xb, wpk, coutp, bias, cout, kd, up=False, out_fp32_planar=False, slope=None, mask=None, split=None):
    lib = _lib.load()
    coutp, variant = coutp if isinstance(coutp, tuple) else (coutp, "t")
    fwd = lib.vxm_conv3d_tcs_fwd if variant == "s" else lib.vxm_conv3d_tct_fwd
    full = xb if xb is not None else xa
    if variant == "s" and _use_s2(xa, xb, coutp, full, up, kd):
        fwd = lib.vxm_conv3d_tcs2_fwd
    B, D, H, W = full.shape[0], full.shape[1], full.shape[2], full.shape[3]
    if xb is None and up:
        D, H, W = (D * 2 if kd == 3 else D), H * 2, W * 2
    Ca = 0 if xa is None else xa.shape[-1]
    Cb = 0 if xb is None else xb.shape[-1]
    dev = full.device
    out2 = None
    if out_fp32_planar:
        out = torch.empty((B, cout, D, H, W), dtype=torch.float32, device=dev)
    elif split:
        out = torch.empty((B, D, H, W, split), dtype=torch.bfloat16, device=dev)
        out2 = torch.empty((B, D, H, W, cout - split), dtype=torch.bfloat16, device=dev)
    else:
        out = torch.empty((B, D, H, W, cout), dtype=torch.bfloat16, device=dev)
    s = -1.0 if slope is None else float(slope)
    _lib.check(fwd(_lib.ptr(xa), _lib.ptr(xb), _lib.ptr(wpk), _lib.ptr(bias), _lib.ptr(out), _lib.ptr(mask),
                   B, D, H, W, Ca, Cb, 1 if up else 0, cout, coutp, kd, 1 if out_fp32_planar else 0, s,
                   _lib.ptr(out2), int(split or 0), _lib.stream_ptr()), "vxm_conv3d_tc%s_fwd" % variant)
    return (out, out2) if split else out


# ---- split-precision (bf16x3) passes ----------------------------------------------------------------------------------

def split_weights(w):
    """fp32 weights -> (hi, lo) fp32 tensors with hi = bf16(w), lo = w - hi (packed as bf16 by the weight packer)."""
    hi = w.detach().to(torch.bfloat16).float()
    return hi, w.detach() - hi


def planar_to_ndhwc8_split(planes):
    """<= 8 planar fp32 volumes -> (hi, lo) bf16 (B,D,H,W,8) tensors with hi + lo = x to 16 mantissa bits."""
    lib = _lib.load()
    ref = planes[0]
    B = ref.shape[0]
    D, H, W = (ref.shape[-3] if ref.dim() == 5 else 1), ref.shape[-2], ref.shape[-1]
    n = len(planes)
    arr_p = (ctypes.c_void_p * 8)(*([p.data_ptr() for p in planes] + [0] * (8 - n)))
    arr_s = (ctypes.c_longlong * 8)(*([p.stride(0) for p in planes] + [0] * (8 - n)))
    hi = torch.empty((B, D, H, W, 8), dtype=torch.bfloat16, device=ref.device)
    lo = torch.empty_like(hi)
    _lib.check(lib.vxm_planar_to_ndhwc8_split_bf16(arr_p, arr_s, n, _lib.ptr(hi), _lib.ptr(lo), B, D * H * W, _lib.stream_ptr()),
               "vxm_planar_to_ndhwc8_split_bf16")
    return hi, lo


def conv_fwd_split(xa, xb, packs, bias, cout, kd, up=False, out_fp32_planar=False, slope=None):
    """One convolution layer in split precision: three passes of the kw-stacked tcgen05 kernel accumulating
    x_lo*w_hi + x_hi*w_lo + x_hi*w_hi in an fp32 channels-last buffer.  xa / xb: (hi, lo) pairs (or None);
    packs: ((wpk_hi, meta), (wpk_lo, meta)) from pack_weights_t.  Returns a (hi, lo) pair of bf16 NDHWC tensors, or the
    fp32 planar tensor when out_fp32_planar."""
    lib = _lib.load()
    (wh, (coutp, variant)), (wl, _) = packs
    if variant != "s":
        raise _lib.VxmError("split-precision convolution needs the swizzled kw-stacked kernel (VXM_B200_TC_KERNEL=auto|s)")
    full = xb if xb is not None else xa
    fh = full[0]
    B, D, H, W = fh.shape[0], fh.shape[1], fh.shape[2], fh.shape[3]
    if xb is None and up:
        D, H, W = (D * 2 if kd == 3 else D), H * 2, W * 2
    Ca = 0 if xa is None else xa[0].shape[-1]
    Cb = 0 if xb is None else xb[0].shape[-1]
    dev = fh.device
    acc = torch.empty((B, D, H, W, coutp), dtype=torch.float32, device=dev)
    s = -1.0 if slope is None else float(slope)

    def part(x, k):
        return None if x is None else x[k]

    def launch(k_x, wpk, out, out_lo, acc_in, mode):
        _lib.check(lib.vxm_conv3d_tcs_fwd_acc(_lib.ptr(part(xa, k_x)), _lib.ptr(part(xb, k_x)), _lib.ptr(wpk), _lib.ptr(bias),
                                              _lib.ptr(out), _lib.ptr(out_lo), _lib.ptr(acc_in), B, D, H, W, Ca, Cb,
                                              1 if up else 0, cout, coutp, kd, mode, s, _lib.stream_ptr()),
                   "vxm_conv3d_tcs_fwd_acc")

    launch(1, wh, acc, None, None, 2)          # x_lo * w_hi   (smallest terms first)
    launch(0, wl, acc, None, acc, 2)           # + x_hi * w_lo
    if out_fp32_planar:
        out = torch.empty((B, cout, D, H, W), dtype=torch.float32, device=dev)
        launch(0, wh, out, None, acc, 1)       # + x_hi * w_hi + bias -> fp32 planar
        return out
    hi = torch.empty((B, D, H, W, cout), dtype=torch.bfloat16, device=dev)
    lo = torch.empty_like(hi)
    launch(0, wh, hi, lo, acc, 3)              # + x_hi * w_hi + bias, activation -> (hi, lo)
    return hi, lo


def pool_split(x, nd):
    lib = _lib.load()
    xh, xl = x
    B, D, H, W, C = xh.shape
    Dc = D // 2 if nd == 3 else D
    yh = torch.empty((B, Dc, H // 2, W // 2, C), dtype=torch.bfloat16, device=xh.device)
    yl = torch.empty_like(yh)
    _lib.check(lib.vxm_pool2_split_ndhwc_bf16(_lib.ptr(xh), _lib.ptr(xl), _lib.ptr(yh), _lib.ptr(yl), B, Dc, H // 2, W // 2, C, nd,
                                              _lib.stream_ptr()), "vxm_pool2_split_ndhwc_bf16")
    return yh, yl


# ---- deferred weight-gradient reduction: every layer's partials reduced by ONE launch at the end of the backward pass ----

class WgradBatch:
    """Collects the pending reductions of the tcgen05 weight-gradient kernels of one backward pass
    (vxm_conv3d_tc_wgrad2_partial) and reduces them all in one launch (`flush`).  One instance per (device, stream);
    the workspace is persistent (512 MB: the default 3-D U-Net at 160x192x224 needs ~200 MB of per-CTA partials)."""

    _cache = {}
    WORK_BYTES = 512 << 20

    @classmethod
    def get(cls, device):
        key = (device.index, torch.cuda.current_stream(device).cuda_stream)
        b = cls._cache.get(key)
        if b is None:
            with _ws_lock:
                b = cls._cache.get(key)
                if b is None:
                    b = cls._cache[key] = cls(device)
        return b

    def __init__(self, device):
        lib = _lib.load()
        self.dev = device
        self.dsz = int(lib.vxm_conv3d_tc_wgrad2_desc_bytes())
        self.maxn = int(lib.vxm_conv3d_tc_wgrad2_max_pending())
        self.host = ctypes.create_string_buffer(self.dsz * self.maxn)
        self.n = ctypes.c_int(0)
        self.used = ctypes.c_size_t(0)
        self.off = 0
        self.work = torch.empty(self.WORK_BYTES, dtype=torch.uint8, device=device)

    def add(self, xa, xb, gz, gw, gb, cin, cout, kd, up, accumulate):
        lib = _lib.load()
        need = int(lib.vxm_conv3d_tc_wgrad2_partial_bytes(kd))
        if self.off + need > self.WORK_BYTES or self.n.value + 2 > self.maxn:
            self.flush()
        B, D, H, W, Cg = gz.shape
        Ca = 0 if xa is None else xa.shape[-1]
        Cb = 0 if xb is None else xb.shape[-1]
        _lib.check(lib.vxm_conv3d_tc_wgrad2_partial(_lib.ptr(xa), _lib.ptr(xb), _lib.ptr(gz), _lib.ptr(gw), _lib.ptr(gb),
                                                    ctypes.c_void_p(self.work.data_ptr() + self.off), self.WORK_BYTES - self.off,
                                                    ctypes.byref(self.used), ctypes.cast(self.host, ctypes.c_void_p), ctypes.byref(self.n),
                                                    B, D, H, W, Ca, Cb, 1 if up else 0, cin, Cg, cout, kd, 1 if accumulate else 0,
                                                    _lib.stream_ptr()), "vxm_conv3d_tc_wgrad2_partial")
        self.off += int(self.used.value)

    def add_khm(self, x, gz, gw2d, gb, cin_real, cout_real):
        """Weight gradient of a kd-folded layer: x (B,D,H,W,Cx) against gz (B,D,H,W,Cg) -> gw2d (cout_real, cin_real, 1, 3, 3)
        (overwritten at flush), gb (cout_real) or None."""
        lib = _lib.load()
        need = int(lib.vxm_conv3d_tc_wgrad2_partial_bytes(1))
        if self.off + need > self.WORK_BYTES or self.n.value + 1 > self.maxn:
            self.flush()
        B, D, H, W, Cg = gz.shape
        _lib.check(lib.vxm_conv3d_tc_wgrad2_partial_khm(_lib.ptr(x), _lib.ptr(gz), _lib.ptr(gw2d), _lib.ptr(gb),
                                                        ctypes.c_void_p(self.work.data_ptr() + self.off), self.WORK_BYTES - self.off,
                                                        ctypes.byref(self.used), ctypes.cast(self.host, ctypes.c_void_p), ctypes.byref(self.n),
                                                        B, D, H, W, x.shape[-1], cin_real, Cg, cout_real, 0, _lib.stream_ptr()),
                   "vxm_conv3d_tc_wgrad2_partial_khm")
        self.off += int(self.used.value)

    def reset(self):
        """Drop pending reductions (after an error mid-backward)."""
        self.n.value = 0
        self.off = 0

    def flush(self):
        if self.n.value:
            _lib.check(_lib.load().vxm_conv3d_tc_wgrad2_flush(ctypes.cast(self.host, ctypes.c_void_p), self.n.value, _lib.stream_ptr()),
                       "vxm_conv3d_tc_wgrad2_flush")
        self.n.value = 0
        self.off = 0


def wgrad_deferrable(xa, xb, gz, planar_x=None, planar_g=None):
    import os
    if os.environ.get("VXM_B200_WGRAD_DEFER", "1") != "1" or os.environ.get("VXM_B200_WGRAD", "")[:1] == "o":
        return False
    if gz is None or planar_x is not None or planar_g is not None:
        return False
    ok = lambda c: c in (8, 16, 32)  # noqa: E731
    ca = 0 if xa is None else xa.shape[-1]
    cb = 0 if xb is None else xb.shape[-1]
    return (ca == 0 or ok(ca)) and (cb == 0 or ok(cb)) and ca + cb > 0 and ok(gz.shape[-1])
