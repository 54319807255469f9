"""
Host-side mirror of the reference network classes (orca_modules.py), backed by liborca_b200.so.

Each class keeps the reference's constructor arguments, ``forward`` signature and
``state_dict`` keys (the parameter containers are the same ``nn.Sequential`` /
``nn.ModuleList`` trees, so reference ``.statedict`` files load with ``strict=True``), but
``forward`` never runs a torch.nn layer: it hands raw device pointers to the C ABI declared
in ``include/orca_b200.h``.  There is no CPU path -- CPU tensors raise ``RuntimeError``.

Reference classes mirrored (file:line in /root/reference/orca_modules.py):
  Encoder :803-980, Encoder2 :984-1169, Encoder2b :1173-1276, Encoder3 :1279-1406,
  Decoder :16-488, Decoder_1m :491-800, Net :1409-1900.
"""
import ctypes

import torch
from torch import nn

from . import _lib

# Decoder dilation schedules (orca_modules.py:22-459 and :499-774)
DECODER_DILATIONS = [1, 2, 4, 8, 16, 32, 64] * 4
DECODER_1M_DILATIONS = [1, 2, 4, 8, 16, 32, 64] + [2, 4, 8, 16, 32, 64] * 2
# Encoder stage table: (pool before the stage, c_in, c_out)   orca_modules.py:811-927
ENCODER_STAGES = [(None, 4, 64), (4, 64, 96), (4, 96, 128), (5, 128, 128), (5, 128, 128),
                  (5, 128, 128), (2, 128, 128)]


# ----------------------------------------------------------------------------------------
# parameter containers (structure only; never executed)
# ----------------------------------------------------------------------------------------
def _seq1d_linear(c_in, c_out, head=None):
    """[head] Conv BN Conv BN  -- the 'l' (linear) half of a residual unit."""
    layers = [] if head is None else [head]
    layers += [nn.Conv1d(c_in, c_out, kernel_size=9, padding=4), nn.BatchNorm1d(c_out),
               nn.Conv1d(c_out, c_out, kernel_size=9, padding=4), nn.BatchNorm1d(c_out)]
    return nn.Sequential(*layers)


def _seq1d_relu(c, last_bn=True):
    """Conv BN ReLU Conv [BN] ReLU."""
    layers = [nn.Conv1d(c, c, kernel_size=9, padding=4), nn.BatchNorm1d(c), nn.ReLU(inplace=True),
              nn.Conv1d(c, c, kernel_size=9, padding=4)]
    if last_bn:
        layers.append(nn.BatchNorm1d(c))
    layers.append(nn.ReLU(inplace=True))
    return nn.Sequential(*layers)


def _seq2d_linear(c_in, c_mid, c_out, d, dropout=False):
    layers = [nn.Dropout(p=0.1)] if dropout else []
    layers += [nn.Conv2d(c_in, c_mid, kernel_size=(3, 3), padding=d, dilation=d), nn.BatchNorm2d(c_mid),
               nn.Conv2d(c_mid, c_out, kernel_size=(3, 3), padding=d, dilation=d), nn.BatchNorm2d(c_out)]
    return nn.Sequential(*layers)


def _seq2d_relu(c_in, c_mid, c_out, d):
    return nn.Sequential(
        nn.Conv2d(c_in, c_mid, kernel_size=(3, 3), padding=d, dilation=d), nn.BatchNorm2d(c_mid),
        nn.ReLU(inplace=True),
        nn.Conv2d(c_mid, c_out, kernel_size=(3, 3), padding=d, dilation=d), nn.BatchNorm2d(c_out),
        nn.ReLU(inplace=True))


def _final2d(num_2d=1):
    """orca_modules.py:423-428 (num_2d = 1); orca_leukemia.py:923-926: hidden width max(num_2d, 5)."""
    h = num_2d if num_2d > 5 else 5
    return nn.Sequential(nn.Conv2d(64, h, kernel_size=(1, 1), padding=0), nn.BatchNorm2d(h),
                         nn.ReLU(inplace=True), nn.Conv2d(h, num_2d, kernel_size=(1, 1), padding=0))


def _check_num_2d(num_2d):
    if not isinstance(num_2d, int) or not 1 <= num_2d <= 8:
        raise ValueError("num_2d must be an int in [1, 8], got %r" % (num_2d,))
    return num_2d


def _add_encoder_stages(mod):
    for k, (pool, c_in, c_out) in enumerate(ENCODER_STAGES, start=1):
        head = None if pool is None else nn.MaxPool1d(kernel_size=pool, stride=pool)
        setattr(mod, "lconv%d" % k, _seq1d_linear(c_in, c_out, head))
        setattr(mod, "conv%d" % k, _seq1d_relu(c_out))


def _add_decoder_1m_body(mod, num_2d=1):
    dil = DECODER_1M_DILATIONS
    mod.lconvtwos = nn.ModuleList(
        [_seq2d_linear(128 if i == 0 else 64, 32, 64, d, dropout=(i == 0)) for i, d in enumerate(dil)])
    mod.convtwos = nn.ModuleList([_seq2d_relu(64, 32, 64, d) for d in dil])
    mod.final = _final2d(num_2d)


# ----------------------------------------------------------------------------------------
# conv extraction: walk the containers in definition order, pair each conv with its BN
# ----------------------------------------------------------------------------------------
def _conv_entries(seq):
    mods = list(seq)
    out = []
    for i, m in enumerate(mods):
        if isinstance(m, (nn.Conv1d, nn.Conv2d)):
            bn = mods[i + 1] if i + 1 < len(mods) and isinstance(mods[i + 1], (nn.BatchNorm1d, nn.BatchNorm2d)) else None
            out.append((m, bn))
    return out


class _NativeModule(nn.Module):
    """Common plumbing: lazy handle creation, invalidation on weight changes."""

    _kind = None

    def __init__(self):
        super().__init__()
        self._handle = None
        self._handle_device = None
        self._handle_version = None
        self._handle_options = None
        self._calibrated_version = None
        # per-module overrides of the process defaults in orca_b200._lib (keys: "impl", "encoder_fp16_stages")
        self.options = {}
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._invalidate())

    # -- to be provided by subclasses --------------------------------------------------
    def _sequentials(self):
        raise NotImplementedError

    def _flags(self):
        return 0

    def _num_1d(self):
        return 0

    # -- handle management ---------------------------------------------------------------
    def _invalidate(self):
        if self._handle is not None:
            _lib.lib().orca_b200_module_destroy(self._handle)
        self._handle = None

    def __del__(self):
        try:
            self._invalidate()
        except Exception:
            pass

    def _param_version(self):
        return tuple(t._version for t in list(self.parameters()) + list(self.buffers()))

    def _apply(self, fn, recurse=True):
        self._invalidate()
        return super()._apply(fn, recurse)

    def train(self, mode=True):
        if mode:
            raise RuntimeError("orca_b200 modules are inference-only (eval-mode BatchNorm/Dropout "
                               "semantics are folded into the kernels); call .eval()")
        return super().train(False)

    def native_handle(self, device):
        """Fold + pack + upload the weights on first use (or after they changed)."""
        version = self._param_version()
        if self._handle is not None and self._handle_device == device and self._handle_version == version:
            stamp = (_lib.options_epoch, tuple(sorted(self.options.items())))
            if stamp != self._handle_options:  # kernel selection / precision are per handle in the C library
                _lib.apply_options(self._handle, self.options)
                self._handle_options = stamp
            return self._handle
        self._invalidate()
        entries = []
        for seq in self._sequentials():
            entries += _conv_entries(seq)
        keep = []  # host copies must outlive the create call
        arr = (_lib.ConvParams * len(entries))()

        def host(t):
            t = t.detach().to("cpu", torch.float32).contiguous()
            keep.append(t)
            return ctypes.cast(t.data_ptr(), ctypes.POINTER(ctypes.c_float))

        for p, (conv, bn) in zip(arr, entries):
            ks = conv.kernel_size
            p.c_in, p.c_out = conv.in_channels, conv.out_channels
            p.kh, p.kw = (1, ks[0]) if len(ks) == 1 else (ks[0], ks[1])
            p.dilation = conv.dilation[0]
            p.weight = host(conv.weight)
            p.bias = host(conv.bias) if conv.bias is not None else None
            if bn is not None:
                p.bn_weight, p.bn_bias = host(bn.weight), host(bn.bias)
                p.bn_mean, p.bn_var = host(bn.running_mean), host(bn.running_var)
                p.bn_eps = bn.eps
        handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(_lib.lib().orca_b200_module_create(self._kind, arr, len(entries), self._flags(),
                                                          self._num_1d(), ctypes.byref(handle)))
        self._handle, self._handle_device, self._handle_version = handle, device, version
        _lib.apply_options(handle, self.options)
        self._handle_options = (_lib.options_epoch, tuple(sorted(self.options.items())))
        return handle

    # -- fp16 range guard (modules with encoder stages) -------------------------------------
    def fp16_guard_fired(self, clear=True):
        """True if a single-pass fp16 encoder stage of this module produced a value beyond the fp16 range guard since
        the last check (orca_b200_module_status; synchronises the device)."""
        if self._handle is None:
            return False
        with torch.cuda.device(self._handle_device):
            return bool(_lib.module_status(self._handle, clear) & _lib.STATUS_FP16_RANGE)

    CALIBRATION_BP = 96000     # length of the window the one-time precision self-check runs on
    CALIBRATION_TOL = 2e-4     # single-pass vs fp32-grade, max-abs / max: synthetic default weights give ~1e-5

    def _calibrate_fp16(self, run_small):
        """One-time self-check per set of weights (first forward after they were loaded): run a small window of the real
        input with the default single-pass fp16 stages AND with every stage in the fp32-grade format; keep the fast format only
        if the two agree to CALIBRATION_TOL.  This catches what no static rule does (e.g. heavy-tailed trained weights
        whose few dominant taps stop the fp16 rounding noise from averaging out).  `run_small()` performs the small
        forward under the module's current options and returns one output tensor."""
        if self._calibrated_version == self._handle_version or "encoder_fp16_stages" in self.options:
            return
        self._calibrated_version = self._handle_version
        handle = self._handle
        if _lib.lib().orca_b200_module_get_option(handle, _lib.OPT_ENCODER_FP16_STAGES) == 0:
            return  # already fp32-grade (the library found the folded weights ill-conditioned, or the default is 0)
        fast = run_small()
        self.options["encoder_fp16_stages"] = 0
        try:
            exact = run_small()
        finally:
            del self.options["encoder_fp16_stages"]
        scale = float(exact.abs().max())
        diff = float((fast - exact).abs().max())
        fired = self.fp16_guard_fired()
        if fired or not (diff <= self.CALIBRATION_TOL * max(scale, 1e-30)):
            import warnings
            warnings.warn("orca_b200: %s runs every encoder stage in the fp32-grade format: on a %d bp window the single-pass "
                          "fp16 stages differ by %.1e of the output range (limit %.0e)%s"
                          % (type(self).__name__, self.CALIBRATION_BP, diff / max(scale, 1e-30), self.CALIBRATION_TOL,
                             "; fp16 range exceeded" if fired else ""), RuntimeWarning, stacklevel=4)
            self.options["encoder_fp16_stages"] = 0

    def _fall_back_to_fp32_grade(self):
        import warnings
        warnings.warn("orca_b200: an activation of %s exceeded the fp16 range in a single-pass encoder stage; this module "
                      "now runs every stage in the three-product fp32-grade format (encoder_fp16_stages = 0)"
                      % type(self).__name__, RuntimeWarning, stacklevel=3)
        self.options["encoder_fp16_stages"] = 0


def _require_cuda(name, t, dtypes=(torch.float32,)):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("orca_b200.%s: expected a CUDA tensor (there is no CPU path)" % name)
    if t.dtype not in dtypes:
        raise RuntimeError("orca_b200.%s: expected %s, got %s" % (name, " or ".join(str(d) for d in dtypes), t.dtype))


def _workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


# ----------------------------------------------------------------------------------------
class Encoder(_NativeModule):
    """bp -> 4 kb encoder; mirrors orca_modules.Encoder (:803-980).

    forward(x: f32[B, 4, L]) -> f32[B, 128, L // 4000]   (a transposed view of channel-last storage)
    """
    _kind = _lib.ENCODER

    def __init__(self):
        super().__init__()
        _add_encoder_stages(self)
        self.chunk_bp = 0  # 0 = library default
        self.eval()

    def _sequentials(self):
        out = []
        for k in range(1, 8):
            out += [getattr(self, "lconv%d" % k), getattr(self, "conv%d" % k)]
        return out

    def forward(self, x, bin_range=None, out=None, reverse_complement=False, window=None, guard=True):
        """Reference call: forward(x).  Extensions used by orca_b200.predict / orca_b200.parallel:

        guard=True checks the fp16 range guard after the call (one device synchronisation) and, if it fired, reruns
            the call with every stage in the fp32-grade format and keeps this module there.  The batched drivers pass
            guard=False and check once per pass (predict.check_fp16_guard) so that nothing synchronises mid-pass.

        reverse_complement=True encodes the opposite strand straight from the same buffer:
            RC(x)[b, c, l] = x[b, 3-c, L-1-l] (orca_predict.py:324-329) is x walked with negated strides.
        bin_range=(b0, b1) computes only those 4 kb bins (a sequence shard); `out` is an existing
            (B, L/4000, 128) channel-last buffer to fill.
        window=(pos0, L_total): x holds only forward-strand positions [pos0, pos0 + x.size(2)) of a
            sequence of L_total bp (a shard uploads its slice plus the 112 kb halo, not the whole input).
        x may also be PACKED bases: a uint8 (B, L) tensor of codes 0..4 / raw ASCII (orca_b200.feeder), 1 B/bp.
        """
        if x.is_cuda and x.size(-1) >= 4000:
            self.native_handle(x.device)
            n_cal = min(self.CALIBRATION_BP, (x.size(-1) // 4000) * 4000)
            self._calibrate_fp16(lambda: self._forward(x[..., :n_cal].contiguous() if x.dtype != torch.uint8 else x[:, :n_cal].contiguous()))
        res = self._forward(x, bin_range, out, reverse_complement, window)
        if guard and self.fp16_guard_fired():
            self._fall_back_to_fp32_grade()
            res = self._forward(x, bin_range, out, reverse_complement, window)
        return res

    def _forward(self, x, bin_range=None, out=None, reverse_complement=False, window=None):
        """One native Encoder call; see forward."""
        _require_cuda("Encoder.forward", x, (torch.float32, torch.uint8))
        packed = x.dtype == torch.uint8
        n = x.size(-1)
        pos0, L = (0, n) if window is None else window
        if (x.dim() != 2 if packed else (x.dim() != 3 or x.size(1) != 4)) or L % 4000 != 0 or L == 0 or n == 0:
            raise RuntimeError("Encoder.forward: expected float32 (B, 4, L) or packed uint8 (B, L) with L a positive "
                               "multiple of 4000, got %s %s" % (x.dtype, tuple(x.shape)))
        B = x.shape[0]
        P = L // 4000
        dev = x.device
        h = self.native_handle(dev)
        lib = _lib.lib()
        with torch.cuda.device(dev):
            if out is None:
                out = torch.empty((B, P, 128), dtype=torch.float32, device=dev)
            elif tuple(out.shape) != (B, P, 128) or not out.is_contiguous():
                raise RuntimeError("Encoder.forward: out must be a contiguous (B, L/4000, 128) tensor")
            b0, b1 = (0, P) if bin_range is None else bin_range
            ws = _workspace(lib.orca_b200_encoder_workspace_bytes(h, B, L, self.chunk_bp), dev)
            xp = x.data_ptr()
            if packed:
                sB, sL = x.stride()
                if reverse_complement:  # walk the same bytes backwards, complementing the codes on the fly
                    xp += (n - 1) * sL
                    sL = -sL
                    pos0 = L - (pos0 + n)
                _lib.check(lib.orca_b200_encoder_forward_packed(h, ctypes.c_void_p(xp), B, L, sB, sL,
                                                                1 if reverse_complement else 0, pos0, n, _ptr(out), b0, b1,
                                                                self.chunk_bp, _ptr(ws), ws.numel(), _stream(dev)))
                return out.transpose(1, 2)
            sB, sC, sL = x.stride()
            if reverse_complement:
                xp += 4 * (3 * sC + (n - 1) * sL)
                sC, sL = -sC, -sL
                pos0 = L - (pos0 + n)
            _lib.check(lib.orca_b200_encoder_forward(h, ctypes.c_void_p(xp), B, L, sB, sC, sL, pos0, n,
                                                     _ptr(out), b0, b1, self.chunk_bp, _ptr(ws), ws.numel(),
                                                     _stream(dev)))
        return out.transpose(1, 2)


class _UNet1d(_NativeModule):
    """Shared forward of Encoder2 / Encoder2b / Encoder3."""
    _n_out = 6

    def forward(self, x, coarsest_only=False):
        _require_cuda(type(self).__name__ + ".forward", x)
        div = 1 << (self._n_out - 1)
        if x.dim() != 3 or x.size(1) != 128 or x.size(2) % div != 0 or x.size(2) == 0:
            raise RuntimeError("%s.forward: expected (B, 128, P) with P a positive multiple of %d, got %s"
                               % (type(self).__name__, div, tuple(x.shape)))
        B, _, P = x.shape
        dev = x.device
        h = self.native_handle(dev)
        lib = _lib.lib()
        with torch.cuda.device(dev):
            outs = [torch.empty((B, P >> i, 128), dtype=torch.float32, device=dev) for i in range(self._n_out)]
            ptrs = (ctypes.c_void_p * self._n_out)(*[o.data_ptr() for o in outs])
            ws = _workspace(lib.orca_b200_encoder2_workspace_bytes(h, B, P), dev)
            _lib.check(lib.orca_b200_encoder2_forward(h, _ptr(x), B, P, x.stride(0), x.stride(1), x.stride(2),
                                                      ptrs, self._n_out, 1 if coarsest_only else 0,
                                                      _ptr(ws), ws.numel(), _stream(dev)))
        if coarsest_only:
            return [outs[-1].transpose(1, 2)]
        return [o.transpose(1, 2) for o in outs]


class Encoder2(_UNet1d):
    """4 kb -> 128 kb U-net; mirrors orca_modules.Encoder2 (:984-1169)."""
    _kind = _lib.ENCODER2
    _n_out = 6

    def __init__(self):
        super().__init__()
        n = self._n_out - 1
        self.lblocks = nn.ModuleList([_seq1d_linear(128, 128, nn.MaxPool1d(kernel_size=2, stride=2)) for _ in range(n)])
        self.blocks = nn.ModuleList([_seq1d_relu(128) for _ in range(n)])
        self.downlblocks = nn.ModuleList([_seq1d_linear(128, 128, nn.Upsample(scale_factor=2)) for _ in range(n)])
        self.downblocks = nn.ModuleList([_seq1d_relu(128, last_bn=False) for _ in range(n)])
        self.eval()

    def _sequentials(self):
        return list(self.lblocks) + list(self.blocks) + list(self.downlblocks) + list(self.downblocks)


class Encoder3(Encoder2):
    """128 kb -> 1024 kb U-net; mirrors orca_modules.Encoder3 (:1279-1406)."""
    _kind = _lib.ENCODER3
    _n_out = 4


class Encoder2b(_UNet1d):
    """Pooling half only (HCTnoc); mirrors orca_modules.Encoder2b (:1173-1276)."""
    _kind = _lib.ENCODER2B
    _n_out = 6

    def __init__(self):
        super().__init__()
        self.lblocks = nn.ModuleList([_seq1d_linear(128, 128, nn.MaxPool1d(kernel_size=2, stride=2)) for _ in range(5)])
        self.blocks = nn.ModuleList([_seq1d_relu(128) for _ in range(5)])
        self.eval()

    def _sequentials(self):
        return list(self.lblocks) + list(self.blocks)


class Decoder(_NativeModule):
    """2D dilated-conv decoder head; mirrors orca_modules.Decoder (:16-488) and, with num_2d > 1, the
    multi-map orca_leukemia.Decoder (orca_leukemia.py:512-993; see orca_b200.leukemia for its signature).

    forward(x f32[B,128,S], distenc f32[B,C,S,S], y f32[B,C,S/2,S/2] | None) -> f32[B,C,S,S]   (C = num_2d)
    """
    _kind = _lib.DECODER

    def __init__(self, upsample_mode="nearest", num_2d=1):
        super().__init__()
        if upsample_mode not in ("nearest", "bilinear"):
            raise ValueError("upsample_mode must be 'nearest' or 'bilinear'")
        self.num_2d = _check_num_2d(num_2d)
        dil = DECODER_DILATIONS
        self.lconvtwos = nn.ModuleList([_seq2d_linear(64, 32, 64, d, dropout=(i == 0)) for i, d in enumerate(dil)])
        self.convtwos = nn.ModuleList([_seq2d_relu(64, 32, 64, d) for d in dil])
        self.final = _final2d(num_2d)
        self.upsample = nn.Upsample(scale_factor=(2, 2), mode=upsample_mode)
        self.lcombiner = _seq2d_linear(64 + num_2d, 64, 64, 1, dropout=True)
        self.combiner = _seq2d_relu(64, 64, 64, 1)
        self.lcombinerD = _seq2d_linear(128 + num_2d, 64, 64, 1)
        self.combinerD = _seq2d_relu(64, 64, 64, 1)
        self.eval()

    def _sequentials(self):
        return (list(self.lconvtwos) + list(self.convtwos)
                + [self.final, self.lcombiner, self.combiner, self.lcombinerD, self.combinerD])

    def _flags(self):
        return _lib.UPSAMPLE_BILINEAR if self.upsample.mode == "bilinear" else _lib.UPSAMPLE_NEAREST

    def forward(self, x, distenc, y=None):
        _require_cuda("Decoder.forward(x)", x)
        _require_cuda("Decoder.forward(distenc)", distenc)
        if x.dim() != 3 or x.size(1) != 128:
            raise RuntimeError("Decoder.forward: x must be (B, 128, S), got %s" % (tuple(x.shape),))
        B, _, S = x.shape
        C = self.num_2d
        if tuple(distenc.shape) != (B, C, S, S):
            raise RuntimeError("Decoder.forward: distenc must be (%d, %d, %d, %d), got %s" % (B, C, S, S, tuple(distenc.shape)))
        if y is not None:
            _require_cuda("Decoder.forward(y)", y)
            if S % 2 or tuple(y.shape) != (B, C, S // 2, S // 2):
                raise RuntimeError("Decoder.forward: y must be (%d, %d, %d, %d), got %s" % (B, C, S // 2, S // 2, tuple(y.shape)))
        dev = x.device
        h = self.native_handle(dev)
        lib = _lib.lib()
        with torch.cuda.device(dev):
            out = torch.empty((B, C, S, S), dtype=torch.float32, device=dev)
            ws = _workspace(lib.orca_b200_decoder_workspace_bytes(h, B, S), dev)
            yp, ys = (None, (0, 0, 0, 0)) if y is None else (_ptr(y), y.stride())
            _lib.check(lib.orca_b200_decoder_forward(
                h, _ptr(x), B, S, x.stride(0), x.stride(1), x.stride(2),
                _ptr(distenc), distenc.stride(0), distenc.stride(1), distenc.stride(2), distenc.stride(3),
                yp, ys[0], ys[1], ys[2], ys[3], _ptr(out), _ptr(ws), ws.numel(), _stream(dev)))
        return out


class Decoder_1m(_NativeModule):
    """1 Mb decoder; mirrors orca_modules.Decoder_1m (:491-800) and, with num_2d > 1, orca_leukemia.Decoder_1m
    (orca_leukemia.py:996-1315).  forward(x f32[B,128,S]) -> f32[B,num_2d,S,S]"""
    _kind = _lib.DECODER_1M

    def __init__(self, num_2d=1):
        super().__init__()
        self.num_2d = _check_num_2d(num_2d)
        _add_decoder_1m_body(self, num_2d)
        self.eval()

    def _sequentials(self):
        return list(self.lconvtwos) + list(self.convtwos) + [self.final]

    def forward(self, x):
        _require_cuda("Decoder_1m.forward", x)
        if x.dim() != 3 or x.size(1) != 128:
            raise RuntimeError("Decoder_1m.forward: x must be (B, 128, S), got %s" % (tuple(x.shape),))
        B, _, S = x.shape
        dev = x.device
        h = self.native_handle(dev)
        lib = _lib.lib()
        with torch.cuda.device(dev):
            out = torch.empty((B, self.num_2d, S, S), dtype=torch.float32, device=dev)
            ws = _workspace(lib.orca_b200_decoder_workspace_bytes(h, B, S), dev)
            _lib.check(lib.orca_b200_decoder_forward(
                h, _ptr(x), B, S, x.stride(0), x.stride(1), x.stride(2), None, 0, 0, 0, 0, None, 0, 0, 0, 0,
                _ptr(out), _ptr(ws), ws.numel(), _stream(dev)))
        return out


class Net(_NativeModule):
    """Orca-1Mb (Encoder body + Decoder_1m body [+ final_1d]); mirrors orca_modules.Net (:1409-1900) and, with
    num_2d > 1, orca_leukemia.Net (orca_leukemia.py:16-509)."""
    _kind = _lib.NET

    def __init__(self, num_1d=None, num_2d=1):
        super().__init__()
        self.num_2d = _check_num_2d(num_2d)
        _add_encoder_stages(self)
        _add_decoder_1m_body(self, num_2d)
        if num_1d is not None:
            self.final_1d = nn.Sequential(
                nn.Conv1d(128, 128, kernel_size=1, padding=0), nn.BatchNorm1d(128), nn.ReLU(inplace=True),
                nn.Conv1d(128, num_1d, kernel_size=1, padding=0), nn.Sigmoid())
        self.num_1d = num_1d
        self.eval()

    def _sequentials(self):
        out = []
        for k in range(1, 8):
            out += [getattr(self, "lconv%d" % k), getattr(self, "conv%d" % k)]
        out += list(self.lconvtwos) + list(self.convtwos) + [self.final]
        if self.num_1d:
            out.append(self.final_1d)
        return out

    def _num_1d(self):
        return int(self.num_1d) if self.num_1d else 0

    def forward(self, x, guard=True):
        """x: float32 (B, 4, L) as the reference takes it, or packed uint8 (B, L) bases (orca_b200.feeder).
        guard: see Encoder.forward (screening loops pass guard=False and call fp16_guard_fired() once per batch)."""
        if x.is_cuda and x.size(-1) >= 4000:
            self.native_handle(x.device)
            n_cal = min(self.CALIBRATION_BP, (x.size(-1) // 4000) * 4000)

            def small():
                o = self._forward(x[:1, ..., :n_cal].contiguous())
                return o[0] if isinstance(o, tuple) else o
            self._calibrate_fp16(small)
        res = self._forward(x)
        if guard and self.fp16_guard_fired():
            self._fall_back_to_fp32_grade()
            res = self._forward(x)
        return res

    def _forward(self, x):
        _require_cuda("Net.forward", x, (torch.float32, torch.uint8))
        packed = x.dtype == torch.uint8
        if (x.dim() != 2 if packed else (x.dim() != 3 or x.size(1) != 4)) or x.size(-1) % 4000 != 0 or x.size(-1) == 0:
            raise RuntimeError("Net.forward: expected float32 (B, 4, L) or packed uint8 (B, L) with L a positive "
                               "multiple of 4000, got %s %s" % (x.dtype, tuple(x.shape)))
        B, L = x.shape[0], x.shape[-1]
        S = L // 4000
        dev = x.device
        h = self.native_handle(dev)
        lib = _lib.lib()
        with torch.cuda.device(dev):
            out = torch.empty((B, self.num_2d, S, S), dtype=torch.float32, device=dev)
            out1d = torch.empty((B, self.num_1d, S), dtype=torch.float32, device=dev) if self.num_1d else None
            ws = _workspace(lib.orca_b200_net_workspace_bytes(h, B, L), dev)
            o1 = None if out1d is None else _ptr(out1d)
            if packed:
                _lib.check(lib.orca_b200_net_forward_packed(h, _ptr(x), B, L, x.stride(0), x.stride(1), 0, _ptr(out), o1,
                                                            _ptr(ws), ws.numel(), _stream(dev)))
            else:
                _lib.check(lib.orca_b200_net_forward(h, _ptr(x), B, L, x.stride(0), x.stride(1), x.stride(2),
                                                     _ptr(out), o1, _ptr(ws), ws.numel(), _stream(dev)))
        # reference: `if self.num_1d: return cur, output1d else: return cur`  (:1897-1900)
        return (out, out1d) if self.num_1d else out
