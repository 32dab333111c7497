"""Drop-in for the reference's quantized Linear modules (modules.py:12-230): `Int4Linear`,
`Int8Linear`, `Any4Linear` with the same constructor arguments, parameter names / shapes
(`weight` int32, `scales_and_zeros` [k/g][n][2], `lut`, `bias`), `kernel` strings, `w_inner_k`
semantics, `reshape_weight()` and `forward()`; state_dicts are interchangeable.

Added on top: `RowShardedLinear`, the row-wise multi-GPU wrapper; `NF4Linear` / `FP4Linear` / `MX4Linear`
(the reference's TODO, modules.py:10); a checkpoint format that remembers how the weight was packed
(8(f)-4: extra state `weight_reshaped` / `w_inner_k` / `kernel`).
"""
import torch

from . import functional as F

_ops = torch.ops.tinygemm


class _PackedCheckpoint:
    """Mixin of the packed-weight modules: the checkpoint format."""

    # ---- checkpoint format ----
    # The reference keeps `weight_reshaped`, `w_inner_k` and `kernel` as plain attributes (modules.py:194): a saved
    # state_dict of a packed model cannot be loaded back, the shape of `weight` alone does not say how it was packed.
    # Here they travel as the module's extra state, and loading adopts the checkpoint's `weight` / `lut` shapes, so a
    # packed model reloads without re-packing.  Reference checkpoints (no extra state) still load: the flags keep
    # their constructor values.
    _EXTRA_KEYS = ("weight_reshaped", "w_inner_k", "kernel", "group_size")

    def get_extra_state(self):
        st = {k: getattr(self, k) for k in self._EXTRA_KEYS}
        st["format"] = "any4_b200/1"
        return st

    def set_extra_state(self, state):
        if not state:
            return
        for k in self._EXTRA_KEYS:
            if k in state:
                setattr(self, k, state[k])

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        for name in ("weight", "lut", "exponents"):
            t = state_dict.get(prefix + name)
            cur = getattr(self, name, None)
            if t is not None and cur is not None and t.shape != cur.shape:
                # packed (4-D) vs unpacked ([out][in]) weight, per-row vs global LUT: take the checkpoint's layout
                setattr(self, name, torch.nn.Parameter(torch.empty(t.shape, device=cur.device, dtype=cur.dtype),
                                                       requires_grad=cur.requires_grad))
        n_missing = len(missing_keys)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)
        extra = prefix + torch.nn.modules.module._EXTRA_STATE_KEY_SUFFIX
        if extra in missing_keys[n_missing:]:
            missing_keys.remove(extra)  # a reference checkpoint: flags stay as constructed ...
            if self.weight.dim() == 4:  # ... except that a 4-D weight can only be a packed one
                layout = getattr(self, "_PACKERS", {}).get(self.kernel, "Bint4")
                per_ik = {"Bint4": 0.5, "Aint4": 1, "Bint8": 1, "Aint8": 2}[layout]  # last dim = inner_k * this
                self.weight_reshaped, self.w_inner_k = True, int(self.weight.shape[3] / per_ik)


class _PackedLinear(_PackedCheckpoint, torch.nn.Module):
    """Shared machinery: parameters, weight packing by kernel name, bias, reshape of activations."""

    # kernel name -> (convert op suffix used by reshape_weight); None = forward-only kernel
    _PACKERS = {}
    # kernel names forward() accepts
    _FORWARD = ()
    _DEFAULT_INNER_K = 4
    _ZERO_INIT = True

    def __init__(self, in_features, out_features, bias, device, dtype, group_size, kernel, w_inner_k):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.group_size = group_size
        make = torch.zeros if self._ZERO_INIT else torch.empty
        self.weight = torch.nn.Parameter(
            make((out_features, in_features), device=device, dtype=torch.int32), requires_grad=False
        )
        self.scales_and_zeros = torch.nn.Parameter(
            make((in_features // group_size, out_features, 2), device=device, dtype=dtype)
        )
        self._init_extra(device, dtype)
        if bias:
            self.bias = torch.nn.Parameter(torch.empty(out_features, device=device, dtype=dtype))
        else:
            self.register_parameter("bias", None)
        self.kernel = kernel
        self.w_inner_k = w_inner_k
        self.weight_reshaped = False

    def _init_extra(self, device, dtype):
        pass

    def reshape_weight(self, w_inner_k=None):
        """Pack `weight` ([out][in] int32 codes) into the tensor-core layout `kernel` consumes."""
        if w_inner_k is None:
            w_inner_k = self._DEFAULT_INNER_K
        layout = self._PACKERS.get(self.kernel)
        if layout is None:
            raise ValueError(f"Unsupported kernel type {self.kernel}")
        convert = getattr(_ops, f"convert_matrix_to_m16n8k16_{layout}_layout")
        self.weight.data = convert(self.weight, w_inner_k)
        self.weight_reshaped = True
        self.w_inner_k = w_inner_k

    def _gemm(self, x2d):
        raise NotImplementedError

    def forward(self, input):
        lead = input.shape[:-1]
        x2d = input.view(-1, input.shape[-1])
        if self.kernel not in self._FORWARD:
            raise ValueError(f"Unsupported kernel type {self.kernel}")
        y = self._gemm(x2d)
        if self.bias is not None:
            y = y + self.bias
        return y.view(*lead, y.shape[-1])

    def extra_repr(self):
        return (f"in_features={self.in_features}, out_features={self.out_features}, "
                f"bias={self.bias is not None}, group_size={self.group_size}")


class Int4Linear(_PackedLinear):
    _PACKERS = {
        "linear_y_f16RM_x_f16RM_W_int4TC": "Bint4",
        "linear_y_f16RM_W_int4TC_x_f16RM": "Aint4",
        "linear_y_f16TC_x_f16TC_W_int4TC": "Bint4",
    }
    _FORWARD = (
        "linear_y_f16RM_x_f16RM_W_int4TC", "linear_y_f16RM_W_int4TC_x_f16RM",
        "linear_y_f16TC_W_int4TC_x_f16TC", "linear_y_f16TC_x_f16TC_W_int4TC",
    )

    def __init__(self, in_features, out_features, bias=True, device=None, dtype=None, group_size=128,
                 kernel="linear_y_f16RM_W_int4TC_x_f16RM", w_inner_k=4):
        super().__init__(in_features, out_features, bias, device, dtype, group_size, kernel, w_inner_k)

    def _gemm(self, x2d):
        fn = getattr(F, self.kernel)
        return fn(x2d, self.weight, self.scales_and_zeros, self.group_size, w_inner_k=self.w_inner_k,
                  reshape_weight=not self.weight_reshaped)


class Int8Linear(_PackedLinear):
    _PACKERS = {
        "linear_y_f16RM_x_f16RM_W_int8TC": "Bint8",
        "linear_y_f16RM_W_int8TC_x_f16RM": "Aint8",
    }
    _FORWARD = (
        "linear_y_f16RM_x_f16RM_W_int8TC", "linear_y_f16RM_W_int8TC_x_f16RM", "linear_y_f16TC_W_int8TC_x_f16TC",
    )
    _DEFAULT_INNER_K = 2

    def __init__(self, in_features, out_features, bias=True, device=None, dtype=None, group_size=128,
                 kernel="linear_y_f16RM_W_int8TC_x_f16RM", w_inner_k=2):
        super().__init__(in_features, out_features, bias, device, dtype, group_size, kernel, w_inner_k)

    def _gemm(self, x2d):
        fn = getattr(F, self.kernel)
        if self.kernel == "linear_y_f16TC_W_int8TC_x_f16TC":
            # the reference never passes reshape_weight here (modules.py:138)
            return fn(x2d, self.weight, self.scales_and_zeros, self.group_size, w_inner_k=self.w_inner_k)
        return fn(x2d, self.weight, self.scales_and_zeros, self.group_size, w_inner_k=self.w_inner_k,
                  reshape_weight=not self.weight_reshaped)


class Any4Linear(_PackedLinear):
    _N_BIT = 4
    _PACKERS = {
        "linear_y_f16RM_x_f16RM_W_any4TC": "Bint4",
        "linear_y_f16RM_W_any4TC_x_f16RM": "Aint4",
    }
    _FORWARD = ("linear_y_f16RM_x_f16RM_W_any4TC", "linear_y_f16RM_W_any4TC_x_f16RM")
    _ZERO_INIT = False

    @property
    def N_BIT(self):
        return self._N_BIT

    def __init__(self, in_features, out_features, bias=True, device=None, dtype=None, group_size=128,
                 kernel="linear_y_f16RM_x_f16RM_W_any4TC", w_inner_k=4, per_row=True):
        self.per_row = per_row
        self.n_bit = 4
        super().__init__(in_features, out_features, bias, device, dtype, group_size, kernel, w_inner_k)

    def __setattr__(self, name, value):
        # per_row / n_bit are set before Module.__init__ so that _init_extra can see them
        if name in ("per_row", "n_bit") and "_parameters" not in self.__dict__:
            object.__setattr__(self, name, value)
        else:
            super().__setattr__(name, value)

    def _init_extra(self, device, dtype):
        shape = (self.out_features, 2 ** self.N_BIT) if self.per_row else (2 ** self.N_BIT,)
        self.lut = torch.nn.Parameter(torch.empty(*shape, device=device, dtype=dtype))

    def _gemm(self, x2d):
        fn = getattr(F, self.kernel)
        return fn(x2d, self.weight, self.lut, self.scales_and_zeros, self.group_size, w_inner_k=self.w_inner_k,
                  reshape_weight=not self.weight_reshaped)

    def bind_host(self, input, out=None):
        """Decode step for HOST activations (include/tinygemm_b200.h: tg_gemm_w4_rm_hostio).  `input`: pinned CPU tensor
        [m][in_features] (m small); `out`: pinned CPU tensor [m][padded out_features] (allocated if None).  Returns
        (launch, out): `launch()` enqueues, on the current stream of the weight's device, a one-CTA kernel that pulls the
        activations out of pinned host memory into a private staging buffer and the GEMV kernel, which writes its
        outputs straight into the pinned host buffer - no copy-engine transfers, no output staging; the GEMV's weight
        stream runs while the activations cross PCIe.  (They are staged because every CTA reads all of them: from the
        device they come out of L2, from host memory each read would cross PCIe.)  All argument checks happen here,
        once.  Synchronize the stream before reading `out`; refill `input` in place between launches - from the HOST,
        after that synchronize: the staging kernel fetches `input` as soon as it starts, possibly while earlier work of
        the stream is still running, so `input` must not be written by device work of the same stream.
        Weight-on-the-right kernel, packed weight, no bias."""
        import ctypes

        from . import _native

        if self.kernel != "linear_y_f16RM_x_f16RM_W_any4TC" or not self.weight_reshaped or self.bias is not None:
            raise RuntimeError("bind_host needs a packed weight-on-the-right any4 layer without bias")
        dt = self.scales_and_zeros.dtype
        if input.is_cuda or not input.is_pinned() or input.dtype != dt or not input.is_contiguous():
            raise RuntimeError(f"bind_host: input must be a contiguous pinned CPU tensor of dtype {dt}")
        x2d = input.view(-1, input.shape[-1])
        m, k = x2d.shape
        w = self.weight
        w_rows = w.shape[0] * 8
        if k != self.in_features:
            raise RuntimeError("bind_host: wrong in_features")
        if out is None:
            out = torch.empty((m, w_rows), dtype=dt).pin_memory()
        if out.is_cuda or not out.is_pinned() or out.dtype != dt or tuple(out.shape) != (m, w_rows) or not out.is_contiguous():
            raise RuntimeError(f"bind_host: out must be a contiguous pinned CPU tensor [{m}][{w_rows}] of dtype {dt}")
        fmt = 2 if self.lut.dim() == 2 else 1  # tg_w4_format: any4 row-wise / global
        fn = _native.capi().tg_gemm_w4_rm_hostio
        dev = w.device
        xd = torch.empty((m, k), device=dev, dtype=dt)
        args = (ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(x2d.data_ptr()), ctypes.c_void_p(xd.data_ptr()),
                ctypes.c_void_p(w.data_ptr()), ctypes.c_void_p(self.scales_and_zeros.data_ptr()),
                ctypes.c_void_p(self.lut.data_ptr()), None,
                m, w_rows, k, self.group_size, w.shape[3] * 2, fmt, 1, 0 if dt == torch.bfloat16 else 1)
        keep = (input, out, xd, w, self.scales_and_zeros, self.lut)  # the bound buffers live as long as the callable
        stream_of = torch.cuda.current_stream

        def launch(_keep=keep):
            if torch.cuda.current_device() != dev.index:
                raise RuntimeError("bind_host: make the weight's device current before launching")
            if fn(*args, ctypes.c_void_p(stream_of().cuda_stream)) != 0:
                raise RuntimeError(_native.last_error())

        return launch, out

    def forward_host(self, input, out=None):
        """One-shot form of `bind_host`: launch and return `out` (pinned CPU; synchronize before reading it)."""
        with torch.cuda.device(self.weight.device):
            launch, out = self.bind_host(input, out)
            launch()
        return out

    def extra_repr(self):
        return super().extra_repr() + f", per_row={self.per_row}"


# The reference leaves these as a TODO (modules.py:10 "add FP4Linear, NF4Linear, MX4Linear").  NF4 and FP4 are any4 with
# ONE fixed 16-entry table for all rows (value = table[code] * scale + zero, scale = the group's absmax / table max,
# zero = 0): same kernels, `lut` is a buffer instead of a learnt parameter.  MX4 (OCP microscaling: fp4 e2m1 codes, one
# shared power-of-two exponent per 32 weights) runs through the mx4 op.
NF4_TABLE = (-1.0, -0.6961928009986877, -0.5250730514526367, -0.39491748809814453, -0.28444138169288635,
             -0.18477343022823334, -0.09105003625154495, 0.0, 0.07958029955625534, 0.16093020141124725,
             0.24611230194568634, 0.33791524171829224, 0.44070982933044434, 0.5626170039176941, 0.7229568362236023, 1.0)
FP4_TABLE = (0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0, -0.0, -0.5, -1.0, -1.5, -2.0, -3.0, -4.0, -6.0)  # e2m1, sign = bit 3


class _FixedTableLinear(Any4Linear):
    """any4 with a fixed global table.  `quantize_weight(w)` fills codes / scales from a float weight (absmax scaling
    per group, nearest table entry) and packs them."""

    TABLE = None

    def __init__(self, in_features, out_features, bias=True, device=None, dtype=None, group_size=128,
                 kernel="linear_y_f16RM_x_f16RM_W_any4TC", w_inner_k=4):
        super().__init__(in_features, out_features, bias, device, dtype, group_size, kernel, w_inner_k, per_row=False)

    def _init_extra(self, device, dtype):
        # a Parameter, as in Any4Linear (the kernels read `self.lut`; state_dicts stay loadable into Any4Linear)
        self.lut = torch.nn.Parameter(torch.tensor(self.TABLE, device=device, dtype=dtype), requires_grad=False)

    @torch.no_grad()
    def quantize_weight(self, w, w_inner_k=None):
        n, k, g = self.out_features, self.in_features, self.group_size
        if tuple(w.shape) != (n, k):
            raise ValueError(f"expected a [{n}][{k}] weight")
        dt = self.scales_and_zeros.dtype
        table = torch.tensor(self.TABLE, device=w.device, dtype=torch.float32)
        wg = w.float().view(n, k // g, g)
        scale = (wg.abs().amax(-1) / table.abs().max()).clamp_min(1e-8).to(dt)            # [n][k/g], rounded as stored
        codes = ((wg / scale.float().unsqueeze(-1)).unsqueeze(-1) - table).abs().argmin(-1)  # nearest table entry
        self.weight = torch.nn.Parameter(codes.view(n, k).to(torch.int32).to(self.weight.device), requires_grad=False)
        sz = torch.stack([scale.t(), torch.zeros_like(scale.t())], 2).contiguous()           # [k/g][n][2]
        self.scales_and_zeros.data = sz.to(self.scales_and_zeros.device)
        self.weight_reshaped = False
        self.reshape_weight(w_inner_k)
        return self


class NF4Linear(_FixedTableLinear):
    """QLoRA NormalFloat4 (the table of kmeans.py:17 in the reference) through the any4 kernels."""
    TABLE = NF4_TABLE


class FP4Linear(_FixedTableLinear):
    """fp4 e2m1 values with a real-valued group scale through the any4 kernels."""
    TABLE = FP4_TABLE


class MX4Linear(_PackedCheckpoint, torch.nn.Module):
    """OCP MX fp4 (e2m1 codes + one e8m0 exponent per 32 weights, tinygemm_lib/utils.py:137-232) through
    `tinygemm_y_f16RM_x_f16RM_w_mx4TC`; bf16 activations only, as the reference op.  Parameters: `weight` int32 codes
    ([out][in], or the packed B int4 layout after `reshape_weight`), `exponents` uint8 [out][in / 32], `bias`."""

    GROUP = 32

    def __init__(self, in_features, out_features, bias=True, device=None, dtype=torch.bfloat16, w_inner_k=4):
        super().__init__()
        if dtype not in (None, torch.bfloat16):
            raise ValueError("MX4Linear: the mx4 kernels take bfloat16 activations only")
        if in_features % self.GROUP:
            raise ValueError("MX4Linear: in_features must be a multiple of 32")
        self.in_features, self.out_features, self.group_size = in_features, out_features, self.GROUP
        self.weight = torch.nn.Parameter(torch.zeros((out_features, in_features), device=device, dtype=torch.int32),
                                         requires_grad=False)
        self.exponents = torch.nn.Parameter(torch.full((out_features, in_features // self.GROUP), 127, device=device,
                                                       dtype=torch.uint8), requires_grad=False)
        if bias:
            self.bias = torch.nn.Parameter(torch.empty(out_features, device=device, dtype=torch.bfloat16))
        else:
            self.register_parameter("bias", None)
        self.kernel = "tinygemm_y_f16RM_x_f16RM_w_mx4TC"
        self.w_inner_k = w_inner_k
        self.weight_reshaped = False

    def reshape_weight(self, w_inner_k=None):
        w_inner_k = self.w_inner_k if w_inner_k is None else w_inner_k
        self.weight.data = _ops.convert_matrix_to_m16n8k16_Bint4_layout(self.weight, w_inner_k)
        self.weight_reshaped, self.w_inner_k = True, w_inner_k

    @torch.no_grad()
    def quantize_weight(self, w, w_inner_k=None):
        from . import utils as U

        if tuple(w.shape) != (self.out_features, self.in_features):
            raise ValueError(f"expected a [{self.out_features}][{self.in_features}] weight")
        codes, exps = U.quantize_mx4(w.to(torch.bfloat16), self.GROUP)
        self.weight = torch.nn.Parameter(codes.to(torch.int32).to(self.weight.device), requires_grad=False)
        self.exponents = torch.nn.Parameter(exps.to(torch.uint8).to(self.exponents.device), requires_grad=False)
        self.weight_reshaped = False
        self.reshape_weight(w_inner_k)
        return self

    def forward(self, input):
        if not self.weight_reshaped:
            raise RuntimeError("MX4Linear: pack the weight first (reshape_weight / quantize_weight)")
        lead = input.shape[:-1]
        y = _ops.tinygemm_y_f16RM_x_f16RM_w_mx4TC(input.view(-1, input.shape[-1]), self.weight, self.GROUP,
                                                   self.exponents, True)
        y = y[:, : self.out_features]
        if self.bias is not None:
            y = y + self.bias
        return y.reshape(*lead, self.out_features)

    def extra_repr(self):
        return f"in_features={self.in_features}, out_features={self.out_features}, bias={self.bias is not None}, group_size=32"


def fuse_rows(linears, interleave=False):
    """Concatenate packed quantized Linears that read the same input along the output-feature axis (q|k|v, gate|up):
    ONE GEMV launch instead of several.  The row-tiled packed layouts concatenate along
    dim 0; every output element is computed from the same products as by the separate layers (rows are independent):
    bit-identical to torch.cat of the separate outputs whenever the launches use the same k-split, else equal up to the
    order of the fp32 partial sums (tests/test_decode_gpu.py::test_fuse_rows).  All layers must be packed (`weight_reshaped`), of the
    same class / kernel / group size / inner-k / dtype, with out_features a multiple of the tile height."""
    first = linears[0]
    if interleave:
        return _fuse_rows_interleaved(linears)
    tile = 8 if first._PACKERS.get(first.kernel, "").startswith("B") else 16
    for lin in linears:
        if type(lin) is not type(first) or not lin.weight_reshaped:
            raise ValueError("fuse_rows needs packed layers of one class")
        if (lin.in_features, lin.group_size, lin.kernel, lin.w_inner_k) != (
                first.in_features, first.group_size, first.kernel, first.w_inner_k):
            raise ValueError("fuse_rows: in_features / group_size / kernel / w_inner_k differ")
        if lin.out_features % tile or (lin.bias is None) != (first.bias is None):
            raise ValueError(f"fuse_rows: out_features must be multiples of {tile}, bias all-or-none")
        if getattr(lin, "per_row", True) is not True:
            raise ValueError("fuse_rows: a global LUT cannot be concatenated (use per_row=True)")
    dev, dt = first.weight.device, first.scales_and_zeros.dtype
    kw = dict(bias=first.bias is not None, device="meta", dtype=dt, group_size=first.group_size, kernel=first.kernel,
              w_inner_k=first.w_inner_k)
    fused = type(first)(first.in_features, sum(l.out_features for l in linears), **kw)
    fused.weight = torch.nn.Parameter(torch.cat([l.weight.data for l in linears], 0), requires_grad=False)
    fused.scales_and_zeros = torch.nn.Parameter(torch.cat([l.scales_and_zeros.data for l in linears], 1).contiguous())
    if hasattr(first, "lut"):
        fused.lut = torch.nn.Parameter(torch.cat([l.lut.data for l in linears], 0))
    if first.bias is not None:
        fused.bias = torch.nn.Parameter(torch.cat([l.bias.data for l in linears], 0))
    fused.weight_reshaped = True
    assert fused.weight.device == dev
    return fused


def _fuse_rows_interleaved(linears):
    """fuse_rows(..., interleave=True): rows alternate between the layers (row 2j = linears[0] row j, row 2j+1 =
    linears[1] row j, ...), the form any4_b200.decode.linear_silu_pairs consumes for gate|up.  Row order inside the
    packed tiles changes, so this works on UNPACKED layers (weight = [out][in] int32 codes) and packs the result."""
    first, r = linears[0], len(linears)
    for lin in linears:
        if type(lin) is not type(first) or lin.weight_reshaped:
            raise ValueError("interleaved fuse_rows needs unpacked layers of one class (call it before reshape_weight)")
        if (lin.in_features, lin.out_features, lin.group_size, lin.kernel, lin.w_inner_k) != (
                first.in_features, first.out_features, first.group_size, first.kernel, first.w_inner_k):
            raise ValueError("interleaved fuse_rows: shapes / group_size / kernel / w_inner_k differ")
        if lin.bias is not None or getattr(lin, "per_row", True) is not True:
            raise ValueError("interleaved fuse_rows: no bias, per-row LUT only")
    n, dt = first.out_features, first.scales_and_zeros.dtype
    kw = dict(bias=False, device="meta", dtype=dt, group_size=first.group_size, kernel=first.kernel,
              w_inner_k=first.w_inner_k)
    fused = type(first)(first.in_features, r * n, **kw)
    fused.weight = torch.nn.Parameter(
        torch.stack([l.weight.data for l in linears], 1).reshape(r * n, first.in_features), requires_grad=False)
    fused.scales_and_zeros = torch.nn.Parameter(
        torch.stack([l.scales_and_zeros.data for l in linears], 2).reshape(-1, r * n, 2).contiguous())
    if hasattr(first, "lut"):
        fused.lut = torch.nn.Parameter(torch.stack([l.lut.data for l in linears], 1).reshape(r * n, -1).contiguous())
    fused.reshape_weight(first.w_inner_k)
    return fused


class _SymmWorkspace:
    """Per (process group, device): the symmetric-memory exchange buffers of tg_gemm_w4_rm_exchange (a ring of `SLOTS`
    buffers of 8-byte tagged words, peer-mapped so every rank can store into every other rank's copy) and a ring of
    `SLOTS` plain local m x n output buffers.  Shared by all RowShardedLinear layers of the process.

    Contract: the tensor a fused sharded forward returns is a VIEW of output slot `call % SLOTS`; it stays valid until
    SLOTS - 1 further fused sharded calls (of any layer) have been issued on this rank.  Consume it (or copy it) before
    that.  Every rank must issue the same sequence of fused sharded calls."""

    SLOTS = 8
    _cache = {}

    def __init__(self, group, device, dtype, m_cap, n_cap):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        group = group if group is not None else dist.group.WORLD
        n_cap += n_cap & 1
        self.m_cap, self.n_cap = m_cap, n_cap
        self.xchg = symm.empty((self.SLOTS, m_cap, n_cap // 2), dtype=torch.int64, device=device)
        self.xchg.zero_()  # tag 0 is never used by a call
        self.hdl = symm.rendezvous(self.xchg, group)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.out = torch.empty((self.SLOTS, m_cap, n_cap), dtype=dtype, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group)  # every rank's exchange buffer is zero before anybody stores into it
        self.calls = 0

    @classmethod
    def get(cls, group, device, dtype, m, n):
        key = (id(group), str(device), dtype)
        ws = cls._cache.get(key)
        if ws is None or ws.m_cap < m or ws.n_cap < n:
            ws = cls(group, device, dtype, max(m, 16 if ws is None else ws.m_cap), max(n, 0 if ws is None else ws.n_cap))
            cls._cache[key] = ws
        return ws


class RowShardedLinear(torch.nn.Module):
    """Row-wise (output-feature) shard of a packed quantized Linear across the ranks of one node
    (BASELINE.json north_star: "weights shard row-wise across the 8 GPUs of one box
    with a single NCCL allreduce on the m x n output").

    Rank r keeps weight rows [r*n/R, (r+1)*n/R) - a contiguous slice of dim 0 of the packed weight
    (row tiles never straddle a shard: n/R must be a multiple of the tile height), the matching
    columns of `scales_and_zeros`, rows of `lut` and entries of `bias`.  forward() computes the
    local n/R outputs into its slice of a zero-filled m x n buffer and sums the buffers with ONE
    all-reduce: every element is value + 0 + ... + 0, so the result is bit-identical to the
    single-GPU output.

    `fused=True` (B200-native path, weight-on-the-right 4-bit kernels): no collective and no barrier kernel at all.  The
    GEMV's epilogue stores this rank's n/R outputs straight into EVERY rank's copy of a symmetric-memory m x n buffer
    over NVLink and completes the exchange itself (include/tinygemm_b200.h: tg_gemm_w4_rm_exchange: a counter in
    symmetric memory, the kernel's last CTA waits until all ranks' shards have landed).  Bit-identical to the
    all-reduce path (same values, no sums).  The result is a view of a ring slot of the shared workspace - see
    _SymmWorkspace for how long it stays valid.  `max_features` sizes the workspace (largest out_features of any
    sharded layer in the process).
    """

    def __init__(self, full: _PackedLinear, rank: int, world: int, group=None, fused: bool = False,
                 max_features: int = 0):
        super().__init__()
        self.fused = bool(fused) and world > 1 and full.kernel in (
            "linear_y_f16RM_x_f16RM_W_any4TC", "linear_y_f16RM_x_f16RM_W_int4TC")
        self.max_features = max(max_features, full.out_features)
        if not full.weight_reshaped:
            raise ValueError("pack the weight (reshape_weight) before sharding it")
        n = full.out_features
        tile = 16 if "_W_" in full.kernel and full.kernel.index("_W_") < full.kernel.index("_x_") else 8
        if n % (world * tile) != 0:
            raise ValueError(f"out_features={n} cannot be split into {world} shards of whole {tile}-row tiles")
        self.rank, self.world, self.group = rank, world, group
        self.out_features = n
        self.lo, self.hi = rank * n // world, (rank + 1) * n // world
        local = type(full).__new__(type(full))
        torch.nn.Module.__init__(local)
        for attr in ("in_features", "group_size", "kernel", "w_inner_k", "per_row", "n_bit"):
            if hasattr(full, attr):
                setattr(local, attr, getattr(full, attr))
        local.out_features = self.hi - self.lo
        local.weight_reshaped = True
        tl, th = self.lo // tile, self.hi // tile
        local.weight = torch.nn.Parameter(full.weight.data[tl:th].contiguous(), requires_grad=False)
        local.scales_and_zeros = torch.nn.Parameter(full.scales_and_zeros.data[:, self.lo:self.hi].contiguous(),
                                                    requires_grad=False)
        if hasattr(full, "lut"):
            lut = full.lut.data
            local.lut = torch.nn.Parameter((lut[self.lo:self.hi] if lut.dim() == 2 else lut).contiguous(),
                                           requires_grad=False)
        local.register_parameter("bias", None)  # bias is added once, after the reduction
        self.local = local
        self.bias = None if full.bias is None else torch.nn.Parameter(full.bias.data.clone(), requires_grad=False)

    def forward(self, input):
        import torch.distributed as dist

        lead = input.shape[:-1]
        x2d = input.view(-1, input.shape[-1])
        if self.fused:
            y = self._forward_fused(x2d)
            if self.bias is not None:
                y = y + self.bias
            return y.reshape(*lead, self.out_features)
        y_local = self.local._gemm(x2d)
        full = torch.zeros((x2d.shape[0], self.out_features), device=x2d.device, dtype=x2d.dtype)
        full[:, self.lo:self.hi] = y_local[:, : self.hi - self.lo]
        if self.world > 1:
            dist.all_reduce(full, op=dist.ReduceOp.SUM, group=self.group)
        if self.bias is not None:
            full = full + self.bias
        return full.view(*lead, self.out_features)

    def forward_silu_pairs(self, input):
        """For a layer whose weight rows alternate gate / up projections (fuse_rows(..., interleave=True)), sharded with
        `fused=True`: silu(gate) * up AND the exchange both happen in the GEMV kernel
        (include/tinygemm_b200.h: tg_gemm_w4_rm_exchange_silu_pairs); returns [..., out_features / 2]."""
        if not self.fused or self.bias is not None:
            raise RuntimeError("forward_silu_pairs needs fused=True and no bias")
        if (self.hi - self.lo) % 4:
            raise RuntimeError("forward_silu_pairs: the shard must hold whole (gate, up) row pairs in pairs of two")
        lead = input.shape[:-1]
        y = self._forward_fused(input.view(-1, input.shape[-1]), silu_pairs=True)
        return y.reshape(*lead, self.out_features // 2)

    def _forward_fused(self, x2d, silu_pairs=False):
        import ctypes

        from . import _native

        loc = self.local
        m, n = x2d.shape[0], self.out_features // (2 if silu_pairs else 1)
        if x2d.device != loc.weight.device:
            raise ValueError("RowShardedLinear: input and weights live on different devices")
        ws = _SymmWorkspace.get(self.group, x2d.device, x2d.dtype, m, self.max_features)
        slot = ws.calls % ws.SLOTS
        ws.calls += 1
        tag = ((ws.calls - 1) % 0xFFFFFFFF) + 1            # non-zero, unique per call
        out = ws.out[slot]
        xoff = slot * ws.m_cap * (ws.n_cap // 2) * 8       # this call's exchange buffer: the start of ring slot `slot`
        peers = (ctypes.c_void_p * self.world)(*[p + xoff for p in ws.ptrs])
        is_any4 = hasattr(loc, "lut")
        fmt = (2 if loc.lut.dim() == 2 else 1) if is_any4 else 0  # tg_w4_format
        lib = _native.capi()
        with torch.cuda.device(x2d.device):
            # the exchange completes INSIDE the kernel (tagged words into every rank's buffer, every CTA collects its
            # slice before it exits): no barrier / collective launch follows, consumers are ordered by stream order
            entry = lib.tg_gemm_w4_rm_exchange_silu_pairs if silu_pairs else lib.tg_gemm_w4_rm_exchange
            rc = entry(
                ctypes.c_void_p(out.data_ptr()), peers, self.rank, tag, self.world, ws.n_cap,
                ctypes.c_void_p(x2d.data_ptr()), ctypes.c_void_p(loc.weight.data_ptr()),
                ctypes.c_void_p(loc.scales_and_zeros.data_ptr()),
                ctypes.c_void_p(loc.lut.data_ptr()) if is_any4 else None, None,
                m, self.hi - self.lo, x2d.shape[1], loc.group_size, loc.weight.shape[3] * 2, fmt,
                0 if x2d.dtype == torch.bfloat16 else 1, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc != 0:
            raise RuntimeError(_native.last_error())
        return out[:m, :n]
