"""Drop-in for the reference's quantized Linear modules (modules.py:12-230): `Int4Linear`,
`Int8Linear`, `Any4Linear` with the same constructor arguments, parameter names / shapes
(`weight` int32, `scales_and_zeros` [k/g][n][2], `lut`, `bias`), `kernel` strings, `w_inner_k`
semantics, `reshape_weight()` and `forward()`; state_dicts are interchangeable.

Added on top: `RowShardedLinear`, the row-wise multi-GPU wrapper with one
all-reduce on the m x n output.
"""
import torch

from . import functional as F

_ops = torch.ops.tinygemm


class _PackedLinear(torch.nn.Module):
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

    def extra_repr(self):
        return super().extra_repr() + f", per_row={self.per_row}"


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
    """

    def __init__(self, full: _PackedLinear, rank: int, world: int, group=None):
        super().__init__()
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
        y_local = self.local._gemm(x2d)
        full = torch.zeros((x2d.shape[0], self.out_features), device=x2d.device, dtype=x2d.dtype)
        full[:, self.lo:self.hi] = y_local[:, : self.hi - self.lo]
        if self.world > 1:
            dist.all_reduce(full, op=dist.ReduceOp.SUM, group=self.group)
        if self.bias is not None:
            full = full + self.bias
        return full.view(*lead, self.out_features)
