"""Drop-in for the reference's `src.utils.TensorflowWrapper` (src/utils/TensorflowWrapper.py:6-156): the fluent
`TensorWrapper` over graph tensors plus `placeholder / constant / zeros / ones / random_mask / parameter_guard`.

The reference builds a TensorFlow-1 graph and evaluates it with `tf.Session().run`. Here the wrapper builds a small
lazy expression graph of its own and `Session().run(fetches, feed_dict)` evaluates it on the B200: `matmul` (and the
`matmul -> add -> sigmoid` chain the encoder is made of, SDAV.py:129) runs as ONE fused tcgen05 kernel through libdlc;
the remaining element-wise / shape ops are not on the hot path and run as plain device tensor ops.
`to_tf()` returns the backend graph node, which is what reference call sites pass back into the wrapper.
"""
import numpy as np

float64 = np.float64
float32 = np.float32
int32 = np.int32
int64 = np.int64


class Node:
    """One vertex of the lazy graph. `static_rank` mirrors len(tensor.get_shape()) in the reference."""
    __slots__ = ("op", "inputs", "attrs", "static_rank", "dtype")

    def __init__(self, op, inputs=(), attrs=None, static_rank=None, dtype=float64):
        self.op = op
        self.inputs = tuple(inputs)
        self.attrs = attrs or {}
        self.static_rank = static_rank
        self.dtype = dtype

    # arithmetic on raw nodes, as tf.Tensor supports it (reference code mixes wrappers and raw tensors)
    def __matmul__(self, other):
        return Node("matmul", (self, _as_node(other)), static_rank=self.static_rank, dtype=self.dtype)

    def __add__(self, other):
        return _binary("add", self, other)

    def __radd__(self, other):
        return _binary("add", other, self)

    def __sub__(self, other):
        return _binary("sub", self, other)

    def __mul__(self, other):
        return _binary("mul", self, other)

    def __rmul__(self, other):
        return _binary("mul", other, self)

    def __truediv__(self, other):
        return _binary("div", self, other)

    def __getitem__(self, item):
        return Node("getitem", (self,), {"item": item}, static_rank=_rank_after_getitem(self.static_rank, item),
                    dtype=self.dtype)


def _rank_after_getitem(rank, item):
    if rank is None:
        return None
    if isinstance(item, (int, np.integer)):
        return max(rank - 1, 0)
    return rank


def _as_node(x):
    if isinstance(x, TensorWrapper):
        return x.x
    if isinstance(x, Node):
        return x
    arr = np.asarray(x)
    return Node("const", attrs={"value": arr}, static_rank=arr.ndim, dtype=arr.dtype.type)


def _binary(op, a, b):
    a, b = _as_node(a), _as_node(b)
    ranks = [r for r in (a.static_rank, b.static_rank) if r is not None]
    return Node(op, (a, b), static_rank=max(ranks) if ranks else None, dtype=a.dtype)


class TensorWrapper:
    def __init__(self, x):
        self.x = x.to_tf() if isinstance(x, TensorWrapper) else _as_node(x)

    # ---- shape helpers (TensorflowWrapper.py:13-32, 40-47)
    def flat_batch(self):
        shape = self.shape()
        return self.reshape([shape[0] * shape[1], shape[2]])

    def batch(self, batch_size):
        batch_size = TensorWrapper(batch_size)
        shape = self.shape()
        return self.reshape([batch_size, (shape[0] / batch_size).to(int32), shape[1]])

    def batch_size(self):
        if self.dimensions() == 3:
            return self.shape()[0]
        return 1

    def parameter_number(self):
        if self.dimensions() == 3:
            return self.shape()[1] * self.shape()[2]
        return self.shape()[0] * self.shape()[1]

    def corrupt(self, corruption_level):
        shape = self.shape()
        shape = shape[1:] if self.dimensions() == 3 else shape
        return self.multiply(random_mask(shape, corruption_level))

    def rank(self):
        return TensorWrapper(Node("rank", (self.x,), static_rank=0, dtype=int32))

    def dimensions(self):
        if self.x.static_rank is None:
            raise ValueError("tensor rank is not statically known")
        return self.x.static_rank

    def shape(self):
        return TensorWrapper(Node("shape", (self.x,), static_rank=1, dtype=int32))

    # ---- ops (TensorflowWrapper.py:49-87)
    def concat(self, y, axis=0):
        y = _as_node(parameter_guard(y))
        return TensorWrapper(Node("concat", (self.x, y), {"axis": axis}, static_rank=self.x.static_rank, dtype=self.x.dtype))

    def reshape(self, shape):
        shape = parameter_guard(shape)
        if isinstance(shape, Node):
            rank, inputs, spec = None, (self.x, shape), None
        else:
            rank = len(shape)
            spec = [None if isinstance(s, Node) else int(s) for s in shape]
            inputs = (self.x,) + tuple(s for s in shape if isinstance(s, Node))
        return TensorWrapper(Node("reshape", inputs, {"spec": spec}, static_rank=rank, dtype=self.x.dtype))

    def matmul(self, y):
        y = _as_node(parameter_guard(y))
        if self.dimensions() == TensorWrapper(y).dimensions():
            return TensorWrapper(self.x @ y)
        # 3-D x 2-D: flatten the batch, multiply, re-batch (TensorflowWrapper.py:64-67)
        batch_size = self.batch_size()
        return self.flat_batch().matmul(y).batch(batch_size)

    def add(self, y):
        return TensorWrapper(self.x + parameter_guard(y))

    def multiply(self, y):
        return TensorWrapper(_binary("mul", self.x, parameter_guard(y)))

    def sigmoid(self):
        return TensorWrapper(Node("sigmoid", (self.x,), static_rank=self.x.static_rank, dtype=self.x.dtype))

    def shuffle(self):
        return TensorWrapper(Node("shuffle", (self.x,), static_rank=self.x.static_rank, dtype=self.x.dtype))

    def round(self):
        return TensorWrapper(Node("round", (self.x,), static_rank=self.x.static_rank, dtype=self.x.dtype))

    def to(self, dtype):
        return TensorWrapper(Node("cast", (self.x,), {"dtype": dtype}, static_rank=self.x.static_rank, dtype=dtype))

    def __truediv__(self, y):
        return TensorWrapper(self.x / parameter_guard(y))

    def __getitem__(self, item):
        return TensorWrapper(self.x[parameter_guard(item)])

    def __mul__(self, y):
        return TensorWrapper(self.x * parameter_guard(y))

    def __rmul__(self, y):
        return TensorWrapper(parameter_guard(y) * self.x)

    def __add__(self, y):
        return TensorWrapper(self.x + parameter_guard(y))

    def __sub__(self, y):
        return TensorWrapper(self.x - parameter_guard(y))

    def to_tf(self):
        return self.x


def parameter_guard(y):
    if isinstance(y, TensorWrapper):
        return y.to_tf()
    if isinstance(y, list):
        return list(map(parameter_guard, y))
    return y


def placeholder(dtype, shape):
    return TensorWrapper(Node("placeholder", attrs={"shape": shape}, static_rank=None if shape is None else len(shape),
                              dtype=dtype))


def constant(value, shape=None, dtype=float64):
    arr = np.asarray(value, dtype=dtype)
    if shape:  # tf.constant(value, shape=...) fills / reshapes
        shape = [int(s) for s in parameter_guard(shape)]
        arr = np.broadcast_to(arr, shape).copy() if arr.size == 1 else arr.reshape(shape)
    return TensorWrapper(Node("const", attrs={"value": arr}, static_rank=arr.ndim, dtype=dtype))


def _fill(shape, value, dtype):
    shape = parameter_guard(shape)
    if isinstance(shape, Node):
        return TensorWrapper(Node("fill", (shape,), {"value": value}, static_rank=1, dtype=dtype))
    if isinstance(shape, (list, tuple)) and not any(isinstance(s, Node) for s in shape):
        return TensorWrapper(Node("const", attrs={"value": np.full([int(s) for s in shape], value, dtype=dtype)},
                                  static_rank=len(shape), dtype=dtype))
    nodes = [_as_node(s) for s in shape]
    return TensorWrapper(Node("fill", tuple(nodes), {"value": value, "list": True}, static_rank=len(nodes), dtype=dtype))


def zeros(shape, dtype=float64):
    return _fill(shape, 0, dtype)


def ones(shape, dtype=float64):
    return _fill(shape, 1, dtype)


def random_mask(shape, zeros_percentage, dtype=float64):
    """TensorflowWrapper.py:148-156: round(n * p) zeros, the rest ones, shuffled, reshaped."""
    shape = TensorWrapper(shape)
    zeros_percentage = TensorWrapper(zeros_percentage)
    parameters = shape[0] * shape[1]
    n_zeros = (parameters.to(float64) * zeros_percentage).round().to(int32)
    n_ones = parameters - n_zeros
    return ones(n_ones, dtype=dtype).concat(zeros(n_zeros, dtype=dtype)).shuffle().reshape(shape)


# --------------------------------------------------------------------------------------------------------------
# Evaluation on the device
# --------------------------------------------------------------------------------------------------------------
class Session:
    """Minimal stand-in for tf.Session: `run(fetches, feed_dict=None)` evaluates graph nodes on cuda:current and
    returns NumPy arrays. Context-manager protocol kept so reference-style `with Session() as sess:` reads the same."""

    def __init__(self, precision="fp16x2", seed=None):
        self.precision = precision
        self.seed = seed

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def run(self, fetches, feed_dict=None):
        import torch

        from . import _cuda
        _cuda.require_cuda()
        gen = torch.Generator(device="cuda")
        if self.seed is not None:
            gen.manual_seed(int(self.seed))
        ev = _Evaluator(feed_dict or {}, self.precision, gen)
        single = not isinstance(fetches, (list, tuple))
        outs = [ev.numpy(_as_node(f)) for f in ([fetches] if single else fetches)]
        return outs[0] if single else outs


_TORCH_DT = None


def _torch_dtype(dt):
    import torch
    global _TORCH_DT
    if _TORCH_DT is None:
        _TORCH_DT = {np.float64: torch.float64, np.float32: torch.float32, np.int32: torch.int32, np.int64: torch.int64,
                     float: torch.float64, int: torch.int64}
    return _TORCH_DT.get(dt, torch.float64)


class _Evaluator:
    def __init__(self, feed, precision, gen):
        self.feed = {(_as_node(k) if not isinstance(k, Node) else k): v for k, v in feed.items()}
        self.precision = precision
        self.gen = gen
        self.memo = {}

    def numpy(self, node):
        return self.value(node).cpu().numpy()

    def value(self, node):
        key = id(node)
        if key not in self.memo:
            self.memo[key] = self._eval(node)
        return self.memo[key]

    def _matmul(self, a, b, bias=None, act="none"):
        """[m,k] @ [k,n] on the tensor cores (float64 in/out like the reference graph; ~22-bit operands, fp32
        accumulate). Exact for small integers - the reference's only test vector."""
        import torch

        from . import ops
        if a.dim() != 2 or b.dim() != 2:
            raise ValueError("matmul expects rank-2 operands at evaluation time")
        out = ops.matmul(a.to(torch.float64).contiguous(), b.to(torch.float64).contiguous(),
                         None if bias is None else bias.to(torch.float64), act, self.precision)
        return out.to(torch.float64)

    def _eval(self, n):
        import torch
        op = n.op
        if op == "placeholder":
            if n not in self.feed:
                raise KeyError("placeholder was not fed")
            return torch.as_tensor(np.asarray(self.feed[n]), dtype=_torch_dtype(n.dtype)).cuda()
        if op == "const":
            return torch.as_tensor(n.attrs["value"]).cuda()
        if op == "sigmoid":
            src = n.inputs[0]
            # fused pattern: sigmoid(add(matmul(x, W), b)) -> one kernel (the encoder layer)
            if src.op == "add" and src.inputs[0].op == "matmul":
                mm = src.inputs[0]
                bias = self.value(src.inputs[1])
                if bias.dim() == 1:
                    return self._matmul(self.value(mm.inputs[0]), self.value(mm.inputs[1]), bias, "sigmoid")
            return torch.sigmoid(self.value(src))
        if op == "matmul":
            return self._matmul(self.value(n.inputs[0]), self.value(n.inputs[1]))
        if op in ("add", "sub", "mul", "div"):
            a, b = self.value(n.inputs[0]), self.value(n.inputs[1])
            if op == "add":
                return a + b
            if op == "sub":
                return a - b
            if op == "mul":
                return a * b
            return a / b
        if op == "shape":
            return torch.tensor(list(self.value(n.inputs[0]).shape), dtype=torch.int32, device="cuda")
        if op == "rank":
            return torch.tensor(self.value(n.inputs[0]).dim(), dtype=torch.int32, device="cuda")
        if op == "getitem":
            return self.value(n.inputs[0])[n.attrs["item"]]
        if op == "cast":
            v = self.value(n.inputs[0])
            dt = _torch_dtype(n.attrs["dtype"])
            return torch.trunc(v).to(dt) if (v.is_floating_point() and not dt.is_floating_point) else v.to(dt)
        if op == "round":
            return torch.round(self.value(n.inputs[0]))  # half to even, like tf.round
        if op == "reshape":
            x = self.value(n.inputs[0])
            spec = n.attrs["spec"]
            if spec is None:
                shape = [int(v) for v in self.value(n.inputs[1]).tolist()]
            else:
                dyn = iter(n.inputs[1:])
                shape = [int(self.value(next(dyn)).item()) if s is None else s for s in spec]
            return x.reshape(shape)
        if op == "concat":
            a, b = self.value(n.inputs[0]), self.value(n.inputs[1])
            return torch.cat([a, b.to(a.dtype)], dim=n.attrs["axis"])
        if op == "fill":
            if n.attrs.get("list"):
                shape = [int(self.value(s).item()) for s in n.inputs]
            else:
                v = self.value(n.inputs[0])
                shape = [int(v.item())] if v.dim() == 0 else [int(t) for t in v.tolist()]
            return torch.full(shape, n.attrs["value"], dtype=_torch_dtype(n.dtype), device="cuda")
        if op == "shuffle":
            x = self.value(n.inputs[0])
            perm = torch.randperm(x.shape[0], device="cuda", generator=self.gen)
            return x[perm]
        raise NotImplementedError("TensorWrapper op %r" % op)
