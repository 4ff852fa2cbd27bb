"""A released MDX-Net ``.onnx`` as a device-resident torch module.

The reference builds ``model_run`` from the ``.onnx`` either through onnxruntime with a HOST round trip per chunk
(``spek.cpu().numpy()``, /root/reference/handlers/patch_separate.py:19-52) or, when the segment size differs from the model's,
through ``onnx2torch.convert`` on the torch device (:54-63).  Neither ``onnx`` nor ``onnxruntime`` nor ``onnx2torch`` exists in
this build, so this file reads the protobuf wire format itself (ModelProto / GraphProto / NodeProto / TensorProto /
AttributeProto field numbers of onnx.proto3) and runs the graph node by node with torch operators -- the equivalent of the
reference's onnx2torch branch: weights and activations stay on the device between al_stft and al_istft.

Operator coverage is what TFC-TDF exports use (Conv, ConvTranspose, BatchNormalization, Relu, MatMul, Gemm, Add / Sub / Mul /
Div, Transpose, Reshape, Flatten, Concat, Slice, Pad, Identity, Constant, Sigmoid, Tanh, Shape / Gather / Unsqueeze / Squeeze /
Cast on shape chains); anything else raises ``NotImplementedError`` naming the operator."""
from __future__ import annotations

import struct
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

# ---- protobuf wire format -------------------------------------------------------------------------------------------------


def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _fields(buf: bytes):
    """Yields (field number, wire type, value): varint -> int, 64-bit / 32-bit -> raw bytes, length-delimited -> bytes."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val, pos = buf[pos:pos + 8], pos + 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val, pos = buf[pos:pos + ln], pos + ln
        elif wt == 5:
            val, pos = buf[pos:pos + 4], pos + 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield fno, wt, val


def _sint64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _packed_varints(val, wt) -> List[int]:
    if wt == 0:
        return [_sint64(val)]
    out, pos = [], 0
    while pos < len(val):
        v, pos = _varint(val, pos)
        out.append(_sint64(v))
    return out


_DTYPES = {1: np.float32, 2: np.uint8, 3: np.int8, 6: np.int32, 7: np.int64, 9: np.bool_, 10: np.float16, 11: np.float64}


def _tensor(buf: bytes) -> Tuple[str, np.ndarray]:
    dims: List[int] = []
    dtype, name, raw = 1, "", None
    floats: List[float] = []
    ints: List[int] = []
    for fno, wt, val in _fields(buf):
        if fno == 1:
            dims += _packed_varints(val, wt)
        elif fno == 2:
            dtype = val
        elif fno == 4:                                   # float_data
            floats += list(struct.unpack(f"<{len(val) // 4}f", val)) if wt == 2 else [struct.unpack("<f", val)[0]]
        elif fno in (5, 7):                              # int32_data / int64_data
            ints += _packed_varints(val, wt)
        elif fno == 8:
            name = val.decode()
        elif fno == 9:
            raw = val
        elif fno == 13:
            raise NotImplementedError("ONNX tensors with external data are not supported")
    if dtype not in _DTYPES:
        raise NotImplementedError(f"ONNX tensor data type {dtype}")
    np_dt = _DTYPES[dtype]
    if raw is not None:
        arr = np.frombuffer(raw, dtype=np.dtype(np_dt).newbyteorder("<")).astype(np_dt)
    elif floats:
        arr = np.asarray(floats, dtype=np_dt)
    else:
        arr = np.asarray(ints, dtype=np_dt)
    return name, arr.reshape(dims) if dims else arr.reshape(())


def _attribute(buf: bytes):
    name, val = "", None
    ints: List[int] = []
    floats: List[float] = []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            name = v.decode()
        elif fno == 2:
            val = struct.unpack("<f", v)[0]
        elif fno == 3:
            val = _sint64(v)
        elif fno == 4:
            val = v.decode(errors="replace")
        elif fno == 5:
            val = _tensor(v)[1]
        elif fno == 7:
            floats += list(struct.unpack(f"<{len(v) // 4}f", v)) if wt == 2 else [struct.unpack("<f", v)[0]]
        elif fno == 8:
            ints += _packed_varints(v, wt)
    if val is None:
        val = ints if ints else (floats if floats else None)
    return name, val


class Node:
    __slots__ = ("op", "inputs", "outputs", "attrs", "name")

    def __init__(self, buf: bytes):
        self.inputs, self.outputs, self.attrs, self.op, self.name = [], [], {}, "", ""
        for fno, _, v in _fields(buf):
            if fno == 1:
                self.inputs.append(v.decode())
            elif fno == 2:
                self.outputs.append(v.decode())
            elif fno == 3:
                self.name = v.decode()
            elif fno == 4:
                self.op = v.decode()
            elif fno == 5:
                k, a = _attribute(v)
                self.attrs[k] = a


def _value_info(buf: bytes) -> Tuple[str, List[Optional[int]]]:
    """ValueInfoProto -> (name, static dims or None per axis)."""
    name, dims = "", []
    for fno, _, v in _fields(buf):
        if fno == 1:
            name = v.decode()
        elif fno == 2:                                    # TypeProto
            for f2, _, v2 in _fields(v):
                if f2 == 1:                               # tensor_type
                    for f3, _, v3 in _fields(v2):
                        if f3 == 2:                       # shape
                            for f4, _, v4 in _fields(v3):
                                if f4 == 1:               # dim
                                    d = None
                                    for f5, _, v5 in _fields(v4):
                                        if f5 == 1:
                                            d = _sint64(v5)
                                    dims.append(d)
    return name, dims


def parse_model(data: bytes):
    """-> (nodes, initializers {name: ndarray}, graph inputs [(name, dims)], graph output names)."""
    graph = None
    for fno, _, v in _fields(data):
        if fno == 7:
            graph = v
    if graph is None:
        raise ValueError("no graph in the ONNX file")
    nodes: List[Node] = []
    inits: Dict[str, np.ndarray] = {}
    inputs, outputs = [], []
    for fno, _, v in _fields(graph):
        if fno == 1:
            nodes.append(Node(v))
        elif fno == 5:
            n, a = _tensor(v)
            inits[n] = a
        elif fno == 11:
            inputs.append(_value_info(v))
        elif fno == 12:
            outputs.append(_value_info(v)[0])
    inputs = [(n, d) for n, d in inputs if n not in inits]
    return nodes, inits, inputs, outputs


# ---- execution ------------------------------------------------------------------------------------------------------------


def _pads(attrs, nd: int) -> List[int]:
    p = attrs.get("pads")
    if p is None:
        return [0] * nd
    if list(p[:nd]) != list(p[nd:]):
        raise NotImplementedError(f"asymmetric convolution padding {p}")
    return list(p[:nd])


class OnnxGraphNet(nn.Module):
    """``forward(spek [B, 4, dim_f, dim_t]) -> same shape``: the graph of the file, evaluated in order on the module's device."""

    def __init__(self, nodes, inits, inputs, outputs):
        super().__init__()
        if len(inputs) != 1 or len(outputs) != 1:
            raise NotImplementedError(f"expected one graph input and one output, got {len(inputs)} / {len(outputs)}")
        self.nodes = nodes
        self.input_name, self.input_dims = inputs[0]
        self.output_name = outputs[0]
        self._names: Dict[str, str] = {}
        for i, (k, a) in enumerate(inits.items()):
            key = f"t{i}"
            self._names[k] = key
            t = torch.from_numpy(np.ascontiguousarray(a))
            self.register_buffer(key, t.float() if t.dtype in (torch.float16, torch.float64) else t, persistent=False)
        unsupported = sorted({n.op for n in nodes if not hasattr(self, "_op_" + n.op)})
        if unsupported:
            raise NotImplementedError(f"ONNX operators outside the MDX-Net subset: {unsupported}")

    @classmethod
    def from_file(cls, path: str) -> "OnnxGraphNet":
        with open(path, "rb") as f:
            return cls(*parse_model(f.read()))

    @property
    def dim_f(self) -> Optional[int]:
        return self.input_dims[2] if len(self.input_dims) == 4 else None

    @property
    def dim_t(self) -> Optional[int]:
        return self.input_dims[3] if len(self.input_dims) == 4 else None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        env: Dict[str, torch.Tensor] = {self.input_name: x}
        get = lambda n: env[n] if n in env else getattr(self, self._names[n])
        for node in self.nodes:
            args = [get(n) if n else None for n in node.inputs]
            out = getattr(self, "_op_" + node.op)(node.attrs, *args)
            if node.op == "Constant" and out.is_floating_point():
                out = out.to(device=x.device, dtype=x.dtype)      # shape constants stay on the host, data constants follow x
            if isinstance(out, tuple):
                for n, o in zip(node.outputs, out):
                    env[n] = o
            else:
                env[node.outputs[0]] = out
        return env[self.output_name]

    # -- operators (ONNX semantics of the default opset range 9..17 for the attributes used) --
    @staticmethod
    def _op_Conv(a, x, w, b=None):
        nd = w.dim() - 2
        conv = {1: F.conv1d, 2: F.conv2d}[nd]
        return conv(x, w.to(x.dtype), None if b is None else b.to(x.dtype), stride=a.get("strides", [1] * nd), padding=_pads(a, nd),
                    dilation=a.get("dilations", [1] * nd), groups=a.get("group", 1))

    @staticmethod
    def _op_ConvTranspose(a, x, w, b=None):
        nd = w.dim() - 2
        conv = {1: F.conv_transpose1d, 2: F.conv_transpose2d}[nd]
        return conv(x, w.to(x.dtype), None if b is None else b.to(x.dtype), stride=a.get("strides", [1] * nd), padding=_pads(a, nd),
                    output_padding=a.get("output_padding", [0] * nd), dilation=a.get("dilations", [1] * nd), groups=a.get("group", 1))

    @staticmethod
    def _op_BatchNormalization(a, x, scale, bias, mean, var):
        return F.batch_norm(x, mean.to(x.dtype), var.to(x.dtype), scale.to(x.dtype), bias.to(x.dtype), False, 0.0, a.get("epsilon", 1e-5))

    @staticmethod
    def _op_Relu(a, x):
        return F.relu(x)

    @staticmethod
    def _op_Sigmoid(a, x):
        return torch.sigmoid(x)

    @staticmethod
    def _op_Tanh(a, x):
        return torch.tanh(x)

    @staticmethod
    def _op_MatMul(a, x, y):
        return torch.matmul(x, y.to(x.dtype))

    @staticmethod
    def _op_Gemm(a, x, w, c=None):
        x = x.transpose(0, 1) if a.get("transA", 0) else x
        w = w.transpose(0, 1) if a.get("transB", 0) else w
        y = a.get("alpha", 1.0) * torch.matmul(x, w.to(x.dtype))
        return y if c is None else y + a.get("beta", 1.0) * c.to(x.dtype)

    @staticmethod
    def _op_Add(a, x, y):
        return x + y

    @staticmethod
    def _op_Sub(a, x, y):
        return x - y

    @staticmethod
    def _op_Mul(a, x, y):
        return x * y

    @staticmethod
    def _op_Div(a, x, y):
        return torch.div(x, y, rounding_mode="trunc") if not (x.is_floating_point() or y.is_floating_point()) else x / y

    @staticmethod
    def _op_Transpose(a, x):
        return x.permute(a["perm"]) if a.get("perm") else x.permute(*reversed(range(x.dim())))

    @staticmethod
    def _op_Reshape(a, x, shape):
        tgt = [int(s) for s in shape.tolist()]
        tgt = [x.shape[i] if s == 0 else s for i, s in enumerate(tgt)]
        return x.reshape(tgt)

    @staticmethod
    def _op_Flatten(a, x):
        ax = a.get("axis", 1)
        return x.reshape(int(np.prod(x.shape[:ax])) if ax else 1, -1)

    @staticmethod
    def _op_Concat(a, *xs):
        return torch.cat(list(xs), dim=a["axis"])

    @staticmethod
    def _op_Identity(a, x):
        return x

    @staticmethod
    def _op_Constant(a):
        v = a.get("value")
        if v is None:
            raise NotImplementedError("Constant without a tensor value")
        return torch.from_numpy(np.ascontiguousarray(v))

    @staticmethod
    def _op_Shape(a, x):
        return torch.tensor(list(x.shape), dtype=torch.int64)

    @staticmethod
    def _op_Gather(a, x, idx):
        ax = a.get("axis", 0)
        idx = idx.to(torch.int64).to(x.device)
        return torch.index_select(x, ax, idx.reshape(-1)).reshape(x.shape[:ax] + tuple(idx.shape) + x.shape[ax + 1:])

    @staticmethod
    def _op_Unsqueeze(a, x, axes=None):
        ax = a.get("axes") if axes is None else [int(v) for v in axes.tolist()]
        for d in sorted(ax):
            x = x.unsqueeze(d)
        return x

    @staticmethod
    def _op_Squeeze(a, x, axes=None):
        ax = a.get("axes") if axes is None else [int(v) for v in axes.tolist()]
        if ax is None:
            return x.squeeze()
        for d in sorted(ax, reverse=True):
            x = x.squeeze(d)
        return x

    @staticmethod
    def _op_Cast(a, x):
        to = {1: torch.float32, 6: torch.int32, 7: torch.int64, 10: torch.float16, 11: torch.float64, 9: torch.bool}[a["to"]]
        return x.to(to)

    @staticmethod
    def _op_Slice(a, x, starts=None, ends=None, axes=None, steps=None):
        if starts is None:                                # opset < 10: attributes
            starts, ends, axes = a["starts"], a["ends"], a.get("axes")
        else:
            starts, ends = starts.tolist(), ends.tolist()
            axes = None if axes is None else axes.tolist()
            steps = None if steps is None else steps.tolist()
        axes = list(range(len(starts))) if axes is None else axes
        steps = [1] * len(starts) if steps is None else steps
        sl = [slice(None)] * x.dim()
        for s, e, ax, st in zip(starts, ends, axes, steps):
            if st <= 0:
                raise NotImplementedError("Slice with a non-positive step")
            n = x.shape[ax]
            s = max(s + n, 0) if s < 0 else min(s, n)
            e = max(e + n, 0) if e < 0 else min(e, n)
            sl[ax] = slice(int(s), int(e), int(st))
        return x[tuple(sl)]

    @staticmethod
    def _op_Pad(a, x, pads=None, value=None, axes=None):
        p = a.get("pads") if pads is None else [int(v) for v in pads.tolist()]
        mode = a.get("mode", "constant")
        nd = x.dim()
        flat = []
        for d in reversed(range(nd)):                     # torch wants the last axis first, (before, after) pairs
            flat += [p[d], p[d + nd]]
        v = 0.0 if value is None else float(value)
        return F.pad(x, flat, mode=mode, value=v) if mode == "constant" else F.pad(x, flat, mode=mode)
