"""nets/onnx_graph.py: the repository's own ONNX reader + graph executor (the reference turns the
.onnx into `model_run` through onnxruntime / onnx2torch, handlers/patch_separate.py:19-63).  There is no `onnx` package here, so
the test ENCODES a small TFC-TDF-shaped graph with a few lines of protobuf writing and compares the imported module with the
same arithmetic written directly in torch; the Separator then loads that file under an MDX model name."""
import struct

import numpy as np
import pytest
import torch
import torch.nn.functional as F


# ---- a minimal protobuf writer (test side only) ---------------------------------------------------------------------------
def _vi(n):
    n &= (1 << 64) - 1
    out = b""
    while True:
        b = n & 0x7F
        n >>= 7
        out += bytes([b | (0x80 if n else 0)])
        if not n:
            return out


def _ld(fno, payload):
    return _vi((fno << 3) | 2) + _vi(len(payload)) + payload


def _iv(fno, v):
    return _vi(fno << 3) + _vi(v)


def _tensor(name, arr, raw=True):
    arr = np.asarray(arr)
    dt = {np.dtype(np.float32): 1, np.dtype(np.int64): 7}[arr.dtype]
    out = b"".join(_iv(1, d) for d in arr.shape) + _iv(2, dt) + _ld(8, name.encode())
    if raw:
        out += _ld(9, arr.astype(arr.dtype.newbyteorder("<")).tobytes())
    elif dt == 1:
        out += _ld(4, struct.pack(f"<{arr.size}f", *arr.reshape(-1).tolist()))
    else:
        out += _ld(7, b"".join(_vi(int(v)) for v in arr.reshape(-1)))
    return out


def _attr(name, v):
    out = _ld(1, name.encode())
    if isinstance(v, float):
        return out + _vi((2 << 3) | 5) + struct.pack("<f", v) + _iv(20, 1)
    if isinstance(v, int):
        return out + _iv(3, v) + _iv(20, 2)
    if isinstance(v, np.ndarray):
        return out + _ld(5, _tensor("", v)) + _iv(20, 4)
    return out + _ld(8, b"".join(_vi(int(i)) for i in v)) + _iv(20, 7)       # packed ints


def _node(op, ins, outs, **attrs):
    return (b"".join(_ld(1, i.encode()) for i in ins) + b"".join(_ld(2, o.encode()) for o in outs) + _ld(4, op.encode())
            + b"".join(_ld(5, _attr(k, v)) for k, v in attrs.items()))


def _value_info(name, dims):
    shape = b"".join(_ld(1, _iv(1, d)) for d in dims)
    return _ld(1, name.encode()) + _ld(2, _ld(1, _iv(1, 1) + _ld(2, shape)))


def _model(nodes, inits, inp, out):
    g = b"".join(_ld(1, n) for n in nodes) + _ld(2, b"g") + b"".join(_ld(5, t) for t in inits) + _ld(11, _value_info(*inp)) \
        + _ld(12, _value_info(*out))
    return _iv(1, 7) + _ld(7, g)


def _tiny_tfc_tdf(dim_f=16, dim_t=8, c=6, seed=0):
    """first conv 1x1 + BN + ReLU -> transpose -> TDF (MatMul over F, Add bias, ReLU, MatMul, Add) -> residual Add -> transpose back
    -> strided Conv (down) -> ConvTranspose (up) -> Mul with the skip -> Concat with itself -> Slice half -> final conv 1x1, plus a
    Constant scale and a Reshape/Shape-free path: the operator set of the released graphs."""
    rs = np.random.RandomState(seed)
    f = lambda *s: (rs.standard_normal(s) * 0.3).astype(np.float32)
    W = dict(w0=f(c, 4, 1, 1), b0=f(c), g=np.abs(f(c)) + 0.5, be=f(c), mu=f(c), var=np.abs(f(c)) + 0.5,
             t1=f(dim_f, dim_f // 4), tb1=f(dim_f // 4), t2=f(dim_f // 4, dim_f), tb2=f(dim_f),
             wd=f(c, c, 2, 2), bd=f(c), wu=f(c, c, 2, 2), bu=f(c), w1=f(4, c, 1, 1), b1=f(4), k=np.asarray(0.5, np.float32))
    nodes = [
        _node("Conv", ["input", "w0", "b0"], ["c0"], kernel_shape=[1, 1], strides=[1, 1], pads=[0, 0, 0, 0], dilations=[1, 1], group=1),
        _node("BatchNormalization", ["c0", "g", "be", "mu", "var"], ["n0"], epsilon=1e-5),
        _node("Relu", ["n0"], ["r0"]),
        _node("Transpose", ["r0"], ["x"], perm=[0, 1, 3, 2]),
        _node("MatMul", ["x", "t1"], ["m1"]), _node("Add", ["m1", "tb1"], ["a1"]), _node("Relu", ["a1"], ["h1"]),
        _node("MatMul", ["h1", "t2"], ["m2"]), _node("Add", ["m2", "tb2"], ["a2"]), _node("Add", ["x", "a2"], ["y"]),
        _node("Transpose", ["y"], ["yt"], perm=[0, 1, 3, 2]),
        _node("Conv", ["yt", "wd", "bd"], ["d"], kernel_shape=[2, 2], strides=[2, 2], pads=[0, 0, 0, 0]),
        _node("ConvTranspose", ["d", "wu", "bu"], ["u"], kernel_shape=[2, 2], strides=[2, 2]),
        _node("Mul", ["u", "yt"], ["s"]),
        _node("Constant", [], ["kc"], value=W["k"]),
        _node("Mul", ["s", "kc"], ["sk"]),
        _node("Concat", ["sk", "yt"], ["cc"], axis=1),
        _node("Slice", ["cc", "sl_s", "sl_e", "sl_a"], ["half"]),
        _node("Conv", ["half", "w1", "b1"], ["output"], kernel_shape=[1, 1]),
    ]
    inits = [_tensor(k, v, raw=(i % 2 == 0)) for i, (k, v) in enumerate(W.items()) if k != "k"]
    inits += [_tensor("sl_s", np.asarray([0], np.int64), raw=False), _tensor("sl_e", np.asarray([c], np.int64)),
              _tensor("sl_a", np.asarray([1], np.int64))]
    data = _model(nodes, inits, ("input", [1, 4, dim_f, dim_t]), ("output", [1, 4, dim_f, dim_t]))

    def direct(x):
        t = {k: torch.from_numpy(v) for k, v in W.items()}
        h = F.relu(F.batch_norm(F.conv2d(x, t["w0"], t["b0"]), t["mu"], t["var"], t["g"], t["be"], False, 0.0, 1e-5))
        xx = h.transpose(-1, -2)
        y = xx + (F.relu(xx @ t["t1"] + t["tb1"]) @ t["t2"] + t["tb2"])
        yt = y.transpose(-1, -2)
        u = F.conv_transpose2d(F.conv2d(yt, t["wd"], t["bd"], stride=2), t["wu"], t["bu"], stride=2)
        return F.conv2d(torch.cat([u * yt * 0.5, yt], 1)[:, :c], t["w1"], t["b1"])

    return data, direct


def test_onnx_graph_import_matches_direct_torch(tmp_path):
    from audiolab_b200.nets.onnx_graph import OnnxGraphNet
    data, direct = _tiny_tfc_tdf()
    p = tmp_path / "tiny.onnx"
    p.write_bytes(data)
    net = OnnxGraphNet.from_file(str(p)).eval()
    assert (net.dim_f, net.dim_t) == (16, 8) and len(net.nodes) == 19
    x = torch.randn(3, 4, 16, 8, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        assert torch.allclose(net(x), direct(x), rtol=1e-5, atol=1e-6)


def test_onnx_graph_rejects_operators_outside_the_subset(tmp_path):
    from audiolab_b200.nets.onnx_graph import OnnxGraphNet, parse_model
    data = _model([_node("LSTM", ["input"], ["output"])], [], ("input", [1, 4, 8, 8]), ("output", [1, 4, 8, 8]))
    with pytest.raises(NotImplementedError, match="LSTM"):
        OnnxGraphNet(*parse_model(data))


@pytest.mark.gpu
def test_separator_runs_an_mdx_model_from_its_onnx_file(tmp_path):
    """Separator.load_model on a real file: the .onnx is read, dim_f / dim_t come from its input, model_run stays on the device,
    and the demix equals the oracle's with the same network."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import copy
    from audiolab_b200.nets.onnx_graph import OnnxGraphNet
    from audiolab_b200.separator import Separator
    from oracle import mdx as omdx
    from oracle.metrics import max_abs_err
    from oracle.synth import synth_mix
    data, _ = _tiny_tfc_tdf(dim_f=3072, dim_t=16, c=4, seed=3)
    (tmp_path / "UVR-MDX-NET-Voc_FT.onnx").write_bytes(data)
    sep = Separator(log_level=40, model_file_dir=str(tmp_path), use_autocast=False, mdx_params={"segment_size": 256})
    inst = sep.load_model("UVR-MDX-NET-Voc_FT.onnx")
    assert inst.dim_t == 16 and isinstance(inst.demixer, object)
    mix = synth_mix(30001, seed=4)
    got = sep.separate_tensor(torch.tensor(mix))["Vocals"].cpu()
    net = OnnxGraphNet.from_file(str(tmp_path / "UVR-MDX-NET-Voc_FT.onnx")).eval()
    ocfg = omdx.MdxConfig(n_fft=6144, dim_f=3072, dim_t_log2=4, overlap=0.25, zero_low_bins=3)
    ref = omdx.demix_windowed(mix, copy.deepcopy(net), ocfg)
    # the random graph amplifies (|output| ~ 1e3): the 1e-4 of unit-scale stems, relative to the peak
    assert max_abs_err(got, ref) <= 1e-4 * max(1.0, float(np.abs(ref).max()))
