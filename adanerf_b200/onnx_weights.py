"""Minimal ONNX initialiser reader (no `onnx` package needed).

Reads the fp32 initialisers of the reference's exported `model{0,1}.onnx`
(writer: /root/reference/src/export.py:81-83, torch.onnx.export of BaseNet / NeRF) by walking the
protobuf wire format directly: ModelProto.graph(7) -> GraphProto.initializer(5) -> TensorProto
{dims(1), data_type(2), float_data(4), name(8), raw_data(9)}.  The C++ twin is
adanerf_b200/csrc/onnx_reader.cpp; both return tensors keyed by the state_dict names
(`layers.0.weight`, `pts_linears.3.bias`, ...) which the exporter preserves 1:1.
"""
import struct

import numpy as np


def _varint(buf, pos):
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not (b & 0x80):
            return result, pos
        shift += 7


def _fields(buf, start, end):
    """Yield (field_number, wire_type, value_or_(start,end)) for one message."""
    pos = start
    while pos < end:
        key, pos = _varint(buf, pos)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
            yield fn, wt, v
        elif wt == 1:
            yield fn, wt, (pos, pos + 8)
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            yield fn, wt, (pos, pos + ln)
            pos += ln
        elif wt == 5:
            yield fn, wt, (pos, pos + 4)
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")


def _tensor(buf, start, end):
    dims, name, raw, dtype, floats = [], None, None, None, []
    for fn, wt, v in _fields(buf, start, end):
        if fn == 1:
            if wt == 0:
                dims.append(v)
            else:  # packed
                p, e = v
                while p < e:
                    d, p = _varint(buf, p)
                    dims.append(d)
        elif fn == 2:
            dtype = v
        elif fn == 4:
            if wt == 5:
                floats.append(struct.unpack_from("<f", buf, v[0])[0])
            else:
                floats.extend(np.frombuffer(buf, dtype="<f4", count=(v[1] - v[0]) // 4, offset=v[0]).tolist())
        elif fn == 8:
            name = bytes(buf[v[0]:v[1]]).decode("utf-8")
        elif fn == 9:
            raw = (v[0], v[1])
    if dtype != 1:  # FLOAT
        return name, None
    if raw is not None:
        arr = np.frombuffer(buf, dtype="<f4", count=(raw[1] - raw[0]) // 4, offset=raw[0]).copy()
    else:
        arr = np.asarray(floats, dtype=np.float32)
    return name, arr.reshape(dims) if dims else arr


def read_onnx_initializers(path):
    """-> dict name -> float32 ndarray (shape as stored: weights [out,in], biases [out])."""
    with open(path, "rb") as f:
        buf = memoryview(f.read())
    out = {}
    for fn, wt, v in _fields(buf, 0, len(buf)):
        if fn == 7 and wt == 2:  # graph
            for gfn, gwt, gv in _fields(buf, v[0], v[1]):
                if gfn == 5 and gwt == 2:
                    name, arr = _tensor(buf, gv[0], gv[1])
                    if arr is not None:
                        out[name] = arr
    return out
