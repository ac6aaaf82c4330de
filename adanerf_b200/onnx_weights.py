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
        if pos >= len(buf) or shift > 63:
            raise ValueError("truncated or corrupt protobuf varint")
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
            if pos + 8 > end:
                raise ValueError("truncated protobuf message")
            yield fn, wt, (pos, pos + 8)
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            if pos + ln > end:
                raise ValueError("truncated protobuf message")
            yield fn, wt, (pos, pos + ln)
            pos += ln
        elif wt == 5:
            if pos + 4 > end:
                raise ValueError("truncated protobuf message")
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


# ---------------------------------------------------------------------------------------------
# Writer (tests / tooling): emits a ModelProto whose graph carries only fp32 initialisers with raw_data,
# i.e. the part of the reference's export (src/export.py:81-83) that the renderer consumes.
def _enc_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _field(fn, wt, payload):
    key = _enc_varint((fn << 3) | wt)
    if wt == 0:
        return key + _enc_varint(payload)
    return key + _enc_varint(len(payload)) + payload


def write_onnx_initializers(path, tensors):
    """tensors: dict name -> float32 ndarray.  Layout: ModelProto{ir_version(1)=4, graph(7)=GraphProto{initializer(5)*}}."""
    graph = b""
    for name, arr in tensors.items():
        a = np.ascontiguousarray(arr, dtype="<f4")
        t = b"".join(_field(1, 0, int(d)) for d in a.shape)
        t += _field(2, 0, 1)                         # data_type = FLOAT
        t += _field(8, 2, name.encode("utf-8"))
        t += _field(9, 2, a.tobytes())               # raw_data, little endian
        graph += _field(5, 2, t)
    model = _field(1, 0, 4) + _field(7, 2, graph)
    with open(path, "wb") as f:
        f.write(model)


def write_export_dir(path, scene, sd0, sd1, thr, K):
    """Writes {config.ini, dataset_info.txt, model0.onnx, model1.onnx} in the reference's export format
    (src/export.py:28-93, src/train_data.py:180-195)."""
    import os
    os.makedirs(path, exist_ok=True)
    as_np = lambda sd: {k: (v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)) for k, v in sd.items()}
    write_onnx_initializers(os.path.join(path, "model0.onnx"), as_np(sd0))
    write_onnx_initializers(os.path.join(path, "model1.onnx"), as_np(sd1))
    ndc = bool(scene.get("use_ndc"))
    with open(os.path.join(path, "dataset_info.txt"), "w") as f:
        f.write(f"view_cell_center = {list(scene['view_cell_center'])}\n")
        f.write(f"view_cell_size = {list(scene['view_cell_size'])}\n")
        f.write(f"depth_range = {list(scene['depth_range'])}\n")
        f.write(f"fov = {scene['fov']}\nfocal = 0.0\ncamera_scale = 1.0\nmax_depth = {scene['max_depth']}\n")
        if ndc and scene.get("w") and scene.get("h"):   # not written by src/export.py; read by our loader when present
            f.write(f"w = {int(scene['w'])}\nh = {int(scene['h'])}\n")
    with open(os.path.join(path, "config.ini"), "w") as f:
        if ndc:   # configs/fine_training_ndc.ini
            f.write("posEnc = [nerf, nerf]\nposEncArgs = [2-2, 10-4]\ninFeatures = [SpherePosDir, RayMarchFromPoses]\n"
                    "outFeatures = [RawSigmoid, RGBARayMarch]\nrayMarchSampler = [none, FromClassifiedDepthAdaptiveNoDepthRange]\n"
                    "rayMarchNormalization = [InverseSqrtDistCentered, None]\nuseNDC = True\n"
                    f"numRaymarchSamples = [{K}, {K}]\ndepthTransform = linear\nzNear = [0.001, 0.001]\nzFar = [1.0, 1.0]\n"
                    f"adaptiveSamplingThreshold = {thr}\nmultiDepthFeatures = [128, 128]\naccumulationMult = alpha\n")
        else:
            f.write("posEnc = [nerf, nerf]\nposEncArgs = [10-4, 10-4]\ninFeatures = [SpherePosDir, RayMarchFromPoses]\n"
                    "outFeatures = [RawSigmoid, RGBARayMarch]\nrayMarchSampler = [none, FromClassifiedDepthAdaptive]\n"
                    "rayMarchNormalization = [InverseSqrtDistCentered, InverseSqrtDistCentered]\n"
                    f"numRaymarchSamples = [{K}, {K}]\ndepthTransform = log\nzNear = [0.001, 0.001]\nzFar = [1.0, 1.0]\n"
                    f"adaptiveSamplingThreshold = {thr}\nmultiDepthFeatures = [128, 128]\naccumulationMult = alpha\n")
