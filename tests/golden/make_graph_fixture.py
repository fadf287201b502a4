"""Generates tests/golden/graph_structure.json from the reference's own checkpoint metadata.

    python tests/golden/make_graph_fixture.py        (build container only: reads /root/reference)

The reference ships no tensor data (.MISSING_LARGE_BLOBS) but it does ship the serialised TensorFlow graphs
``kodak_tensorflow/eae/results/*/model_*.ckpt.meta`` (MetaGraphDef protobufs). This script walks them with a
minimal pure-Python protobuf reader (no TensorFlow, no protobuf package) and writes, per model, the
inference-relevant structure: the variables (name, shape, dtype) and, in graph order, every Conv2D /
Conv2DBackpropInput / BiasAdd / MatMul / Sqrt / RealDiv / Mul node of the encoder and the decoder with its inputs and
its attributes (strides, padding, data_format, transpose flags, static output_shape of the transposed convolutions).
tests/test_graph_structure.py asserts that oracle/transforms.py and weights.py assume exactly that structure
(eae/graph/components.py:11-142, EntropyAutoencoder.py:108-224, tfutils.py:363-397, 480-509).
"""
import glob
import json
import os
import re
import struct
import sys

REF = '/root/reference/kodak_tensorflow/eae/results'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'graph_structure.json')


def varint(buf, at):
    (value, shift) = (0, 0)
    while True:
        b = buf[at]
        at += 1
        value |= (b & 0x7F) << shift
        if not (b & 0x80):
            return (value, at)
        shift += 7


def fields(buf):
    """Yields (field number, wire type, value) of one message; length-delimited values are memoryviews."""
    at = 0
    n = len(buf)
    while at < n:
        (key, at) = varint(buf, at)
        (num, wt) = (key >> 3, key & 7)
        if wt == 0:
            (v, at) = varint(buf, at)
        elif wt == 1:
            v = bytes(buf[at:at + 8]); at += 8
        elif wt == 2:
            (ln, at) = varint(buf, at)
            v = buf[at:at + ln]; at += ln
        elif wt == 5:
            v = bytes(buf[at:at + 4]); at += 4
        else:
            raise ValueError('unsupported wire type {}'.format(wt))
        yield (num, wt, v)


def signed(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def packed_ints(buf):
    (out, at) = ([], 0)
    while at < len(buf):
        (v, at) = varint(buf, at)
        out.append(signed(v))
    return out


def shape_of(buf):
    """TensorShapeProto: repeated Dim dim = 2 {int64 size = 1}; bool unknown_rank = 3."""
    dims = []
    for (num, wt, v) in fields(buf):
        if num == 2:
            size = 0
            for (n2, _, v2) in fields(v):
                if n2 == 1:
                    size = signed(v2)
            dims.append(size)
    return dims


def tensor_ints(buf):
    """TensorProto holding a small int32 vector (the `output_shape` constant of conv2d_transpose)."""
    (dtype, content, vals) = (None, None, [])
    for (num, wt, v) in fields(buf):
        if num == 1:
            dtype = v
        elif num == 4:
            content = bytes(v)
        elif num == 7:                      # int_val
            vals += packed_ints(v) if wt == 2 else [signed(v)]
    if content is not None and dtype == 3:  # DT_INT32
        return list(struct.unpack('<{}i'.format(len(content)//4), content))
    return vals


def attr_value(buf):
    """AttrValue: list = 1, s = 2, i = 3, f = 4, b = 5, type = 6, shape = 7, tensor = 8."""
    for (num, wt, v) in fields(buf):
        if num == 2:
            return bytes(v).decode('latin1')
        if num == 3:
            return signed(v)
        if num == 4:
            return struct.unpack('<f', v)[0]
        if num == 5:
            return bool(v)
        if num == 6:
            return {'dtype': v}
        if num == 7:
            return {'shape': shape_of(v)}
        if num == 8:
            return {'tensor': tensor_ints(v)}
        if num == 1:
            out = []
            for (n2, w2, v2) in fields(v):
                if n2 == 3:
                    out += packed_ints(v2) if w2 == 2 else [signed(v2)]
                elif n2 == 2:
                    out.append(bytes(v2).decode('latin1'))
            return out
    return None


def node_of(buf):
    node = {'name': '', 'op': '', 'input': [], 'attr': {}}
    for (num, wt, v) in fields(buf):
        if num == 1:
            node['name'] = bytes(v).decode()
        elif num == 2:
            node['op'] = bytes(v).decode()
        elif num == 3:
            node['input'].append(bytes(v).decode())
        elif num == 5:
            (key, value) = (None, None)
            for (n2, _, v2) in fields(v):
                if n2 == 1:
                    key = bytes(v2).decode()
                elif n2 == 2:
                    value = attr_value(v2)
            node['attr'][key] = value
    return node


def graph_nodes(meta_bytes):
    buf = memoryview(meta_bytes)
    for (num, wt, v) in fields(buf):
        if num == 2:                        # MetaGraphDef.graph_def
            for (n2, _, v2) in fields(v):
                if n2 == 1:                 # GraphDef.node
                    yield node_of(v2)


KEEP_OPS = ('Conv2D', 'Conv2DBackpropInput', 'BiasAdd', 'MatMul', 'Sqrt', 'RealDiv', 'Div', 'Mul', 'Square', 'Reshape',
            'Add')
KEEP_ATTRS = ('strides', 'padding', 'data_format', 'transpose_a', 'transpose_b', 'use_cudnn_on_gpu')


def structure(path):
    nodes = list(graph_nodes(open(path, 'rb').read()))
    by_name = {n['name']: n for n in nodes}
    variables = {}
    for n in nodes:
        if n['op'] in ('VariableV2', 'Variable') and n['name'].split('/')[0] in ('encoder', 'decoder', 'piecewise_linear_function'):
            if '/Adam' in n['name'] or n['name'].endswith(('_power', '/Momentum')):
                continue
            variables[n['name']] = {'shape': n['attr'].get('shape', {}).get('shape'),
                                    'dtype': n['attr'].get('dtype', {}).get('dtype')}
    placeholders = {n['name']: n['attr'].get('shape', {}).get('shape') for n in nodes if n['op'] == 'Placeholder'}
    # The inference chain: TensorFlow names the ops of components.encoder / components.decoder and of tfuls.gdn /
    # tfuls.inverse_gdn after their Python calls, outside any variable scope: Conv2D[_k], BiasAdd[_k], Square[_k],
    # MatMul[_k], Add[_k], Sqrt[_k], Div[_k] (GDN) / Mul[_k] (IGDN), conv2d_transpose[_k]. Gradients, the optimiser,
    # the piecewise-linear density and operator overloads (lower-case add_k / mul_k / div_k) are left out.
    chain = re.compile(r'^(Conv2D|BiasAdd|Square|MatMul|Add|Sqrt|Div|Mul|conv2d_transpose)(_\d+)?$')
    ops = []
    for n in nodes:
        if not chain.match(n['name']) or n['op'] not in KEEP_OPS:
            continue
        entry = {'name': n['name'], 'op': n['op'], 'input': n['input'],
                 'attr': {k: v for (k, v) in n['attr'].items() if k in KEEP_ATTRS}}
        if n['op'] == 'Conv2DBackpropInput':      # input 0 is the output_shape tensor
            const = by_name.get(n['input'][0].split(':')[0])
            if const is not None and const['op'] == 'Const':
                entry['output_shape'] = const['attr'].get('value', {}).get('tensor')
        ops.append(entry)
    return {'variables': variables, 'placeholders': placeholders, 'ops': ops, 'nb_nodes': len(nodes)}


def main():
    if not os.path.isdir(REF):
        sys.exit('reference tree absent: this generator runs in the build container only')
    out = {}
    for path in sorted(glob.glob(os.path.join(REF, '*', 'model_*.ckpt.meta'))):
        model = os.path.basename(os.path.dirname(path))
        out[model] = structure(path)
        out[model]['source'] = os.path.relpath(path, '/root/reference')
    with open(OUT, 'w') as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print('wrote {} ({} models)'.format(OUT, len(out)))


if __name__ == '__main__':
    main()
