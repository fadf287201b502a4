"""Pins the STRUCTURE of the transform oracle (and of the weight format) to a reference-held artefact.

The reference ships its serialised TensorFlow graphs (kodak_tensorflow/eae/results/*/model_*.ckpt.meta); the tensor data
is absent, the GraphDef is not. tests/golden/make_graph_fixture.py (run in the build container) parses the nine files
without TensorFlow into tests/golden/graph_structure.json. Here the inference chain found in every one of them is
compared, op by op, with the layer tables oracle/transforms.py executes (ENCODER_LAYERS / DECODER_LAYERS) and with the
variable names / shapes weights.py assumes: op order, which variable feeds which op, strides, padding, data format, the
operand order of the GDN matmul (x^2 @ gamma, no transposes) and Div (GDN) vs Mul (IGDN)
(eae/graph/components.py:11-142, tfutils/tfutils.py:363-397, 480-509, EntropyAutoencoder.py:108-224).
"""
import json
import os

import pytest

from autoencoder_based_image_compression_b200 import weights as wts
from oracle import transforms as T
from tests import util

with open(os.path.join(util.GOLDEN_DIR, 'graph_structure.json')) as _f:
    GRAPHS = json.load(_f)


def var_of(graph_input):
    """'encoder/weights_1/read' -> 'encoder/weights_1'."""
    return graph_input[:-len('/read')] if graph_input.endswith('/read') else graph_input


class Chain(object):
    def __init__(self, ops):
        self.ops = list(ops)
        self.at = 0

    def take(self, op_type):
        op = self.ops[self.at]
        assert op['op'] == op_type, 'op {} is a {}, the oracle expects a {}'.format(op['name'], op['op'], op_type)
        self.at += 1
        return op

    def take_norm(self, gamma, beta_scope, combine):
        """tfuls.gdn / inverse_gdn: Square -> MatMul(x^2, gamma) -> Add(beta) -> Sqrt -> Div | Mul (tfutils.py:393-397, 506-509)."""
        square = self.take('Square')
        matmul = self.take('MatMul')
        assert matmul['input'][0] == square['name'] and var_of(matmul['input'][1]) == gamma       # x^2 @ gamma[j, i]
        assert matmul['attr']['transpose_a'] is False and matmul['attr']['transpose_b'] is False
        add = self.take('Add')
        assert add['input'][0] == matmul['name']
        sqrt = self.take('Sqrt')
        assert sqrt['input'] == [add['name']]
        out = self.take(combine)
        assert out['input'][0] == square['input'][0] and out['input'][1] == sqrt['name']          # x (/ or *) sqrt(...)
        return out


@pytest.mark.parametrize('model', sorted(GRAPHS))
def test_oracle_layer_tables_equal_the_reference_graph(model):
    graph = GRAPHS[model]
    learned = model.startswith('learning_bw')
    chain = Chain(graph['ops'])
    # ---- components.encoder
    for layer in T.ENCODER_LAYERS:
        conv = chain.take('Conv2D')
        assert var_of(conv['input'][1]) == layer['conv']
        assert conv['attr']['strides'] == [1, layer['stride'], layer['stride'], 1]
        assert conv['attr']['padding'] == 'SAME' and conv['attr']['data_format'] == 'NHWC'
        bias = chain.take('BiasAdd')
        assert bias['input'][0] == conv['name'] and var_of(bias['input'][1]) == layer['bias']
        if not (layer['optional'] and learned):
            chain.take_norm(layer['gdn'][0], layer['gdn'][1], 'Div')
    # ---- components.decoder
    for layer in T.DECODER_LAYERS:
        if not (layer['optional'] and learned):
            chain.take_norm(layer['igdn'][0], layer['igdn'][1], 'Mul')
        tconv = chain.take('Conv2DBackpropInput')      # inputs: output_shape, filter, the tensor being up-sampled
        assert var_of(tconv['input'][1]) == layer['tconv']
        assert tconv['attr']['strides'] == [1, layer['stride'], layer['stride'], 1]
        assert tconv['attr']['padding'] == 'SAME' and tconv['attr']['data_format'] == 'NHWC'
        if layer['bias'] is not None:
            bias = chain.take('BiasAdd')
            assert bias['input'][0] == tconv['name'] and var_of(bias['input'][1]) == layer['bias']
    assert chain.at == len(chain.ops), 'the reference graph has ops the oracle does not run: {}'.format(
        [op['name'] for op in chain.ops[chain.at:]])
    assert [T.ENCODER_LAYERS[i]['stride'] for i in range(3)] == list(T.STRIDES)


@pytest.mark.parametrize('model', sorted(GRAPHS))
def test_weight_format_equals_the_reference_variables(model):
    variables = GRAPHS[model]['variables']
    learned = model.startswith('learning_bw')
    keys = [k for k in wts.ENCODER_KEYS + wts.DECODER_KEYS if not (learned and k in wts.OPTIONAL_KEYS)]
    graph_keys = sorted(k for k in variables if k.split('/')[0] in ('encoder', 'decoder'))
    assert graph_keys == sorted(keys)
    for key in keys:
        short = key.split('/')[1]
        expected = wts.SHAPES[short] if short.startswith('weights') else ((128, 128) if short.startswith('gamma') else (128,))
        assert tuple(variables[key]['shape']) == tuple(expected), key
        assert variables[key]['dtype'] == 1                                                       # DT_FLOAT
    assert tuple(variables[wts.BIN_WIDTHS_KEY]['shape']) == (128,)
    # what random_init draws is a complete, valid set for this architecture
    wts.validate(wts.random_init(0, learned), learned)
