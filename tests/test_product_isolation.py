"""The product never imports, links or executes anything under oracle/ (or the reference)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'autoencoder_based_image_compression_b200')


def product_sources():
    for (folder, _, names) in os.walk(PKG):
        if os.path.basename(folder) in ('build', '__pycache__'):
            continue
        for name in names:
            if name.endswith(('.py', '.cu', '.cuh', '.h', '.cpp')):
                yield os.path.join(folder, name)
    yield os.path.join(ROOT, 'include', 'eae_b200.h')


def test_product_does_not_touch_the_oracle_or_the_reference():
    pattern = re.compile(r'(^\s*(from|import)\s+oracle\b)|liboracle|libref_coder|oracle/|/root/reference', re.M)
    offenders = []
    for path in product_sources():
        with open(path) as f:
            if pattern.search(f.read()):
                offenders.append(path)
    assert offenders == []


def test_oracle_says_it_is_test_infrastructure():
    for name in ('coder_oracle.c', 'ref_shim.cpp', 'coder.py', 'transforms.py', 'glue.py'):
        with open(os.path.join(ROOT, 'oracle', name)) as f:
            assert 'TEST INFRASTRUCTURE ONLY' in f.read(2000), name
