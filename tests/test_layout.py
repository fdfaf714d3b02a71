"""The oracle is test infrastructure: the product package must never import or call it."""

import os
import re

from helpers import ROOT

PKG = os.path.join(ROOT, "aind_exaspim_neuron_segmentation_b200")


def _sources(top, exts):
    for base, _, files in os.walk(top):
        if "__pycache__" in base:
            continue
        for f in files:
            if f.endswith(exts):
                yield os.path.join(base, f)


def test_product_does_not_touch_oracle_or_reference():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|/root/reference|oracle/_ref", re.M)
    for path in _sources(PKG, (".py", ".cu", ".cuh", ".h")):
        text = open(path).read()
        assert not pat.search(text), f"{path} refers to the oracle / reference tree"


def test_no_compat_layers_on_product_path():
    banned = re.compile(r"^\s*(import|from)\s+(triton|tilelang)\b|torch\.compile\(|cudnn|cublas", re.M | re.I)
    for path in _sources(PKG, (".py", ".cu", ".cuh", ".h")):
        text = open(path).read()
        assert not banned.search(text), f"{path} uses a banned library / compat layer"


def test_symlinked_package_name_exists():
    assert os.path.isdir(os.path.join(ROOT, "aind-exaspim-neuron-segmentation_b200", "csrc"))
