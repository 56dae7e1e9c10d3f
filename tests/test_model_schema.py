"""Drop-in boundary checks that need no GPU: the state_dict schema equals the reference's (strict load), the C-ABI
library loads and exports every symbol include/roitr_b200.h declares, and the product path refuses to run on CPU."""
import ctypes
import os
import re

import pytest
import torch

from roitr_b200 import _lib, model
from tests.helpers import CONFIG_3D, CONFIG_4D, schema, weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("cfg,factor", [(CONFIG_3D, 1), (CONFIG_4D, 2)])
def test_state_dict_schema_matches_reference(cfg, factor):
    m = model.create_model(cfg)
    got = [(k, list(v.shape)) for k, v in m.state_dict().items()]
    assert got == [(k, list(s)) for k, s in schema(factor)]          # names, order and shapes
    m.load_state_dict(weights(factor), strict=True)


def test_config_attribute_or_item_access():
    class Attr:
        pass
    a = Attr()
    for k, v in CONFIG_3D.items():
        setattr(a, k, v)
    assert model.create_model(a).cfg == model.create_model(CONFIG_3D).cfg


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "roitr_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = re.findall(r"\b(roitr_\w+)\s*\(", hdr)
    assert len(names) >= 20
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in set(names):
        assert hasattr(L, n), "libroitr_b200.so does not export %s" % n
    assert L.roitr_abi_version() == 1


def test_no_cpu_fallback():
    m = model.create_model(CONFIG_3D).eval()
    x = torch.zeros(64, 3)
    with pytest.raises(RuntimeError):
        m(x, x, torch.ones(64, 1), torch.ones(64, 1), x, x, torch.eye(3), torch.zeros(3, 1), x)
    with pytest.raises(RuntimeError):
        model.create_model(CONFIG_3D).train()(x, x, x, x, x, x, x, x, x)
