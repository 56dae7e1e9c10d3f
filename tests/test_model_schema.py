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


def test_geo_table_interpolation_bound_cpu():
    """Host logic of the tabulated geometric embedding: the table step chosen by engine.build_geo_tables keeps the 4-point
    Lagrange interpolation within its bound of the fp64 truth (emulating the kernel's fp32 arithmetic in numpy)."""
    import numpy as np
    import torch
    from roitr_b200 import engine
    from tests.helpers import weights
    sd = weights(1)
    e = "backbone.global_transformer.embedding"
    Wd, bd, Wa, ba, dv = [sd[e + k] for k in (".proj_d.weight", ".proj_d.bias", ".proj_a.weight", ".proj_a.bias",
                                               ".embedding.div_term")]
    T = engine.build_geo_tables(Wd, bd, Wa, ba, dv)
    assert T["bound"] <= engine.GEO_TABLE_TOL and T["inv_h"] == 2.0 ** round(np.log2(T["inv_h"]))
    rng = np.random.default_rng(0)
    for tab, Wm, b, tmax in ((T["tab_d"], Wd, bd, 200.0), (T["tab_a"], Wa, ba, 12.0)):
        G, R, _ = tab.shape
        full = tab.permute(1, 0, 2).reshape(R, -1).numpy()
        t = rng.uniform(0, tmax, 4000).astype(np.float32)
        x = t * np.float32(T["inv_h"]); fl = np.floor(x); u = (x - fl).astype(np.float32); i = fl.astype(np.int64)
        w = [np.float32(-1 / 6) * u * (u - 1) * (u - 2), np.float32(.5) * (u + 1) * (u - 1) * (u - 2),
             np.float32(-.5) * (u + 1) * u * (u - 2), np.float32(1 / 6) * (u + 1) * u * (u - 1)]
        got = sum(w[k][:, None] * full[i + k] for k in range(4))
        om = torch.tensor(t, dtype=torch.float64)[:, None] * dv.double()[None, :]
        emb = torch.stack([torch.sin(om), torch.cos(om)], 2).reshape(len(t), -1)
        truth = (emb @ Wm.double().t() + b.double()).numpy()
        assert np.abs(got - truth).max() < 1e-6
