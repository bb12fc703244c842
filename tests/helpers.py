"""Shared test utilities: rebuild seeded inputs from a golden recipe and check them."""
import json

import numpy as np
import torch

from oracle import cases


def words_inputs(g):
    kw = json.loads(str(g["recipe"]))
    c = cases.words_case(**kw)
    got = np.array(cases.checksum(c["img"]) + cases.checksum(c["words"]))
    np.testing.assert_allclose(got, g["in_checksum"], rtol=1e-12, err_msg="seeded inputs drifted from the fixture")
    return kw, c


def gag_inputs(g):
    kw = json.loads(str(g["recipe"]))
    c = cases.gag_case(**kw)
    got = np.array(cases.checksum(c["x"]) + cases.checksum(c["key"]))
    np.testing.assert_allclose(got, g["in_checksum"], rtol=1e-12, err_msg="seeded inputs drifted from the fixture")
    return kw, c


def relmax(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def finite_close(a, b, atol):
    """allclose that also requires the -inf pattern to match exactly."""
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    ia, ib = torch.isinf(a), torch.isinf(b)
    assert torch.equal(ia, ib), "-inf pattern differs"
    d = (a[~ia].double() - b[~ib].double()).abs().max().item() if (~ia).any() else 0.0
    assert d <= atol, "max abs diff %g > %g" % (d, atol)
    return d
