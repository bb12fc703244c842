"""CPU: the C-ABI library loads and exports every symbol include/*.h declares, and the
ctypes table in eegan_b200/_lib.py names exactly that set.  No compute calls (no GPU)."""
import ctypes
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names |= set(re.findall(r"\b(eegan_[a-z0-9_]+)\s*\(", src))
    return names


def test_header_symbols_exported():
    from eegan_b200 import _lib
    assert os.path.isfile(_lib.LIB_PATH), "build the library first: make -C eegan_b200/csrc"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    decl = declared_symbols()
    assert len(decl) >= 20
    for name in sorted(decl):
        assert hasattr(handle, name), "symbol %s declared in include/ but not exported" % name
    assert decl == set(_lib.SIGNATURES), "ctypes table and header disagree: %s" % (decl ^ set(_lib.SIGNATURES))


def test_version_and_error_channel():
    from eegan_b200 import _lib
    L = _lib.lib()
    assert L.eegan_abi_version() == 1
    # argument validation happens before any CUDA call, so it is observable without a GPU
    rc = L.eegan_damsm_pair_fwd(None, None, None, 0, 0, 256, 289, 18, 5.0, 5.0, None, None, 0, None, 0, None)
    assert rc == 1 and b"empty batch" in L.eegan_last_error()
    rc = L.eegan_gag_fwd(None, None, None, None, 0, 2, 32, 64, 40, None, None, None)
    assert rc == 1 and b"T=40" in L.eegan_last_error()
    assert L.eegan_damsm_pair_workspace_bytes(48, 48, 256, 289, 18) > 3 * 48 * 48 * 18 * 289 * 4


def test_no_cpu_fallback():
    import torch
    import eegan_b200
    x = torch.randn(2, 8), torch.randn(2, 8)
    with pytest.raises(RuntimeError, match="no CPU path"):
        eegan_b200.sent_loss(x[0], x[1], torch.arange(2), None, 2)
    with pytest.raises(RuntimeError, match="no CPU path"):
        eegan_b200.words_loss(torch.randn(2, 8, 3, 3), torch.randn(2, 8, 4), torch.arange(2), torch.tensor([4, 3]), None, 2)
