"""CPU: host-side pieces of the drop-in boundary (no kernels)."""
import sys
import types

import pytest
import torch

import eegan_b200
from eegan_b200 import config, damsm_losses


def test_install_routes_reference_import_names():
    saved = {k: sys.modules.get(k) for k in ("miscc", "miscc.DAMSM_losses", "sync_batchnorm")}
    try:
        sys.modules["miscc"] = types.ModuleType("miscc")
        eegan_b200.install()
        from miscc.DAMSM_losses import sent_loss, words_loss  # train.py:24
        from sync_batchnorm import DataParallelWithCallback, SynchronizedBatchNorm2d  # models.py:8, train.py:25
        assert words_loss is damsm_losses.words_loss and sent_loss is damsm_losses.sent_loss
        assert SynchronizedBatchNorm2d.__module__.startswith("eegan_b200")
        assert DataParallelWithCallback.__module__.startswith("eegan_b200")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_public_signatures_match_reference():
    import inspect
    expect = {
        "cosine_similarity": ["x1", "x2", "dim", "eps"],
        "func_attention": ["query", "context", "gamma1"],
        "sent_similarity": ["cnn_code", "rnn_code", "class_ids", "batch_size", "eps"],
        "words_similarity": ["img_features", "words_emb", "cap_lens", "class_ids", "batch_size"],
        "sent_loss": ["cnn_code", "rnn_code", "labels", "class_ids", "batch_size", "eps"],
        "words_loss": ["img_features", "words_emb", "labels", "cap_lens", "class_ids", "batch_size"],
    }
    for name, params in expect.items():
        assert list(inspect.signature(getattr(damsm_losses, name)).parameters) == params
    gag = damsm_losses.GlobalAttentionGeneral(32, 256)
    assert list(inspect.signature(gag.forward).parameters) == ["input", "context_key", "content_value"]
    assert hasattr(gag, "applyMask") and gag.mask is None


def test_cfg_follows_reference_object_when_present():
    saved = sys.modules.get("miscc.config")
    try:
        sys.modules.pop("miscc.config", None)
        assert config.gammas() == (5.0, 5.0, 10.0)  # miscc/config.py:47-51
        fake = types.ModuleType("miscc.config")
        fake.cfg = types.SimpleNamespace(TRAIN=types.SimpleNamespace(
            SMOOTH=types.SimpleNamespace(GAMMA1=4.0, GAMMA2=3.0, GAMMA3=2.0)))
        sys.modules["miscc.config"] = fake
        assert config.gammas() == (4.0, 3.0, 2.0)
    finally:
        if saved is None:
            sys.modules.pop("miscc.config", None)
        else:
            sys.modules["miscc.config"] = saved


def test_att_maps_have_reference_shapes_and_type():
    att = torch.arange(3 * 4 * 9, dtype=torch.float32).reshape(3, 4, 9)
    lens = torch.tensor([4, 2, 3])
    # host caption lengths: the reference's exact type (a list), no laziness needed
    maps = damsm_losses._att_maps(att, lens, (3, 3))
    assert type(maps) is list and len(maps) == 3
    assert [tuple(m.shape) for m in maps] == [(1, 4, 3, 3), (1, 2, 3, 3), (1, 3, 3, 3)]
    assert torch.equal(maps[1].reshape(2, 9), att[1, :2])
    # device caption lengths -> lazy Sequence: never an empty list to C-level consumers
    lazy = damsm_losses._LazyAttMaps(att, lens, (3, 3))
    assert not isinstance(lazy, list) and len(lazy) == 3
    assert torch.equal(torch.cat(list(lazy), 1), torch.cat(maps, 1))
    lazy2 = damsm_losses._LazyAttMaps(att, lens, (3, 3))
    assert type(lazy2 + []) is list and len(lazy2 + []) == 3 and type(lazy2.tolist()) is list
    import pickle
    back = pickle.loads(pickle.dumps(damsm_losses._LazyAttMaps(att, lens, (3, 3))))
    assert type(back) is list and torch.equal(back[2], maps[2])


def test_labels_none_short_circuits_like_reference():
    # DAMSM_losses.py:265-269: labels=None -> (None, None) — decided before any device work
    # for sent_loss the inputs are validated first, so CPU tensors still raise
    with pytest.raises(RuntimeError):
        damsm_losses.sent_loss(torch.zeros(2, 4), torch.zeros(2, 4), None, None, 2)


def test_device_ids_helper():
    t = damsm_losses._device_i64([3, 1, 2, 9], "cpu", 3)
    assert t.dtype == torch.int64 and t.tolist() == [3, 1, 2]
    assert damsm_losses._device_i64(None, "cpu") is None


def test_shard_mode_selection():
    """The grid partition of the N > 1 path: package default (images), explicit choice, and a loud error for anything else."""
    from eegan_b200 import sharded
    assert sharded.SHARD_BY in ("images", "captions")
    assert sharded._shard_mode(None) == sharded.SHARD_BY
    assert sharded._shard_mode("captions") == "captions" and sharded._shard_mode("images") == "images"
    import pytest
    with pytest.raises(ValueError):
        sharded._shard_mode("rows")
