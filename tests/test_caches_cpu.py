"""Derived-tensor caches of the host side: a cache entry must never survive its model -- the reference's own
QP 22/27/32/37 sweep builds, frees and rebuilds models in one process, CPython reuses the freed addresses and every parameter of a
freshly loaded checkpoint has the same _version."""
import gc

import torch

from cdfo_b200 import _lib, hotpath
from cdfo_b200.model import CVSR_V8, LLongRangAttention, _CrossScaleBlock


def test_tensor_cache_identity_version_and_cleanup():
    c = _lib.TensorCache()
    w = torch.nn.Parameter(torch.ones(4))
    calls = []
    build = lambda: calls.append(1) or (w.detach() * 2).clone()  # noqa: E731
    a = c.get(w, build)
    assert c.get(w, build) is a and len(calls) == 1
    with torch.no_grad():
        w.add_(1.0)                                   # in-place edit bumps _version
    b = c.get(w, build)
    assert len(calls) == 2 and float(b[0]) == 4.0
    w.data = torch.zeros(4)                           # .to(device) / load through .data: new storage address
    assert float(c.get(w, build)[0]) == 0.0 and len(calls) == 3
    assert c.get(w, build, kind="other") is not c.get(w, build)
    del w
    gc.collect()
    assert len(c) == 0                                # the weak reference's callback dropped the entries


def test_module_caches_do_not_leak_between_models():
    outs = []
    for seed in (1, 2, 3, 4):
        torch.manual_seed(seed)
        blk = _CrossScaleBlock()
        wts = hotpath._block_weights(blk)
        assert hotpath._block_weights(blk) is wts     # hit
        ref = hotpath._compose_1x1_after_3x3(blk.up._modules["0"].weight.detach(), blk.up._modules["0"].bias.detach(),
                                             blk.body._modules["2"].weight.detach(), blk.body._modules["2"].bias.detach())
        assert torch.equal(wts["up_body2"][0], ref[0]) and torch.equal(wts["up_body2"][1], ref[1])
        outs.append(wts["up_body2"][0].clone())
        del blk, wts
        gc.collect()
    assert not torch.equal(outs[0], outs[1])


def test_module_cache_follows_load_state_dict_and_clear():
    torch.manual_seed(0)
    a, b = LLongRangAttention(64), LLongRangAttention(64)
    ta = hotpath._lra_tap_tables(a)
    a.load_state_dict(b.state_dict())                 # copy_ bumps the versions
    tb = hotpath._lra_tap_tables(a)
    assert tb is not ta and torch.equal(tb[0], hotpath._lra_tap_tables(b)[0])
    m = CVSR_V8()
    hotpath._lra_tap_tables(m.RDAB)
    assert "_cdfo_derived" in m.RDAB.__dict__
    assert "_cdfo_derived" not in "".join(m.state_dict().keys())
    hotpath.clear_derived(m)
    assert "_cdfo_derived" not in m.RDAB.__dict__
