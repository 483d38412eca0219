"""Host-side logic that needs no GPU: the folded-weight cache (torch.nn.DataParallel replicas share it), the conditions
under which a module takes the fused path, and pickling / deepcopy of modules that hold caches."""
import copy
import pickle
import threading

import pytest
import torch

from pn2_b200 import _lib, models
from pn2_b200 import pointnet_util as pu


def _sa():
    torch.manual_seed(0)
    return pu.PointNetSetAbstraction(16, 0.3, 8, 3 + 3, [8, 16], False).eval()


def test_fold_cache_hits_until_a_folded_tensor_changes():
    sa = _sa()
    a = sa.folded()
    assert sa.folded() is a
    with torch.no_grad():
        sa.mlp_bns[0].running_var.mul_(2.0)      # in-place update bumps the version counter
    b = sa.folded()
    assert b is not a and not torch.equal(a.layers[0][0], b.layers[0][0])
    sa.train()                                    # train() drops the entries; eval() re-folds
    sa.eval()
    assert sa.folded() is not b


def test_fold_cache_is_shared_safely_between_replica_threads():
    """DataParallel replicas shallow-copy __dict__: they share the cache object.  Replicas never cache (their parameter
    pointers are recycled broadcast copies) and concurrent get() calls each return a consistent value."""
    sa = _sa()
    replica = sa._replicate_for_data_parallel()
    assert replica._fold is sa._fold
    replica._is_replica = True
    # replicas keep parameters as plain attributes; give them the originals for this host-only check
    for m_r, m in zip(replica.modules(), sa.modules()):
        m_r._parameters = m._parameters
    r1, r2 = replica.folded(), replica.folded()
    assert r1 is not r2 and torch.equal(r1.layers[1][0], r2.layers[1][0])
    got, errs = [], []

    def work():
        try:
            for _ in range(50):
                f = sa.folded()
                assert f.cin == 6 and f.cout == 16 and len(f.layers) == 2
                got.append(f)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=work) for _ in range(8)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs and len(got) == 400


def test_modules_with_caches_pickle_and_deepcopy():
    net = models.PointNet2SemSeg(5).eval()
    net.sa1.folded()
    net._fp1_with_head()
    clone = copy.deepcopy(net)
    assert clone.sa1._fold is not net.sa1._fold and not clone.sa1._fold._entries
    again = pickle.loads(pickle.dumps(net))
    assert again._streams == {} and again.sa1.folded().cout == 64
    assert sorted(again.state_dict()) == sorted(net.state_dict())


def test_fused_path_only_for_eval_fp32_cuda_and_supported_shapes():
    sa = _sa()
    x, f = torch.zeros(1, 3, 64), torch.zeros(1, 3, 64)
    assert not pu._fusable(sa, x, f)                      # host tensors: the composed path raises Pn2Error (no CPU fallback)
    assert not pu._fusable(sa.train(), x, f)
    sa.eval()
    meta = lambda t, dt=torch.float32: torch.empty(t.shape, dtype=dt, device="meta")
    assert not pu._all_f32(meta(x, torch.float16))
    assert not pu._all_f32(meta(x), meta(f, torch.float64))
    assert pu._all_f32(meta(x), None, meta(f))
    assert not pu._fusable(copy.deepcopy(sa).half(), x, f)
    # shapes the fused kernel does not serve are reported by the library, so the modules fall back instead of raising
    folded = sa.folded()
    assert pu._fp32_supported(folded, 6, 8) and pu._fp32_supported(folded, 6, 128)
    assert not pu._fp32_supported(folded, 6, 24) and not pu._fp32_supported(folded, 6, 256)
    assert not pu._fp32_supported(folded, 7, 8)           # channel mismatch
    deep = pu.PointNetSetAbstraction(16, 0.3, 8, 6, [8] * 7, False).eval()
    assert len(deep.mlp_convs) > _lib.PN2_MAX_LAYERS and not deep._fused_ok(x, f)
    # shared-memory plan of the fp32 kernel (row_mlp.cu make_layout): layers of one n-tile run in place, so a stack of
    # <= 128-wide layers needs ONE activation buffer (2512 x 20 floats at the 16-row tile: fits); a 256-wide layer has two
    # n-tiles and needs the second buffer as well (does not fit), as does a wider input
    wide_in = lambda widths, c0=2500: pu.PointNetFeaturePropagation(c0, widths).eval().folded()
    assert pu._fp32_supported(wide_in([128, 128]), 2500, 0)
    assert not pu._fp32_supported(wide_in([256, 128]), 2500, 0)
    assert not pu._fp32_supported(wide_in([128, 128], 3000), 3000, 0)


def test_dtype_checks_refuse_reinterpretation():
    with pytest.raises(_lib.Pn2Error):
        _lib.check_f32(torch.zeros(4, dtype=torch.float16, device="meta").to("cpu") if False else torch.zeros(4), "x")  # host tensor
    assert _lib.check_f32(None) is None and _lib.check_i32(None) is None


def test_default_precision_is_fp32():
    assert pu.get_mlp_precision() == "fp32"
    prev = pu.set_mlp_precision("bf16")
    assert prev == "fp32" and pu.set_mlp_precision(prev) == "bf16"
    with pytest.raises(ValueError):
        pu.set_mlp_precision("fp16")
