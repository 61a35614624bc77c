"""CPU tests of the host side: C-ABI library loads and exports every declared symbol,
struct layouts agree, batch planning, weight packing, drop-in model container."""
import argparse
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from fcl_taco2_b200 import _lib, hparams, pack, plan as planmod, synth
from tests.helpers import weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "fcl_taco2.h")).read()
    declared = set(re.findall(r"^(?:int|const char\*)\s+(fcl_\w+)\s*\(", hdr, flags=re.M))
    assert declared == set(_lib.ENTRY_POINTS) | set(_lib.PLAIN_SYMBOLS), declared
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.fcl_abi_version() == _lib.ABI_VERSION
    for i, st in enumerate(_lib.STRUCTS):
        assert lib.fcl_struct_size(i) == ctypes.sizeof(st)
    assert lib.fcl_struct_size(99) == -1


def test_entry_points_reject_bad_arguments_without_gpu():
    lib = _lib.load()
    p = _lib.ConvGemmParams()                     # all NULL
    rc = lib.fcl_conv_gemm_f32(ctypes.byref(p), None)
    assert rc == -1 and b"null pointer" in lib.fcl_last_error()
    with pytest.raises(_lib.FclError):
        _lib.call("fcl_len_reg_scan", _lib.LenRegParams(), 0)


def test_engine_refuses_cpu():
    from fcl_taco2_b200.engine import Engine
    hp = hparams.preset("S")
    with pytest.raises(_lib.FclError):
        Engine(hp, pack.pack_fp32(weights("S", 0), hp), "cpu")


def test_plan_packing_and_order():
    xs = [np.arange(1, 6), np.arange(1, 10), np.arange(1, 3)]
    ds = [np.full(5, 2), np.full(9, 3), np.full(2, 1)]
    pl = planmod.make_plan(xs, ds, utt_ids=[10, 11, 12])
    assert pl.perm.tolist() == [1, 0, 2] and pl.n_rows == 16
    assert pl.utt_off.tolist() == [0, 9, 14, 16]
    assert pl.row_utt[:9].tolist() == [11] * 9 and pl.row_utt[9:14].tolist() == [10] * 5
    assert pl.row_phone[9:14].tolist() == [0, 1, 2, 3, 4]
    assert pl.seg_lo[9] == 9 and pl.seg_hi[9] == 14
    assert pl.dur[:9].tolist() == [3] * 9
    with pytest.raises(ValueError):
        planmod.make_plan([np.arange(3)], [np.arange(2)])
    with pytest.raises(ValueError):
        planmod.make_plan([np.arange(0)])
    with pytest.raises(ValueError):
        planmod.make_plan([np.arange(3)], f0s=[np.zeros(3)])


def test_plan_matches_per_utterance_loop():
    """make_plan (vectorised, int32 np.repeat arithmetic) against the obvious per-utterance loop, for ragged random
    batches incl. lists / (N,1) columns as inputs, forced prosody and compensating length errors."""
    rs = np.random.RandomState(7)
    for B in (1, 2, 17, 64):
        lens = rs.randint(1, 40, size=B)
        xs = [rs.randint(1, 76, size=n) for n in lens]
        ds = [rs.randint(1, 9, size=n) for n in lens]
        f0 = [rs.randn(n).astype(np.float32) for n in lens]
        en = [rs.randn(n, 1).astype(np.float32) for n in lens]            # (N, 1) columns take the slow path
        ids = rs.permutation(1000)[:B]
        pl = planmod.make_plan([x.tolist() if i % 3 == 0 else x for i, x in enumerate(xs)], ds, f0, en, utt_ids=ids)
        order = sorted(range(B), key=lambda i: (-lens[i], i))
        assert pl.perm.tolist() == order and pl.n_utts == B and pl.n_rows == int(lens.sum())
        o = 0
        for k, i in enumerate(order):
            n = int(lens[i])
            sl = slice(o, o + n)
            assert pl.utt_off[k] == o and pl.utt_off[k + 1] == o + n
            assert np.array_equal(pl.ids[sl], xs[i]) and np.array_equal(pl.dur[sl], ds[i])
            assert np.array_equal(pl.pitch[sl], f0[i]) and np.array_equal(pl.energy[sl], en[i][:, 0])
            assert (pl.row_utt[sl] == ids[i]).all() and np.array_equal(pl.row_phone[sl], np.arange(n))
            assert (pl.seg_lo[sl] == o).all() and (pl.seg_hi[sl] == o + n).all()
            o += n
        for name in ("utt_off", "row_utt", "row_phone", "seg_lo", "seg_hi", "dur"):
            assert getattr(pl, name).dtype == np.int32, name
    # two length errors that cancel in the total must still be caught
    with pytest.raises(ValueError):
        planmod.make_plan([np.arange(1, 4), np.arange(1, 4)], [np.ones(4, int), np.ones(2, int)])
    with pytest.raises(ValueError):
        planmod.make_plan([np.arange(1, 4)], [np.array([1, -1, 2])])


def test_shard_utterances_balanced_and_complete():
    rs = np.random.RandomState(0)
    costs = rs.randint(50, 1000, size=103)
    for ws in (1, 2, 4, 8):
        sh = planmod.shard_utterances(costs, ws)
        assert sorted(sum(sh, [])) == list(range(103))
        loads = [costs[s].sum() for s in sh]
        assert max(loads) - min(loads) <= costs.max()


def test_pack_layouts():
    hp = hparams.preset("S")
    sd = weights("S", 0)
    pk = pack.pack_fp32(sd, hp)
    H, E, U = hp.dunits, hp.eunits, hp.prenet_units
    # gate interleave: column u*4+g of the packed matrix is row g*H+u of the torch matrix
    w = sd["dec.lstm.1.cell.weight_hh"]
    assert torch.equal(pk["dec_w1"][H + 3, 7 * 4 + 2], w[2 * H + 7, 3])
    assert torch.equal(pk["dec_w0"][5, 9 * 4 + 1], sd["dec.lstm.0.cell.weight_ih"][1 * H + 9, E + 5])
    assert torch.equal(pk["dec_wpos"][9 * 4 + 3], sd["dec.lstm.0.cell.weight_ih"][3 * H + 9, E + U])
    assert pk["dec_g0h_w"].shape == (1, E, 4 * H) and pk["dec_y0h_w"].shape == (1, E, 80)
    # BN fold reproduces conv+BN
    x = torch.randn(1, E, 11)
    ref = torch.nn.functional.batch_norm(
        torch.nn.functional.conv1d(x, sd["enc.convs.1.0.weight"], None, 1, 2), sd["enc.convs.1.1.running_mean"],
        sd["enc.convs.1.1.running_var"], sd["enc.convs.1.1.weight"], sd["enc.convs.1.1.bias"], False, 0.0, 1e-5)
    wk = pk["enc_conv1_w"].permute(2, 1, 0).contiguous()
    got = torch.nn.functional.conv1d(x, wk, pk["enc_conv1_b"], 1, 2)
    assert (ref - got).abs().max() < 1e-5


def test_model_container_matches_reference_layout():
    from fcl_taco2_b200 import model as M
    for kind in "ST":
        m = M.from_preset(kind, seed=None, device="cpu", kd_keys=(kind == "S"))
        spec = synth.state_dict_spec(hparams.preset(kind), kind == "S", hparams.preset("T"))
        sd = m.state_dict()
        assert list(sd.keys()) == list(spec.keys())
        assert all(tuple(sd[k].shape) == tuple(spec[k][0]) for k in spec)
        m.load_state_dict(weights(kind, 0), strict=True)
        with pytest.raises(RuntimeError):
            bad = dict(weights(kind, 0))
            bad.pop("dec.feat_out.weight")
            m.load_state_dict(bad, strict=True)
    # student checkpoints with or without KD-only tensors both load
    m = M.from_preset("S", seed=None, device="cpu", kd_keys=False)
    m.load_state_dict(weights("S", 0), strict=True)


def test_unsupported_configs_raise():
    from fcl_taco2_b200 import model as M
    ns = argparse.Namespace(**{k: v for k, v in hparams.preset("S").to_dict().items() if k not in ("idim", "odim")})
    ns.reduction_factor = 2
    with pytest.raises(ValueError):
        M.Tacotron2_sa(76, 80, ns, argparse.Namespace(use_fe_condition=True, append_position=True))
    with pytest.raises(ValueError):
        M.Tacotron2_sa(76, 80)      # reference flag defaults (use_residual=True) are not the FCL configs


def test_add_arguments_mirror():
    from fcl_taco2_b200 import model as M
    p = argparse.ArgumentParser()
    M.Tacotron2_sa.add_arguments(p)
    a = p.parse_args(["--dunits", "256", "--use-residual", "false"])
    assert a.dunits == 256 and a.use_residual is False and a.prenet_units == 256 and a.zoneout_rate == 0.1
