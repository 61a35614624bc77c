"""Generate tests/golden_tf/*.npz: the reference's TEACHER-FORCED path (forward() without the losses) run on single,
unpadded utterances in eval mode (build container only).

    python -m oracle.make_golden_tf

TEST INFRASTRUCTURE. `reference_forward_tf` executes the UNMODIFIED reference modules in the order of
Tacotron2_sa.forward (nets/teacher_training/e2e_tts_tacotron2_sa.py:545-595): enc(xs, ilens) -> duration / pitch /
energy predictors -> pitch/energy embeddings of the GROUND-TRUTH f0 / energy -> dec(...) (nets/modules/decoder_sa.py:431-542)
with the re-organised targets the CustomConverter builds (tts.py:236-301: new_ys, non_zero_lens_mask, ds_nonzeros,
output_masks, position; phonemes of duration 0 are dropped). The losses and the reporter are not called.
Prenet dropout: rate 0 or the counter-based mask of oracle/philox.py (keyed by the ORIGINAL phoneme index).
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader, philox            # noqa: E402
from fcl_taco2_b200 import synth, hparams        # noqa: E402

CASES = [
    # name, kind, weight seed, input seed, N, dropout rate, dropout seed, utt index
    ("S_tf_n40_nodrop", "S", 0, 10, 40, 0.0, 0, 0),
    ("S_tf_n40_drop", "S", 1, 11, 40, 0.5, 4711, 3),
    ("T_tf_n24_drop", "T", 0, 12, 24, 0.5, 99, 1),
]


def reference_forward_tf(m, x, ys, d, f0, energy, dropout_fn=None):
    """-> dict(after (L, odim), before (L, odim), d_outs, p_outs, e_outs) from the reference modules, batch of one."""
    from espnet.nets.pytorch_backend.nets_utils import make_pad_mask, make_non_pad_mask, pad_list
    xs = x.unsqueeze(0)
    ilens = torch.LongTensor([x.shape[0]])
    olens = torch.LongTensor([ys.shape[0]])
    # ---- CustomConverter, tts.py:236-301 (reduction_factor 1)
    new_ys, mask, ds_nonzeros, position = [], [], [], []
    for it in range(x.shape[0]):
        start, end = int(d[:it].sum()), int(d[:it + 1].sum())
        if start != end:
            new_ys.append(ys[start:end].float())
            mask.append(1)
            ds_nonzeros.append(int(d[it]))
            position.append(torch.FloatTensor(list(range(end - start))) / (end - start))
        else:
            mask.append(0)
    new_ys = pad_list(new_ys, 0)
    non_zero_lens_mask = pad_list([torch.tensor(mask)], 0)
    ds_nonzeros = torch.tensor(ds_nonzeros)
    output_masks = make_non_pad_mask(ds_nonzeros)
    position = pad_list(position, 0)
    f0b, enb = f0.reshape(1, -1, 1).float(), energy.reshape(1, -1, 1).float()
    # ---- Tacotron2_sa.forward, e2e_tts_tacotron2_sa.py:553-594
    hs, hlens = m.enc(xs, ilens)[:2]
    d_masks = make_pad_mask(ilens)
    d_outs = m.duration_predictor(hs, d_masks)
    p_outs = m.pitch_predictor(hs, d_masks.unsqueeze(-1))
    e_outs = m.energy_predictor(hs, d_masks.unsqueeze(-1))
    p_embs = m.pitch_embed(f0b.transpose(1, 2)).transpose(1, 2)
    e_embs = m.energy_embed(enb.transpose(1, 2)).transpose(1, 2)
    ds = d.reshape(1, -1).float()

    def run():
        import inspect
        if "f0" in inspect.signature(m.dec.forward).parameters:          # teacher (decoder_sa.py:431-432)
            return m.dec(hs, hlens, ds, ys.unsqueeze(0), olens, new_ys, non_zero_lens_mask, ds_nonzeros, output_masks, position,
                         f0b, enb, p_embs, e_embs)
        return m.dec(hs, hlens, ds, ys.unsqueeze(0), olens, new_ys, non_zero_lens_mask, ds_nonzeros, output_masks, position,
                     p_embs, e_embs)                                     # student (kd_student.py:744-747, decoder_sa_kd.py:523)
    if dropout_fn is None:
        res = run()
    else:
        with ref_loader.prenet_dropout(dropout_fn):
            res = run()
    after, before = res[0], res[1]
    return dict(after=after[0], before=before[0], d_outs=d_outs[0], p_outs=p_outs[0, :, 0], e_outs=e_outs[0, :, 0])


def make_inputs(kind, iseed, n):
    """ids, durations with zeros sprinkled in (first / a run / last), targets, f0, energy."""
    xs, ds = synth.synth_batch(1, iseed, fixed_len=n)
    rs = np.random.RandomState(iseed)
    d = ds[0].copy()
    d[[0, n // 3, n // 3 + 1, n - 1]] = 0
    L = int(d.sum())
    ys = (rs.randn(L, 80) * 0.8).astype(np.float32)
    return xs[0], d, ys, rs.randn(n).astype(np.float32), rs.randn(n).astype(np.float32)


def run_case(kind, wseed, iseed, n, rate, dseed, utt):
    hp = hparams.preset(kind)
    sd = synth.random_state_dict(hp, wseed, kind == "S", hparams.preset("T"))
    m = ref_loader.build(kind)
    m.load_state_dict(sd, strict=True)
    x, d, ys, f0, en = make_inputs(kind, iseed, n)
    keep = np.nonzero(d > 0)[0]
    m.dec.prenet.dropout_rate = rate
    fn = None
    if rate > 0.0:
        def fn(inp, p, i):
            km = philox.keep_mask(dseed, np.full(len(keep), utt), keep, i // 2, i % 2, inp.shape[1], p)
            return inp * (torch.from_numpy(km).float() * (1.0 / (1.0 - p)))
    with torch.no_grad():
        r = reference_forward_tf(m, torch.from_numpy(x), torch.from_numpy(ys), torch.from_numpy(d), torch.from_numpy(f0),
                                 torch.from_numpy(en), fn)
    return dict(kind=kind, weight_seed=wseed, weights_sha256=synth.state_dict_digest(sd), x=x, dur=d, ys=ys, f0=f0, energy=en,
                after=r["after"].numpy(), before=r["before"].numpy(), d_outs=r["d_outs"].numpy(), p_outs=r["p_outs"].numpy(),
                e_outs=r["e_outs"].numpy(), dropout_rate=rate, dropout_seed=dseed, utt_index=utt)


def main():
    assert ref_loader.available(), "needs /root/reference"
    outdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden_tf")
    os.makedirs(outdir, exist_ok=True)
    torch.manual_seed(0)
    for name, *cfg in CASES:
        r = run_case(*cfg)
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **r)
        print(name, r["after"].shape, float(np.abs(r["after"]).mean()))


if __name__ == "__main__":
    main()
