"""Drop-in `Tacotron2_sa` for the reference's inference boundary.

Same constructor, `add_arguments`, `state_dict` key layout and `inference()`
signature as
  nets/teacher_training/e2e_tts_tacotron2_sa.py:136-289,624-683            (FCL-taco2-T)
  nets/knowledge_distillation/e2e_tts_tacotron2_sa_kd_student.py:392,804-863 (FCL-taco2-S)
but the arithmetic runs in the sm_100a kernels of libfcl_taco2.so through
`fcl_taco2_b200.engine.Engine`. The torch Parameters held here are only the
checkpoint container: they are repacked once (fcl_taco2_b200.pack) when the
first inference call is made after loading / moving the model.

New, not in the reference: `inference_batch()` (many utterances per call; the
reference is single-utterance, tts.py:655-674) and explicit control of the
always-on prenet dropout (`set_prenet_dropout`).
Out of scope (raises): `forward()` / losses (training), speaker embeddings, r > 1.
"""
from __future__ import annotations

import argparse
import itertools

import numpy as np
import torch

from . import pack, plan as planmod, synth
from .engine import Engine
from .hparams import HParams, from_namespace

try:  # the reference's runners assert isinstance(model, TTSInterface) (tts.py:620)
    from espnet.nets.tts_interface import TTSInterface as _TTSBase
except Exception:  # espnet is not installed in this image
    class _TTSBase:  # noqa: D401
        """Stand-in with the TTSInterface surface the runners touch."""

        def __init__(self):
            self.reporter = None

        @property
        def attention_plot_class(self):
            return None

        @property
        def base_plot_keys(self):
            return []


def _strtobool(x):
    s = str(x).strip().lower()
    if s in ("y", "yes", "t", "true", "on", "1"):
        return True
    if s in ("n", "no", "f", "false", "off", "0"):
        return False
    raise ValueError(f"invalid truth value {x!r}")


# (flag, default, type) -- the model-specific flags of e2e_tts_tacotron2_sa.py:138-287
_ARGS = [
    ("--embed-dim", 512, int), ("--elayers", 1, int), ("--eunits", 512, int), ("--econv-layers", 3, int),
    ("--econv-chans", 512, int), ("--econv-filts", 5, int), ("--dlayers", 2, int), ("--dunits", 1024, int),
    ("--prenet-layers", 2, int), ("--prenet-units", 256, int), ("--postnet-layers", 5, int),
    ("--postnet-chans", 512, int), ("--postnet-filts", 5, int), ("--output-activation", None, str),
    ("--use-batch-norm", True, _strtobool), ("--use-concate", True, _strtobool), ("--use-residual", True, _strtobool),
    ("--dropout-rate", 0.5, float), ("--zoneout-rate", 0.1, float), ("--reduction-factor", 1, int),
    ("--spk-embed-dim", None, int), ("--spc-dim", None, int), ("--pretrained-model", None, str),
    ("--use-masking", False, _strtobool), ("--use-weighted-masking", False, _strtobool),
    ("--duration-predictor-layers", 2, int), ("--duration-predictor-chans", 384, int),
    ("--duration-predictor-kernel-size", 3, int), ("--duration-predictor-dropout-rate", 0.1, float),
]

_KD_ONLY_PREFIXES = ("enc.embed_proj.", "enc.convs_proj.", "enc.blstm_proj.", "dec.prenet_proj.", "dec.lstm_proj.",
                     "dec.lstm0_proj.", "dec.lstm1_proj.", "dec.post_proj.", "dec.post0_proj.", "dec.post1_proj.",
                     "dec.post2_proj.", "dec.post3_proj.", "pemb_proj.", "eemb_proj.")


def _register(root: torch.nn.Module, dotted: str, tensor: torch.Tensor, is_buffer: bool):
    parts = dotted.split(".")
    m = root
    for name in parts[:-1]:
        if name not in m._modules:
            m.add_module(name, torch.nn.Module())
        m = m._modules[name]
    if is_buffer:
        m.register_buffer(parts[-1], tensor)
    else:
        m.register_parameter(parts[-1], torch.nn.Parameter(tensor, requires_grad=False))


class Tacotron2_sa(_TTSBase, torch.nn.Module):
    """FCL-taco2 acoustic model, inference only, B200-native."""

    _student = False

    @staticmethod
    def add_arguments(parser):
        group = parser.add_argument_group("tacotron 2 model setting")
        for flag, default, typ in _ARGS:
            group.add_argument(flag, default=default, type=typ)
        return parser

    def __init__(self, idim, odim, args=None, com_args=None, teacher_args=None, precision="fp32"):
        _TTSBase.__init__(self)
        torch.nn.Module.__init__(self)
        # fill missing arguments from the flag defaults (espnet fill_missing_args semantics)
        defaults = {f.lstrip("-").replace("-", "_"): d for f, d, _ in _ARGS}
        d = dict(defaults)
        if args is not None:
            d.update(vars(args))
        ns = argparse.Namespace(**d)
        self.hp: HParams = from_namespace(idim, odim, ns, com_args)
        self.hp.validate()
        self.idim, self.odim = idim, odim
        self.embed_dim = self.hp.embed_dim
        self.spk_embed_dim = None
        self.reduction_factor = 1
        self.use_fe_condition = True
        self.append_position = True
        self.precision = precision

        teacher_hp = None
        if self._student and teacher_args is not None:
            teacher_hp = from_namespace(idim, odim, teacher_args, com_args)
        self._kd_keys = teacher_hp is not None
        spec = synth.state_dict_spec(self.hp, student_kd_keys=self._kd_keys, teacher=teacher_hp)
        for name, (shape, kind) in spec.items():
            is_buf = kind in ("bn_mean", "bn_var", "bn_count")
            t = torch.zeros(shape, dtype=torch.int64 if kind == "bn_count" else torch.float32)
            if kind in ("bn_var", "bn_weight", "ln_weight"):
                t.fill_(1.0)
            _register(self, name, t, is_buf)
        self._engine = None
        self._dropout_rate = float(self.hp.dropout_rate)     # prenet dropout is ON at inference (decoder_sa.py:156-157)
        self._dropout_seed = None                             # a fresh mask per call, like the reference's F.dropout
        self._calls = itertools.count()
        self.eval()

    # ------------------------------------------------------------------ checkpoint container
    def init_random(self, seed: int = 0):
        """Seeded random-init weights (no checkpoints exist offline); see synth.random_state_dict."""
        teacher = None
        if self._kd_keys:
            from .hparams import preset
            teacher = preset("T")
        sd = synth.random_state_dict(self.hp, seed, self._kd_keys, teacher)
        self.load_state_dict(sd, strict=True)
        return self

    def load_state_dict(self, state_dict, strict=True, **kw):
        """Strict on every inference tensor; KD-only projection tensors (present in student
        checkpoints, unused by inference -- SURVEY.md Appendix B) may be present or absent."""
        own = set(self.state_dict().keys())
        given = set(state_dict.keys())
        missing = [k for k in own - given if not k.startswith(_KD_ONLY_PREFIXES)]
        unexpected = [k for k in given - own if not k.startswith(_KD_ONLY_PREFIXES)]
        if strict and (missing or unexpected):
            raise RuntimeError(f"load_state_dict: missing {sorted(missing)}, unexpected {sorted(unexpected)}")
        filtered = {k: v for k, v in state_dict.items() if k in own}
        res = torch.nn.Module.load_state_dict(self, filtered, strict=False)
        self._engine = None
        return res

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    @property
    def device(self):
        return next(self.parameters()).device

    def engine(self) -> Engine:
        if self._engine is None:
            self._engine = Engine(self.hp, pack.pack_fp32(self.state_dict(), self.hp), self.device, self.precision)
        return self._engine

    # ------------------------------------------------------------------ dropout policy
    def set_prenet_dropout(self, rate=None, seed=None):
        """rate 0 -> deterministic (exact-parity mode); default rate = conf dropout-rate (0.5) with a fresh mask on
        every call, the reference's behaviour (decoder_sa.py:156-157). `seed` (an int) fixes the counter-based mask for
        reproducible output (tests, bench); None -> a fresh seed per call."""
        if rate is not None:
            if not 0.0 <= rate < 1.0:
                raise ValueError("dropout rate must be in [0, 1)")
            self._dropout_rate = float(rate)
        self._dropout_seed = seed
        return self

    def _seed_for_call(self):
        n = next(self._calls)
        return int(self._dropout_seed) if self._dropout_seed is not None else (0x9E3779B97F4A7C15 * (n + 1)) & ((1 << 64) - 1)

    # ------------------------------------------------------------------ inference
    @torch.no_grad()
    def forward_teacher_forced(self, xs, ys, durs, f0s, energies, utt_ids=None):
        """Teacher-forced pass over a batch of utterances -- the computation of the reference's `forward()` without the
        losses (e2e_tts_tacotron2_sa.py:520-595 -> decoder_sa.py:431-542; what the KD trainer asks of the teacher at every
        step, tts_distill.py:159-162), in eval-mode semantics and per utterance (no padding, so no pad leakage):
        ground-truth durations / f0 / energy condition the decoder, the decoder input of step m of a phoneme is its
        ground-truth frame m-1 (`prev_out = y`), phonemes of duration 0 get no decoder row (decoder_sa.py:459-463).

        xs: list of (N_i,) ids; ys: list of (L_i, odim) target mels with L_i = sum(durs[i]); durs / f0s / energies:
        lists of (N_i,). -> list of dicts {after (L_i, odim), before (L_i, odim), d_outs (N_i,) log-durations of the
        duration predictor, p_outs (N_i,), e_outs (N_i,)} (device tensors), in the order given."""
        to_np = lambda v: v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
        ys = [np.ascontiguousarray(to_np(y), dtype=np.float32).reshape(-1, self.odim) for y in ys]
        pl = self._plan(xs, durs, f0s, energies, utt_ids)
        per_utt = np.add.reduceat(pl.dur.astype(np.int64), pl.utt_off[:-1].astype(np.int64))
        for k, i in enumerate(pl.perm):
            if ys[int(i)].shape[0] != int(per_utt[k]):
                raise ValueError(f"ys[{int(i)}] has {ys[int(i)].shape[0]} frames for durations summing to {int(per_utt[k])}")
        if int(per_utt.sum()) == 0:
            raise ValueError("every duration is zero: no frames to decode")
        tf_y = torch.from_numpy(np.concatenate([ys[int(i)] for i in pl.perm])).to(self.device)      # processing order
        eng = self.engine()
        eng.skip_zero_durations = True
        try:
            res = eng.run(pl, self.hp.zoneout_rate, self._dropout_rate, self._seed_for_call(), tf_y=tf_y)
        finally:
            eng.skip_zero_durations = False
        ex, outs = res.extras, [None] * pl.n_utts
        for k, i in enumerate(pl.perm):
            f0_, f1_ = int(res.utt_frame_off[k]), int(res.utt_frame_off[k + 1])
            r0, r1 = int(pl.utt_off[k]), int(pl.utt_off[k + 1])
            outs[int(i)] = {"after": res.out[f0_:f1_], "before": ex["before"][f0_:f1_], "d_outs": ex["dlog"][r0:r1],
                            "p_outs": ex["pitch_pred"][r0:r1], "e_outs": ex["energy_pred"][r0:r1]}
        return outs

    @torch.no_grad()
    def forward(self, xs, ilens, ys, olens, spembs=None, extras=None, new_ys=None, non_zero_lens_mask=None, ds_nonzeros=None,
                output_masks=None, position=None, f0=None, energy=None, *args, **kwargs):
        """The reference's `forward()` argument list (e2e_tts_tacotron2_sa.py:520-523; padded batch tensors from the
        CustomConverter, tts.py:215-302) WITHOUT the losses: -> (after_outs, before_outs), both (B, Lmax, odim) zero-padded,
        as the KD teacher's forward returns them first (e2e_tts_tacotron2_sa_kd_teacher.py:603). Training (losses,
        gradients) stays out of scope. `extras` = durations (B, Tmax[, 1]); f0 / energy (B, Tmax[, 1]) are required
        (use_fe_condition). The re-organised `new_ys / non_zero_lens_mask / ds_nonzeros / output_masks / position` are
        derived on the device from the durations and ignored here. Each utterance is processed unpadded."""
        if spembs is not None:
            raise ValueError("speaker embeddings (spk_embed_dim) are out of scope of the B200 path")
        if extras is None or f0 is None or energy is None:
            raise ValueError("forward() needs the ground-truth durations (extras), f0 and energy")
        B = int(xs.shape[0])
        il, ol = [int(v) for v in ilens], [int(v) for v in olens]
        ds = extras.reshape(B, -1)
        xs_l = [xs[b, :il[b]] for b in range(B)]
        ys_l = [ys[b, :ol[b]] for b in range(B)]
        ds_l = [ds[b, :il[b]].long() for b in range(B)]
        f0_l = [f0.reshape(B, -1)[b, :il[b]].float() for b in range(B)]
        en_l = [energy.reshape(B, -1)[b, :il[b]].float() for b in range(B)]
        outs = self.forward_teacher_forced(xs_l, ys_l, ds_l, f0_l, en_l)
        lmax = max(ol)
        after = torch.zeros((B, lmax, self.odim), dtype=torch.float32, device=self.device)
        before = torch.zeros_like(after)
        for b in range(B):
            after[b, :ol[b]] = outs[b]["after"]
            before[b, :ol[b]] = outs[b]["before"]
        return after, before

    def _plan(self, xs, durs=None, f0s=None, energies=None, utt_ids=None):
        """Host side of a batch: validate, order longest-first, flatten (pure numpy, no CUDA calls)."""
        to_np = lambda v: v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
        xs = [x if type(x) is np.ndarray else to_np(x) for x in xs]
        if any(x.ndim != 1 for x in xs):
            raise ValueError("each utterance must be a 1-D id sequence (encoder_sa.py:157)")
        conv = lambda vs: None if vs is None else [v if type(v) is np.ndarray else to_np(v) for v in vs]
        pl = planmod.make_plan(xs, conv(durs), conv(f0s), conv(energies), utt_ids)
        if pl.ids.min() < 0 or pl.ids.max() >= self.idim:
            raise ValueError("phoneme id out of range")
        return pl

    @torch.no_grad()
    def inference_batch(self, xs, durs=None, f0s=None, energies=None, utt_ids=None, return_result=False,
                        skip_zero_durations=False):
        """Batched form of `inference`: xs is a list of 1-D id sequences (LongTensor / ndarray / list).
        -> list of (L_i, odim) float32 tensors on the model device, in the order given.
        `skip_zero_durations=True` (extension; the reference's inference asserts, decoder_sa.py:575): phonemes whose
        forced or predicted duration is 0 still feed the encoder / predictors / prosody embeddings but get no decoder
        row and no frames -- the semantics of the reference's teacher-forced forward() (decoder_sa.py:459-463)."""
        pl = self._plan(xs, durs, f0s, energies, utt_ids)
        eng = self.engine()
        eng.skip_zero_durations = bool(skip_zero_durations)
        try:
            res = eng.run(pl, self.hp.zoneout_rate, self._dropout_rate, self._seed_for_call())
        finally:
            eng.skip_zero_durations = False
        return res if return_result else res.per_utterance()

    @torch.no_grad()
    def inference_stream(self, batches, depth: int = 3, before_batch=None, planner_thread: bool = False):
        """Decode a sequence of batches with the device->host transfer of batch i overlapped with the compute of
        batch i+1 (what a decode driver over a data set wants: the mels of a batch are ~180 MB, their PCIe transfer
        takes about as long as half of the pass that produced them).

        `batches`: iterable of `xs` lists or of dicts with the keyword arguments of `inference_batch`
        (xs, durs, f0s, energies, utt_ids). Yields, in order, one list of (L_i, odim) float32 numpy arrays per batch
        (caller's utterance order). The arrays are views of a pinned staging buffer that is reused `depth` batches
        later: consume or copy them before advancing that far. `before_batch()` (optional) is called on the calling
        thread right before a batch is enqueued (bench.py flushes L2 there). `planner_thread=True` moves the host-side
        planning of the following batches to a helper thread."""
        dev = self.device
        copy_stream = getattr(self, "_copy_stream", None)
        if copy_stream is None:
            copy_stream = self._copy_stream = torch.cuda.Stream(dev)
            self._host_ring = []
        pending = []                                                  # (result, host buffer, "copied" event)

        def finish(item):
            res, buf, ev = item
            ev.synchronize()
            host = buf[: res.out.shape[0]].numpy()
            lo, hi = res.utt_frame_off[:-1].tolist(), res.utt_frame_off[1:].tolist()
            return [host[lo[k]:hi[k]] for k in np.argsort(res.perm, kind="stable").tolist()]   # caller's order

        # host planning (flattening ~1000 small arrays, ~2 ms) can run one or two batches ahead in a helper thread
        # (numpy and the ctypes launches release the GIL). Measured on S batch 1024 it does not pay (5.52 vs 5.42 ms per
        # batch inline: the host is already ahead of the GPU), so it is opt-in for slower hosts / larger batches.
        import queue
        import threading
        plans = queue.Queue(maxsize=2)

        def planner():
            try:
                for b in batches:
                    kw = b if isinstance(b, dict) else {"xs": b}
                    plans.put(self._plan(**kw))
                plans.put(None)
            except BaseException as e:                                 # re-raised in the consumer
                plans.put(e)

        if planner_thread:
            threading.Thread(target=planner, daemon=True).start()
            plan_iter = iter(plans.get, None)
        else:                                                         # plan on the calling thread, batch by batch
            plan_iter = (self._plan(**(b if isinstance(b, dict) else {"xs": b})) for b in batches)
        engine = self.engine()
        n = 0
        for pl in plan_iter:
            if isinstance(pl, BaseException):
                raise pl
            if before_batch is not None:
                before_batch()
            res = engine.run(pl, self.hp.zoneout_rate, self._dropout_rate, self._seed_for_call())   # no host sync
            slot = n % (depth + 1)                                    # depth in flight + the one the caller holds
            n += 1
            need = res.out.shape[0]
            while len(self._host_ring) < depth + 1:
                self._host_ring.append(None)
            if n == 1:                                                # pinning is slow: size every slot on the first batch
                for i in range(depth + 1):
                    b_ = self._host_ring[i]
                    if b_ is None or b_.shape[0] < need or b_.shape[1] != res.out.shape[1]:
                        self._host_ring[i] = torch.empty((need * 5 // 4 + 1, res.out.shape[1]), dtype=torch.float32,
                                                         pin_memory=True)
            buf = self._host_ring[slot]
            if buf.shape[0] < need or buf.shape[1] != res.out.shape[1]:
                buf = self._host_ring[slot] = torch.empty((need * 5 // 4 + 1, res.out.shape[1]), dtype=torch.float32,
                                                          pin_memory=True)
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(dev))
            copied = torch.cuda.Event()
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done)
                buf[:need].copy_(res.out, non_blocking=True)
                res.out.record_stream(copy_stream)
                copied.record(copy_stream)
            pending.append((res, buf, copied))
            if len(pending) >= depth:
                yield finish(pending.pop(0))
        while pending:
            yield finish(pending.pop(0))

    @torch.no_grad()
    def inference(self, x, inference_args=None, spemb=None, dur=None, f0=None, energy=None, utt_id=None,
                  *args, **kwargs):
        """x (T,) -> (L, odim). `inference_args` is ignored, as in the reference (:624-683)."""
        if spemb is not None:
            raise ValueError("speaker embeddings (spk_embed_dim) are out of scope of the B200 path")
        uid = kwargs.pop("dropout_utt_index", 0)
        outs = self.inference_batch([x], None if dur is None else [dur], None if f0 is None else [f0],
                                    None if energy is None else [energy], utt_ids=[uid])
        return outs[0]


class Tacotron2_sa_student(Tacotron2_sa):
    """nets.knowledge_distillation.e2e_tts_tacotron2_sa_kd_student:Tacotron2_sa -- same inference path;
    KD projection tensors are held (when teacher_args is given) only so checkpoints round-trip."""
    _student = True


def from_preset(kind: str, idim: int = 76, odim: int = 80, seed: int | None = 0, device="cuda",
                precision: str = "fp32", kd_keys: bool = False):
    """Build FCL-taco2-S ('S') or -T ('T') from the built-in restatement of conf/*.yaml
    (hparams.PRESETS), optionally random-initialised, on `device`."""
    from .hparams import preset
    hp = preset(kind, idim=idim, odim=odim)
    ns = argparse.Namespace(**{k: v for k, v in hp.to_dict().items() if k not in ("idim", "odim")})
    com = argparse.Namespace(use_fe_condition=True, append_position=True)
    if kind == "S":
        tns = None
        if kd_keys:
            t = preset("T")
            tns = argparse.Namespace(**{k: v for k, v in t.to_dict().items() if k not in ("idim", "odim")})
        m = Tacotron2_sa_student(idim, odim, ns, com, tns, precision=precision)
    else:
        m = Tacotron2_sa(idim, odim, ns, com, precision=precision)
    if seed is not None:
        m.init_random(seed)
    return m.to(device)
