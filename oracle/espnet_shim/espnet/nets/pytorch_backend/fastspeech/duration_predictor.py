"""Stub of espnet's FastSpeech DurationPredictor (espnet 0.8 era, un-vendored;
reference import: nets/teacher_training/e2e_tts_tacotron2_sa.py:18-19, ctor call
:406-412, inference call :645). Restated from the published algorithm: the
same conv stack as the reference's variance_predictor.py:48-66 (which says it
was derived from this class), Linear(n_chans, 1), and at inference
clamp(round(exp(x) - offset), min=0).long(). PARITY UNPINNED: no reference
test or golden vector pins this class."""
import torch

from espnet.nets.pytorch_backend.transformer.layer_norm import LayerNorm


class DurationPredictor(torch.nn.Module):
    def __init__(self, idim, n_layers=2, n_chans=384, kernel_size=3,
                 dropout_rate=0.1, offset=1.0):
        super().__init__()
        self.offset = offset
        self.conv = torch.nn.ModuleList()
        for i in range(n_layers):
            cin = idim if i == 0 else n_chans
            self.conv.append(torch.nn.Sequential(
                torch.nn.Conv1d(cin, n_chans, kernel_size, stride=1,
                                padding=(kernel_size - 1) // 2),
                torch.nn.ReLU(),
                LayerNorm(n_chans, dim=1),
                torch.nn.Dropout(dropout_rate),
            ))
        self.linear = torch.nn.Linear(n_chans, 1)

    def _run(self, xs, x_masks, is_inference):
        xs = xs.transpose(1, -1)
        for f in self.conv:
            xs = f(xs)
        xs = self.linear(xs.transpose(1, -1)).squeeze(-1)
        if is_inference:
            xs = torch.clamp(torch.round(xs.exp() - self.offset), min=0).long()
        if x_masks is not None:
            xs = xs.masked_fill(x_masks, 0.0)
        return xs

    def forward(self, xs, x_masks=None):
        return self._run(xs, x_masks, False)

    def inference(self, xs, x_masks=None):
        return self._run(xs, x_masks, True)


class DurationPredictorLoss(torch.nn.Module):
    def __init__(self, offset=1.0, reduction="mean"):
        super().__init__()
        self.criterion = torch.nn.MSELoss(reduction=reduction)
        self.offset = offset

    def forward(self, outputs, targets):
        return self.criterion(outputs, torch.log(targets.float() + self.offset))
