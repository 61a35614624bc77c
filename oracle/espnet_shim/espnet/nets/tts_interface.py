"""Stub of espnet.nets.tts_interface.TTSInterface (reference import:
nets/teacher_training/e2e_tts_tacotron2_sa.py:15). Inference needs only a
`.reporter` attribute whose `.report()` is a no-op."""


class _NullReporter:
    def report(self, *a, **k):
        pass


class TTSInterface:
    @staticmethod
    def add_arguments(parser):
        return parser

    def __init__(self):
        self.reporter = _NullReporter()

    def forward(self, *a, **k):
        raise NotImplementedError

    def inference(self, *a, **k):
        raise NotImplementedError

    def load_pretrained_model(self, model_path):
        raise NotImplementedError("stub")

    @property
    def attention_plot_class(self):
        return None

    @property
    def base_plot_keys(self):
        return []
