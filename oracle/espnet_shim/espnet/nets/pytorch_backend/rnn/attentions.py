"""Stub: the reference imports AttForwardTA (nets/modules/decoder_sa.py:11) but never uses it."""


class AttForwardTA:
    pass
