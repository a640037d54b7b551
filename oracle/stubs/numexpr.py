"""Stub of `numexpr` (TEST INFRASTRUCTURE ONLY): the reference calls
ne.set_vml_accuracy_mode at import time (/root/reference/layers/batch_norm.py:6);
ne.evaluate is only reached with use_express=True, which no caller passes."""


def set_vml_accuracy_mode(mode):
    return None


def evaluate(*a, **k):
    raise NotImplementedError("numexpr stub: use_express path is dead code in the reference")
