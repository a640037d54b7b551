"""Stub of `h5py` (TEST INFRASTRUCTURE ONLY): imported by
/root/reference/network/feed_forward_network.py:3; checkpoints are out of scope."""


class File:
    def __init__(self, *a, **k):
        raise NotImplementedError("h5py is not installed in this image")
