"""`h5py` for the reference modules loaded from oracle/_ref (TEST INFRASTRUCTURE ONLY).

h5py is not in this image.  The reference's save_weights_to_h5 / load_network_from_json_and_h5
(/root/reference/network/feed_forward_network.py:90-139) are exercised in the checkpoint-compatibility tests through
the pure-Python HDF5 subset of the product, loaded here by file path so that importing this stub never imports the
`dorknet_b200` package (whose layer modules share names with the reference's)."""
import importlib.util
import os
import sys

_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "dorknet_b200", "minih5.py")
_spec = importlib.util.spec_from_file_location("_dk_minih5_for_reference", _path)
_mod = importlib.util.module_from_spec(_spec)
sys.modules[_spec.name] = _mod
_spec.loader.exec_module(_mod)

File = _mod.File
Empty = _mod.Empty
