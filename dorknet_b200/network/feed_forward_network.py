"""FeedForwardNetwork: our own sequential container with the reference's interface
(network/feed_forward_network.py:15-139).  The layers also drop into the reference's UNMODIFIED
container (see dorknet_b200.dropin and INTEGRATION.md); this one exists so the package stands alone."""
import json

import numpy as np

from ..array import asarray, asnumpy


class FeedForwardNetwork:
    def __init__(self, name):
        self.name = name
        self.is_on_gpu = False
        self.layers = []
        self.loss_layer = None

    def __repr__(self):
        return "{}: \n".format(self.name) + "".join("\t" + repr(l) + "\n" for l in self.layers)

    def add_layer(self, layer):
        self.layers.append(layer)

    def set_loss_layer(self, loss_layer):
        self.loss_layer = loss_layer

    def to_gpu(self):
        if self.is_on_gpu:
            print("Model already on GPU, ignoring request")
            return
        for layer in self.layers:
            try:
                layer.to_gpu()
            except Exception as e:
                print("Error putting layer {} on GPU, error was: {}".format(layer, e))
                raise
        self.is_on_gpu = True

    def forward(self, X, y_one_hot, test_mode=False, terminal_layer_name=None):
        """feed_forward_network.py:47-62: returns (loss + regularisation terms, scores)."""
        loss = 0
        reg_terms = []
        for layer in self.layers:
            X = layer.forward(X, test_mode=test_mode)
            if layer.layer_name == terminal_layer_name:
                return loss, X
            if not test_mode and hasattr(layer, "regulariser_forward"):
                reg_terms.append(layer.regulariser_forward())
        if self.loss_layer is not None:
            this_loss, X = self.loss_layer.forward(X, y_one_hot, test_mode=test_mode)
            loss += this_loss
            loss += sum(reg_terms)
        return loss, X

    def backward(self):
        """feed_forward_network.py:64-70"""
        if self.loss_layer is None:
            raise ValueError("Network doesn't have a loss, can't run backward pass.")
        upstream_dx = self.loss_layer.backward()
        for layer in self.layers[::-1]:
            upstream_dx = layer.backward(upstream_dx)

    def test(self, data_loader, batch_size, test_set_size):
        """feed_forward_network.py:72-88"""
        correct = 0
        for X_b, y_b, _ in data_loader:
            _, scores = self.forward(asarray(X_b), y_one_hot=None, test_mode=True)
            correct += int(np.sum(np.asarray(y_b) == np.argmax(asnumpy(scores), axis=1)))
        return float(correct) / test_set_size

    def save_layer_structure_to_json(self, fname):
        structure = {"name": self.name}
        for layer in self.layers:
            structure[layer.layer_name] = repr(layer)
        if self.loss_layer is not None:
            structure[self.loss_layer.layer_name] = repr(self.loss_layer)
        with open(fname, "w") as f:
            json.dump(structure, f, indent=4)

    def save_weights_to_h5(self, fname, optimiser=None):
        """feed_forward_network.py:90-95: one group per layer in the reference's HDF5 layout.  `optimiser`
        (extension): also store its velocities / running squared gradients."""
        from ..checkpoint import h5_module, save_optimiser_state
        with h5_module().File(fname, "w") as f:
            for layer in self.layers:
                layer.save_to_h5(f)
            if self.loss_layer is not None:
                self.loss_layer.save_to_h5(f)
            if optimiser is not None:
                save_optimiser_state(optimiser, f)

    def load_network_from_json_and_h5(self, json_fname, h5_fname, optimiser=None):
        """feed_forward_network.py:106-139: layer order and names from the JSON structure file, types and
        parameters from the HDF5 file."""
        from ..checkpoint import h5_module, load_optimiser_state, make_layer
        with open(json_fname, "r") as f:
            structure = json.load(f)
        with h5_module().File(h5_fname, "r") as f:
            self.name = structure.pop("name")
            for layer_name in structure.keys():
                l_type = f[layer_name + "/layer_info"].attrs["type"]
                layer = make_layer(l_type, layer_name)
                layer.load_from_h5(f)
                if type(layer).__name__ == "SoftmaxWithCrossEntropy":
                    self.loss_layer = layer
                else:
                    self.layers.append(layer)
            if optimiser is not None:
                load_optimiser_state(optimiser, f)
        self.is_on_gpu = False
