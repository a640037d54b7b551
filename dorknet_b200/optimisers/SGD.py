"""SGD (reference: optimisers/SGD.py:1-24)."""
from .. import runtime
from .._lib import api
from ._multi import MultiTensorOptimiser, collect_layers


class SGD(MultiTensorOptimiser):
    needs_state = False

    def __init__(self, network, learning_rate):
        super().__init__(network, learning_rate)
        self.learnable_layers = collect_layers(network, descend=False)

    def update_weights(self):
        """w += -lr * g for every tensor of the update set, one launch."""
        def plain(tab, n, max_n):
            api.dk_opt_sgd_multi(tab, n, max_n, float(self.learning_rate), float(self.grad_scale), self.push_hyper(),
                                 runtime.stream())
        self._update(0, plain)
