"""SGDMomentum (reference: optimisers/SGDMomentum.py:4-38)."""
from .. import runtime
from .._lib import api
from ._multi import MultiTensorOptimiser, collect_layers


class SGDMomentum(MultiTensorOptimiser):
    def __init__(self, network, learning_rate, momentum, include_skip_projections=False):
        super().__init__(network, learning_rate)
        self.learnable_layers = collect_layers(network, descend=True, include_skip=include_skip_projections)
        self.momentum = momentum
        self.grad_cache = {}  # layer -> {param name -> DeviceArray velocity}; allocated on first update

    def _second_hyper(self):
        return self.momentum

    def update_weights(self):
        """v = -lr*g + momentum*v ; w += v (SGDMomentum.py:31-39), one launch."""
        def plain(tab, n, max_n):
            api.dk_opt_sgdm_multi(tab, n, max_n, float(self.learning_rate), float(self.momentum),
                                  float(self.grad_scale), self.push_hyper(), runtime.stream())
        self._update(1, plain)
