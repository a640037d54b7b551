"""RMSProp (reference: optimisers/RMSProp.py:5-36)."""
from .. import runtime
from .._lib import api
from ._multi import MultiTensorOptimiser, collect_layers


class RMSProp(MultiTensorOptimiser):
    def __init__(self, network, learning_rate, decay_rate, fixed_traversal=False, include_skip_projections=False):
        """fixed_traversal=False is the reference (RMSProp.py:12-15 re-tests the block, so nothing inside a
        ResidualBlock is updated: SURVEY.md F5); True descends as SGDMomentum does; include_skip_projections also
        updates the skip projections (both opt-in)."""
        super().__init__(network, learning_rate)
        self.learnable_layers = collect_layers(network, descend=fixed_traversal or include_skip_projections,
                                               include_skip=include_skip_projections)
        self.decay_rate = decay_rate
        self.grad_cache = {}

    def _second_hyper(self):
        return self.decay_rate

    def update_weights(self):
        """c = d*c + (1-d)*g^2 ; w -= lr*g/sqrt(c + 1e-5) (RMSProp.py:28-36), one launch."""
        def plain(tab, n, max_n):
            api.dk_opt_rmsprop_multi(tab, n, max_n, float(self.learning_rate), float(self.decay_rate),
                                     float(self.grad_scale), self.push_hyper(), runtime.stream())
        self._update(2, plain)
