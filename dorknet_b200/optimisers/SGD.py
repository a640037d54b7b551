"""SGD (reference: optimisers/SGD.py:1-24)."""
from .. import runtime
from .._lib import api
from ._multi import MultiTensorOptimiser, collect_layers


class SGD(MultiTensorOptimiser):
    needs_state = False

    def __init__(self, network, learning_rate, fixed_traversal=False, include_skip_projections=False):
        """fixed_traversal=False is the reference (SGD.py:8-11 never updates anything inside a ResidualBlock:
        SURVEY.md F5); True descends into blocks as SGDMomentum does; include_skip_projections additionally updates
        the skip projections (both opt-in)."""
        super().__init__(network, learning_rate)
        self.learnable_layers = collect_layers(network, descend=fixed_traversal or include_skip_projections,
                                               include_skip=include_skip_projections)

    def update_weights(self):
        """w += -lr * g for every tensor of the update set, one launch."""
        def plain(tab, n, max_n):
            api.dk_opt_sgd_multi(tab, n, max_n, float(self.learning_rate), float(self.grad_scale), self.push_hyper(),
                                 runtime.stream())
        self._update(0, plain)
