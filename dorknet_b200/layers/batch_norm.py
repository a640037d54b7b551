"""BatchNormLayer (reference: layers/batch_norm.py:9-232, layers/batch_norm_stats_cy.pyx)."""
import numpy as np

from .layer import Layer, api, runtime, asarray, empty
from ..array import FoldedBNGrad, LazyBNOutput, LazyDWOutput, LazyStridedGrad, ZeroSumGrad


class _RunningStats(dict):
    """non_learned_params of a BatchNormLayer: reading an entry first runs a still-deferred forward kernel, so the
    running mean / std are always the ones the reference would hold at this point (batch_norm.py:76-89)."""

    def __init__(self, owner, *a, **kw):
        super().__init__(*a, **kw)
        self._owner = owner

    def __getitem__(self, k):
        self._owner._flush()
        return super().__getitem__(k)

    def get(self, k, default=None):
        self._owner._flush()
        return super().get(k, default)

    def items(self):
        self._owner._flush()
        return super().items()

    def values(self):
        self._owner._flush()
        return super().values()


class BatchNormLayer(Layer):
    """https://arxiv.org/pdf/1502.03167.pdf -- same constructor as batch_norm.py:13-14."""
    _h5_attrs = ("input_dimension", "run_momentum", "incoming_chans", "eps")
    _h5_params = ("gamma", "beta")  # layers/batch_norm.py:176-232
    _h5_state = ("running_mean", "running_std")


    def __init__(self, layer_name, input_dimension=4, incoming_chans=None, run_momentum=0.95, is_on_gpu=True):
        super().__init__(layer_name)
        self.eps = 1e-5
        self.input_dimension = input_dimension
        self._pending = None
        self.non_learned_params = _RunningStats(self, {"running_mean": None, "running_std": None})
        self.run_momentum = run_momentum
        if self.input_dimension not in {2, 4}:
            raise ValueError("BatchNorm input_dimension should have length 2 or 4...")
        self.av_axis = (0, 2, 3) if self.input_dimension == 4 else 0
        self.incoming_chans = incoming_chans
        if incoming_chans is not None:
            gamma = np.ones(incoming_chans, dtype=np.float32)
            beta = np.zeros(incoming_chans, dtype=np.float32)
            if self.input_dimension == 4:
                gamma = gamma[np.newaxis, :, np.newaxis, np.newaxis]
                beta = beta[np.newaxis, :, np.newaxis, np.newaxis]
            self.learned_params = {"gamma": gamma, "beta": beta}
            self.grads = {"gamma": np.zeros_like(gamma), "beta": np.zeros_like(beta)}
        else:
            self.learned_params = {}
            self.grads = {}
        self._x = None
        self._relu_fused = False
        self._pending = None
        self._relu_claimed = False
        self._relu_done = False
        self._relu_out = None
        self.defer_apply = True  # return a lazy output so that a following ReLu can fuse (False: run the kernel now)

    def fused_relu_apply(self, y):
        """y = relu(batchnorm(x)) in the SAME kernel as the statistics (called by the ReLu that consumes our lazy
        output); the mask is re-derived from x in backward, so this layer's backward must start with it."""
        pending, self._pending = self._pending, None
        if pending is None:
            if self._relu_done:  # _flush already ran the fused pass into the ReLu's buffer
                return
            raise RuntimeError("BatchNormLayer {}: output already materialised".format(self.layer_name))
        pending(y, 1)
        self._relu_fused = True

    def expect_fused_relu(self, relu_out):
        """A ReLu took our deferred output (and will write into `relu_out`): whatever ends up running the normalisation
        pass, it is the fused one and backward masks dY."""
        self._relu_claimed = True
        self._relu_out = relu_out

    def fused_relu_apply_strided(self, out, stride):
        """out[n,c,oh,ow] = relu(batchnorm(x))[n,c,oh*s,ow*s] only (what a stride-s PointwiseConvLayer reads): the
        statistics pass, then a normalisation pass over one pixel in s*s.  The full-size output is not produced."""
        pending, self._pending = self._pending, None
        if pending is None:
            raise RuntimeError("BatchNormLayer {}: output already materialised".format(self.layer_name))
        pending(None, 0)  # statistics, running mean / std, saved scale / shift
        self._relu_fused = True
        N, C, HW = self._dims(self.input_shape)
        H, W = self.input_shape[2], self.input_shape[3]
        base = self._bufs["saved"].ptr
        api.dk_bn_apply_strided(self._x.ptr, out.ptr, base + 8 * C, base + 12 * C, 1, N, C, H, W, int(stride),
                                runtime.stream())

    def can_fold(self, policy=True):
        """May the PointwiseConvLayer that received our deferred output fold the normalisation into its GEMMs (bn_fold.cu)?
        policy True: only when the statistics come for free from the producing depthwise kernel (a still-unlaunched
        LazyDWOutput), which is what makes the fold a win; "always": whenever the output is still deferred."""
        if self._pending is None or self._relu_claimed or self.input_dimension != 4:
            return False
        if policy == "always":
            return True
        return bool(policy) and isinstance(self._x, LazyDWOutput) and not self._x.is_materialised

    def fold_statistics(self):
        """A PointwiseConvLayer took our deferred output and folds the normalisation into its GEMMs: only the statistics
        are produced (batch mean / std, running statistics, saved scale / shift) -- by the depthwise forward kernel itself
        when our input is a still-unlaunched LazyDWOutput, else by the statistics pass.  Returns (x, saved base pointer,
        has_residual); backward() is served by that layer (array.FoldedBNGrad)."""
        pending, self._pending = self._pending, None
        if pending is None:
            raise RuntimeError("BatchNormLayer {}: output already materialised".format(self.layer_name))
        has_resid = self._fold_stats()
        self._relu_fused = False
        return self._x, self._bufs["saved"].ptr, has_resid

    def apply_saved(self, y, relu):
        """y = relu?(x*scale + shift) from the statistics of the last training forward (a late reader of an output
        whose full-size pass was skipped)."""
        N, C, HW = self._dims(self.input_shape)
        base = self._bufs["saved"].ptr
        api.dk_bn_apply(self._x.ptr, y.ptr, base + 8 * C, base + 12 * C, int(relu), N, C, HW, runtime.stream())

    def fused_add_relu_apply(self, y, skip):
        """y = relu(batchnorm(x) + skip) in the SAME kernel as the statistics: the join of a ResidualBlock whose branch
        ends in this layer (residual_block.py:75).  The mask of that ReLU depends on `skip`, so backward stays unfused:
        the block's ReLu masks dY with its own output and this layer's backward sees a plain gradient."""
        pending, self._pending = self._pending, None
        if pending is None:
            raise RuntimeError("BatchNormLayer {}: output already materialised".format(self.layer_name))
        pending(y, 1, skip)
        self._relu_fused = False

    def _flush(self):
        """Run a still-deferred training forward (somebody needs the statistics before any consumer used y)."""
        if self._pending is not None:
            pending, self._pending = self._pending, None
            if self._relu_claimed:
                # a ReLu holds the (still unread) output: run the fused pass into its buffer now
                pending(self._relu_out, 1)
                self._relu_fused = True
                self._relu_done = True
            else:
                pending(self._bufs["y"], 0)

    def __repr__(self):
        return "BatchNormLayer({}, input_dimension={}, incoming_chans={}, run_momentum={})".format(
            self.layer_name, self.input_dimension, self.incoming_chans, self.run_momentum)

    @staticmethod
    def _dims(shape):
        if len(shape) == 4:
            return shape[0], shape[1], shape[2] * shape[3]
        return shape[0], shape[1], 1

    def _stat_shape(self, C):
        return (1, C, 1, 1) if self.input_dimension == 4 else (C,)

    def forward(self, X, test_mode=False, use_express=False):
        """batch_norm.py:54-115.  Train: batch mean / biased var (one pass, Chan-merged), std =
        sqrt(var+eps), running mean / running STD EMA, y = gamma*x_hat + beta.  Instead of caching
        X_demean and X_hat (two extra activation-sized writes) only mean/invstd/scale/shift per
        channel are kept and x_hat is recomputed from the saved input in backward."""
        self._ensure_gpu()
        X = asarray(X)
        if X.ndim != self.input_dimension:
            raise ValueError("BatchNormLayer {} expects {}-D input, got shape {}".format(
                self.layer_name, self.input_dimension, X.shape))
        N, C, HW = self._dims(X.shape)
        self.input_shape = X.shape
        gamma, beta = self._param("gamma"), self._param("beta")
        y = self._buf("y", X.shape)
        st = runtime.stream()
        if not test_mode:
            first = self.non_learned_params["running_mean"] is None
            if first:
                self.non_learned_params["running_mean"] = empty(self._stat_shape(C))
                self.non_learned_params["running_std"] = empty(self._stat_shape(C))
            rm, rs = self.non_learned_params["running_mean"], self.non_learned_params["running_std"]
            sv = self._buf("saved", (5, C))  # mean, invstd, scale, shift, TF32 truncation residual (folded path only)
            ws, wsn = self._zeroed_ws(api.dk_bn_ws_bytes(C))
            base = sv.ptr
            self._flush()
            self._x = X
            self._relu_fused = False
            self._relu_claimed = False
            self._relu_done = False
            mom, eps = float(self.run_momentum), float(self.eps)

            def run(out, relu, add=None):
                # statistics + running mean/std + normalisation (+ReLU) in one call: a cluster kernel that keeps the
                # channel in shared memory between the two passes (bn_fused.cu), or the split kernels
                out_ptr = out.ptr if out is not None else None
                if add is not None:
                    api.dk_bn_fwd_train_add(X.ptr, add.ptr, out_ptr, gamma.ptr, beta.ptr, rm.ptr, rs.ptr, int(first), mom,
                                            eps, base, base + 4 * C, base + 8 * C, base + 12 * C, relu, N, C, HW, ws, wsn,
                                            runtime.stream())
                    return
                api.dk_bn_fwd_train(X.ptr, out_ptr, gamma.ptr, beta.ptr, rm.ptr, rs.ptr, int(first), mom, eps,
                                    base, base + 4 * C, base + 8 * C, base + 12 * C, relu, N, C, HW, ws, wsn,
                                    runtime.stream())
            def fold_stats():
                if isinstance(X, LazyDWOutput) and not X.is_materialised:
                    X.layer.forward_with_bn_statistics(X, gamma.ptr, beta.ptr, rm.ptr, rs.ptr, first, mom, eps, base)
                    return True
                run(None, 0)
                return False
            self._fold_stats = fold_stats
            if self.defer_apply:
                # nothing is launched until the consumer is known: a ReLu asks for the fused variant, anybody else
                # (or a reader of the running statistics) gets the plain one
                self._pending = run

                def apply():
                    pending, self._pending = self._pending, None
                    if pending is not None:
                        pending(y, 0)
                return LazyBNOutput(y, apply, self)
            run(y, 0)
        else:
            rm, rs = self.non_learned_params["running_mean"], self.non_learned_params["running_std"]
            if rm is None:
                raise ValueError("BatchNormLayer {}: test_mode forward before any training batch".format(self.layer_name))
            rm, rs = asarray(rm), asarray(rs)
            api.dk_bn_fwd_infer(X.ptr, y.ptr, gamma.ptr, beta.ptr, rm.ptr, rs.ptr, 0, N, C, HW, st)
        return y

    def backward(self, upstream_dx):
        """batch_norm.py:118-174: grads["gamma"], grads["beta"] and dx in two passes over (dY, X)."""
        if isinstance(upstream_dx, FoldedBNGrad) and upstream_dx.bn is self and not upstream_dx.is_materialised:
            return upstream_dx.dx  # dgamma / dbeta / dx were produced by the pointwise layer this BatchNorm is folded into
        dY = asarray(upstream_dx)
        self._flush()
        N, C, HW = self._dims(self.input_shape)
        sv = self._bufs["saved"]
        base = sv.ptr
        dx = self._buf("dx", self.input_shape)
        dx = ZeroSumGrad(dx.t, dx.shape)
        dg, db = self._grad("gamma"), self._grad("beta")
        ws, wsn = self._zeroed_ws(api.dk_bn_ws_bytes(C))
        if (isinstance(dY, LazyStridedGrad) and not dY.is_materialised and dY.stride == 2 and self.input_dimension == 4
                and tuple(dY.shape) == tuple(self.input_shape) and self.input_shape[3] % 8 == 0
                and self.input_shape[2] % 2 == 0):
            # the gradient of a stride-2 pointwise convolution, taken in its compact form: the 3/4 zeros are neither
            # written by the dgrad GEMM nor read here
            H, W = self.input_shape[2], self.input_shape[3]
            dys = dY.compact()
            api.dk_bn_bwd_strided(dys.ptr, self._x.ptr, self._param("gamma").ptr, base, base + 4 * C, base + 8 * C,
                                  base + 12 * C, dx.ptr, dg.ptr, db.ptr, 1 if self._relu_fused else 0, N, C, H, W, 2,
                                  ws, wsn, runtime.stream())
            return dx
        api.dk_bn_bwd(dY.ptr, self._x.ptr, self._param("gamma").ptr, base, base + 4 * C, base + 8 * C, base + 12 * C,
                      dx.ptr, dg.ptr, db.ptr, 1 if self._relu_fused else 0, N, C, HW, ws, wsn, runtime.stream())
        return dx

    def backward_join(self, upstream_dx, block_out, joined_dx):
        """ResidualBlock whose branch ends here: joined_dx = upstream_dx * (block_out > 0) (the block's ReLU backward,
        activations.py:44-47), kept for the skip path, and this layer's backward of it -- in one kernel."""
        dY = asarray(upstream_dx)
        self._flush()
        N, C, HW = self._dims(self.input_shape)
        base = self._bufs["saved"].ptr
        dx = self._buf("dx", self.input_shape)
        dx = ZeroSumGrad(dx.t, dx.shape)
        dg, db = self._grad("gamma"), self._grad("beta")
        ws, wsn = self._zeroed_ws(api.dk_bn_ws_bytes(C))
        api.dk_bn_bwd_join(dY.ptr, block_out.ptr, self._x.ptr, self._param("gamma").ptr, base, base + 4 * C, base + 8 * C,
                           base + 12 * C, dx.ptr, joined_dx.ptr, dg.ptr, db.ptr, N, C, HW, ws, wsn, runtime.stream())
        return dx

    # the reference exposes these pieces separately (batch_norm.py:124-174)
    @property
    def std(self):
        """sqrt(var + eps) of the last training batch, shape (1,C,1,1) / (C,) (batch_norm.py:69-72)."""
        self._flush()
        sv = self._bufs["saved"].get()
        return (1.0 / sv[1]).reshape(self._stat_shape(sv.shape[1])).astype(np.float32)
