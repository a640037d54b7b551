"""SoftmaxWithCrossEntropy (reference: layers/losses.py:5-41)."""
from .layer import Layer, api, runtime, asarray, DeviceScalar
from ..array import alloc_scalar_slot, current_epoch, seal_epoch
from ..regularisers.l2 import flush_pending


class SoftmaxWithCrossEntropy(Layer):

    def __init__(self, layer_name):
        super().__init__(layer_name)

    def __repr__(self):
        return "SoftmaxWithCrossEntropy({})".format(self.layer_name)

    def forward(self, X, y_one_hot=None, test_mode=False):
        """Returns (loss, softmax scores) like losses.py:13-27 (softmax without max subtraction;
        loss = mean(-log(sum_j p_j*y_j)), so soft / mixup labels behave as in the reference).
        `loss` is a DeviceScalar: adding the regularisation terms to it costs no kernel and float()
        is the only synchronisation point."""
        self._ensure_gpu()
        flush_pending()  # every layer's l2 term of this step, one launch
        X = asarray(X)
        B, K = X.shape
        p = self._buf("p", (B, K))
        if test_mode:
            api.dk_softmax_xent_fwd(X.ptr, None, p.ptr, None, B, K, runtime.stream())
            return 0, p
        self.y_one_hot = asarray(y_one_hot)
        if self.y_one_hot.shape != (B, K):
            raise ValueError("SoftmaxWithCrossEntropy: labels {} do not match scores {}".format(
                self.y_one_hot.shape, (B, K)))
        if getattr(self, "_loss_slot", None) is None:
            self._loss_slot = alloc_scalar_slot()
        loss, idx = self._loss_slot
        api.dk_softmax_xent_fwd(X.ptr, self.y_one_hot.ptr, p.ptr, loss.ptr, B, K, runtime.stream())
        self.downstream_x = p
        # the step's scalar results (this loss + the l2 terms flushed above) are complete: snapshot them, so that the
        # value the caller holds is the value of THIS step (array.DeviceScalar).  Under stream capture nothing can be
        # snapshotted (GraphedTrainStep seals after every replay instead).
        import torch
        ep = current_epoch() if torch.cuda.is_current_stream_capturing() else seal_epoch()
        return DeviceScalar([(idx, 1.0)], 0.0, ep), p

    def backward(self, upstream_dx=None):
        """(p - y)/B (losses.py:29-34); upstream_dx is not used."""
        B, K = self.downstream_x.shape
        dx = self._buf("dx", (B, K))
        api.dk_softmax_xent_bwd(self.downstream_x.ptr, self.y_one_hot.ptr, dx.ptr, B, K, runtime.stream())
        return dx
