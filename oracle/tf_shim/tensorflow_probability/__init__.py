"""TEST INFRASTRUCTURE ONLY - stand-in for the two tensorflow_probability symbols gpbasics imports
(Optimizer/Fitter.py:10,104-105,152-158).  `VariationalSGD` here is a plain SGD step at TFP's burn-in learning rate;
the optimiser is a caller of the hot path and its update is not a parity target.  The gradients
it saw are kept in `last_grads` so that the golden generator can record exactly what the reference's
`sgd_opt.minimize(opt, hp)` differentiated."""
import types

import tensorflow as tf
import torch


class VariationalSGD:
    def __init__(self, batch_size, total_num_examples, max_learning_rate=1.0, preconditioner_decay_rate=0.95, burnin=25,
                 burnin_max_learning_rate=1e-6, use_single_learning_rate=False, name=None):
        self.lr = burnin_max_learning_rate
        self.last_grads = None
        self.last_loss = None

    def minimize(self, loss, var_list, tape=None):
        value = loss() if callable(loss) else loss
        grads = torch.autograd.grad(value._t.sum(), [v._t for v in var_list], allow_unused=True)
        self.last_loss = value
        self.last_grads = [None if g is None else tf.Tensor(g.detach().clone()) for g in grads]
        self.apply_gradients([(g, v) for g, v in zip(self.last_grads, var_list)])

    def apply_gradients(self, grads_and_vars):
        for g, v in grads_and_vars:
            if g is None:
                continue
            with torch.no_grad():
                v._t.sub_(self.lr * tf._raw(g, v._t.dtype).reshape(v._t.shape))


def _value_and_gradient(f, xs, **kw):
    single = not isinstance(xs, (list, tuple))
    vs = [xs] if single else list(xs)
    ts = [tf.Variable(v) for v in vs]
    y = f(*ts)
    gs = torch.autograd.grad(y._t.sum(), [t._t for t in ts], allow_unused=True)
    gs = [None if g is None else tf.Tensor(g) for g in gs]
    return y, (gs[0] if single else gs)


optimizer = types.SimpleNamespace(VariationalSGD=VariationalSGD)
math = types.SimpleNamespace(value_and_gradient=_value_and_gradient)
