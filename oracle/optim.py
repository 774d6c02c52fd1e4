"""Row-wise optimizer rules restated in numpy (TEST INFRASTRUCTURE — see oracle/__init__.py).

adam_lazy   paddle.optimizer.Adam(lazy_mode=True) as used by models/rank/deepfm/static_model.py:
            101-103: only rows present in the (merged) SelectedRows gradient are updated; bias
            correction uses the global step.  Formula from Paddle's public Adam documentation:
              m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g*g
              lr_t = lr*sqrt(1-b2^t)/(1-b1^t) ; w -= lr_t * m/(sqrt(v) + eps*sqrt(1-b2^t))
sgd         w -= lr*g  (models/rank/din/dygraph_model.py:64-73)
adagrad_row SparseAdaGradSGDRule (models/rank/slot_dnn/config_online.yaml:57-79,
            doc/online_trainer.md:135-144): ONE g2sum per row;
              w -= lr*g*sqrt(g0/(g0+g2sum)); clamp to bounds; g2sum += mean(g*g)
"""
import numpy as np


def merge_rows(ids: np.ndarray, grads: np.ndarray, padding_idx=None):
    """SelectedRows merge-add: returns (unique ids ascending, summed grads [U,D])."""
    ids = ids.reshape(-1)
    grads = grads.reshape(len(ids), -1).astype(np.float64)
    keep = np.ones(len(ids), bool) if padding_idx is None else ids != padding_idx
    uniq, inv = np.unique(ids[keep], return_inverse=True)
    out = np.zeros((len(uniq), grads.shape[1]))
    np.add.at(out, inv, grads[keep])
    return uniq, out


def adam_lazy(W, m, v, rows, g, lr, beta1, beta2, eps, t):
    W, m, v = W.copy(), m.copy(), v.copy()
    b1p, b2p = beta1 ** t, beta2 ** t
    lr_t = lr * np.sqrt(1 - b2p) / (1 - b1p)
    m[rows] = beta1 * m[rows] + (1 - beta1) * g
    v[rows] = beta2 * v[rows] + (1 - beta2) * g * g
    W[rows] -= lr_t * m[rows] / (np.sqrt(v[rows]) + eps * np.sqrt(1 - b2p))
    return W, m, v


def sgd(W, rows, g, lr):
    W = W.copy()
    W[rows] -= lr * g
    return W


def adagrad_row(W, g2sum, rows, g, lr, g0, lo, hi):
    W, g2sum = W.copy(), g2sum.copy()
    scale = np.sqrt(g0 / (g0 + g2sum[rows]))[:, None]
    W[rows] = np.clip(W[rows] - lr * g * scale, lo, hi)
    g2sum[rows] += (g * g).mean(1)
    return W, g2sum


def clip_then_l2(grads, params, l2_coeffs, clip_norm=None):
    """Dense-gradient preparation order of paddle.optimizer.Optimizer.apply_gradients (public API
    docs): ClipGradByGlobalNorm first (models/rank/dcn_v2/dygraph_model.py:81-88), then each
    parameter's ParamAttr regulariser, L2Decay(c): g += c * w (models/rank/dcn_v2/net.py:166-168)."""
    grads = [np.asarray(g, np.float64) for g in grads]
    if clip_norm is not None:
        norm = np.sqrt(sum(float((g * g).sum()) for g in grads))
        grads = [g * (clip_norm / max(norm, clip_norm)) for g in grads]
    return [g + c * np.asarray(w, np.float64) for g, w, c in zip(grads, params, l2_coeffs)]
