"""Registered model families: analytic ``grad_log_post`` (TEST INFRASTRUCTURE).

Restates ``sgmcmcjax/util.py:9-50``: ``log_post(theta, batch) = logprior(theta)
+ Ndata * mean_i loglik(theta, batch_i)`` (util.py:43-44), differentiated w.r.t.
every parameter leaf (util.py:46-49).  With B rows in the batch the gradient is
``grad logprior + (Ndata / B) * sum_i grad loglik_i``.

Parameters are lists of numpy leaves in jax pytree leaf order (lists/tuples in
order), see SURVEY.md A.5.  Every family exposes

    grad_log_post(leaves, batch_arrays, n_data, dtype) -> list of leaf gradients

The per-datum log-likelihoods / priors restated here:

* gaussian_mean  README.md:22-26 == tests/models.py:6-11 (one leaf) and
                 tests/models.py:15-24 (two leaves), generalised to per-leaf
                 coefficients: loglik = -sum_l a_l |x - th_l|^2,
                 logprior = -sum_l c_l |th_l|^2.
* logistic_regression  sgmcmcjax/models/logistic_regression.py:65-72.
* bayesian_mlp   sgmcmcjax/models/bayesian_NN/NN_model.py:31-58 (softmax hidden
                 activation, log_softmax output, N(0,1) prior).
* pmf            NOT in the reference; defined in SURVEY.md section 8 row a15.
"""
from __future__ import annotations

import numpy as np


class GaussianMean:
    """loglik(theta, x) = -sum_l a_l |x - theta_l|^2 ; logprior = -sum_l c_l |theta_l|^2."""

    name = "gaussian_mean"

    def __init__(self, lik_coef=(0.5,), prior_coef=(0.005,)):
        self.a = tuple(float(v) for v in lik_coef)
        self.c = tuple(float(v) for v in prior_coef)

    def grad_log_post(self, leaves, batch, n_data, dtype=np.float32):
        (x,) = batch
        x = np.asarray(x, dtype=dtype)
        scale = dtype(n_data) / dtype(x.shape[0])
        out = []
        for th, a, c in zip(leaves, self.a, self.c):
            th = np.asarray(th, dtype=dtype)
            lik = (dtype(2.0 * a) * (x - th[None, :])).sum(axis=0, dtype=dtype)
            out.append((scale * lik - dtype(2.0 * c) * th).astype(dtype))
        return out


def _gauss_log_post(self, leaves, batch, n_data, dtype=np.float64):
    """util.py:43-44: logprior + Ndata * mean_i loglik_i (value), evaluated in `dtype`."""
    (x,) = batch
    x = np.asarray(x, dtype=dtype)
    lik = dtype(0.0)
    pri = dtype(0.0)
    for th, a, c in zip(leaves, self.a, self.c):
        th = np.asarray(th, dtype=dtype)
        lik = lik - dtype(a) * ((x - th[None, :]) ** 2).sum(dtype=dtype)
        pri = pri - dtype(c) * (th * th).sum(dtype=dtype)
    return dtype(pri + dtype(n_data) * lik / dtype(x.shape[0]))


GaussianMean.log_post = _gauss_log_post


def _sigmoid(z):
    out = np.empty_like(z)
    pos = z >= 0
    out[pos] = 1.0 / (1.0 + np.exp(-z[pos]))
    e = np.exp(z[~pos])
    out[~pos] = e / (1.0 + e)
    return out


class LogisticRegression:
    """loglik = -logsumexp([0, (1-2y) theta.x]) ; logprior = -(0.5/10) theta.theta
    (models/logistic_regression.py:65-72).  d loglik / d theta = (y - sigmoid(theta.x)) x."""

    name = "logistic_regression"

    def __init__(self, prior_precision=0.1):
        self.prior_precision = float(prior_precision)

    def grad_log_post(self, leaves, batch, n_data, dtype=np.float32):
        x, y = batch
        (th,) = leaves
        x = np.asarray(x, dtype=dtype)
        th = np.asarray(th, dtype=dtype)
        z = (x @ th).astype(dtype)
        r = (np.asarray(y, dtype=dtype) - _sigmoid(z)).astype(dtype)
        scale = dtype(n_data) / dtype(x.shape[0])
        return [(scale * (r @ x).astype(dtype) - dtype(self.prior_precision) * th).astype(dtype)]


def _logreg_log_post(self, leaves, batch, n_data, dtype=np.float64):
    """logprior + Ndata * mean_i loglik_i with loglik = -logsumexp([0, (1-2y) theta.x]) (logistic_regression.py:65-72)."""
    x, y = batch
    (th,) = leaves
    x = np.asarray(x, dtype=dtype)
    th = np.asarray(th, dtype=dtype)
    z = (dtype(1.0) - dtype(2.0) * np.asarray(y, dtype=dtype)) * (x @ th)
    lik = -(np.maximum(z, 0) + np.log1p(np.exp(-np.abs(z)))).sum(dtype=dtype)
    return dtype(-dtype(0.5 * self.prior_precision) * (th * th).sum(dtype=dtype) + dtype(n_data) * lik / dtype(x.shape[0]))


LogisticRegression.log_post = _logreg_log_post


def _softmax(z):
    z = z - z.max(axis=-1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=-1, keepdims=True)


class BayesianMLP:
    """Fully connected net, leaves [W1[n1,n0], b1[n1], W2[n2,n1], b2[n2], ...];
    hidden activation softmax, output log_softmax (NN_model.py:31-41);
    loglik = sum_k y_k log_softmax(z)_k (NN_model.py:48-50); prior N(0,1) on every
    entry (NN_model.py:53-58) so grad logprior = -w."""

    name = "bayesian_mlp"

    def grad_log_post(self, leaves, batch, n_data, dtype=np.float32):
        x, y = batch
        a = np.asarray(x, dtype=dtype)
        y = np.asarray(y, dtype=dtype)
        ws = [np.asarray(w, dtype=dtype) for w in leaves[0::2]]
        bs = [np.asarray(b, dtype=dtype) for b in leaves[1::2]]
        acts = [a]
        for w, b in zip(ws[:-1], bs[:-1]):
            a = _softmax(a @ w.T + b).astype(dtype)
            acts.append(a)
        logits = acts[-1] @ ws[-1].T + bs[-1]
        dz = (y - y.sum(axis=-1, keepdims=True) * _softmax(logits)).astype(dtype)
        scale = dtype(n_data) / dtype(y.shape[0])
        grads = [None] * len(leaves)
        for layer in range(len(ws) - 1, -1, -1):
            a_in = acts[layer]
            grads[2 * layer] = (scale * (dz.T @ a_in) - ws[layer]).astype(dtype)
            grads[2 * layer + 1] = (scale * dz.sum(axis=0) - bs[layer]).astype(dtype)
            if layer:
                da = dz @ ws[layer]
                dz = (a_in * (da - (a_in * da).sum(axis=-1, keepdims=True))).astype(dtype)
        return grads


def _mlp_log_post(self, leaves, batch, n_data, dtype=np.float64):
    """logprior + Ndata * mean_i loglik_i (util.py:43-44); loglik = sum_k y_k log_softmax(z)_k (NN_model.py:48-50),
    logprior = sum of N(0, 1) log-densities of every entry (NN_model.py:53-58)."""
    x, y = batch
    a = np.asarray(x, dtype=dtype)
    y = np.asarray(y, dtype=dtype)
    ws = [np.asarray(w, dtype=dtype) for w in leaves[0::2]]
    bs = [np.asarray(b, dtype=dtype) for b in leaves[1::2]]
    for w, b in zip(ws[:-1], bs[:-1]):
        a = _softmax(a @ w.T + b)
    z = a @ ws[-1].T + bs[-1]
    zmax = z.max(axis=-1, keepdims=True)
    logsm = z - zmax - np.log(np.exp(z - zmax).sum(axis=-1, keepdims=True))
    lik = (y * logsm).sum(dtype=dtype)
    n_par = sum(l.size for l in leaves)
    pri = -0.5 * sum((np.asarray(l, dtype=dtype) ** 2).sum(dtype=dtype) for l in leaves) - 0.5 * np.log(2 * np.pi) * n_par
    return dtype(pri + dtype(n_data) * lik / dtype(y.shape[0]))


BayesianMLP.log_post = _mlp_log_post


class PMF:
    """Probabilistic matrix factorisation (ours, not the reference's).
    leaves (U[n_users,d], V[n_items,d]); data (ij int32[N,2], r f32[N]);
    loglik = -0.5*alpha*(r - U[i].V[j])^2 ; logprior = -0.5*lam*(|U|^2+|V|^2)."""

    name = "pmf"

    def __init__(self, alpha=2.0, lam=1.0):
        self.alpha = float(alpha)
        self.lam = float(lam)

    def grad_log_post(self, leaves, batch, n_data, dtype=np.float32):
        ij, r = batch
        u, v = (np.asarray(l, dtype=dtype) for l in leaves)
        ii = np.asarray(ij)[:, 0]
        jj = np.asarray(ij)[:, 1]
        ui, vj = u[ii], v[jj]
        e = (dtype(self.alpha) * (np.asarray(r, dtype=dtype) - (ui * vj).sum(axis=1))).astype(dtype)
        scale = dtype(n_data) / dtype(len(ii))
        gu = np.zeros_like(u)
        gv = np.zeros_like(v)
        np.add.at(gu, ii, e[:, None] * vj)
        np.add.at(gv, jj, e[:, None] * ui)
        return [(scale * gu - dtype(self.lam) * u).astype(dtype),
                (scale * gv - dtype(self.lam) * v).astype(dtype)]


def _pmf_log_post(self, leaves, batch, n_data, dtype=np.float64):
    """logprior + Ndata * mean_i loglik_i with loglik = -alpha/2 (r - U_i.V_j)^2, logprior = -lam/2 (|U|^2 + |V|^2)."""
    ij, r = batch
    u, v = (np.asarray(l, dtype=dtype) for l in leaves)
    ii, jj = np.asarray(ij)[:, 0], np.asarray(ij)[:, 1]
    res = np.asarray(r, dtype=dtype) - (u[ii] * v[jj]).sum(axis=1)
    lik = (-0.5 * self.alpha * res * res).sum(dtype=dtype)
    pri = -0.5 * self.lam * ((u * u).sum(dtype=dtype) + (v * v).sum(dtype=dtype))
    return dtype(pri + dtype(n_data) * lik / dtype(len(ii)))


PMF.log_post = _pmf_log_post
