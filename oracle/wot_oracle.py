"""CPU oracle for the Waddington-OT transport-map hot path.  TEST INFRASTRUCTURE ONLY.

This module is a float64 NumPy restatement of the reference algorithm.  It exists so the CUDA path
can be checked; it is never imported by the product package ``wot_b200``.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``
may import it.

Parity status: PINNED.  ``tests/golden/*.npz`` hold outputs of the *unmodified* reference
(``/root/reference/wot/ot/optimal_transport.py`` and ``wot/ot/ot_model.py``, imported by path in
the build container by ``tests/golden/make_golden.py``); ``tests/test_oracle.py`` checks this
restatement against them (couplings to 1e-12, identical iteration / batch counts) and against the
reference's own golden case (``/root/reference/tests/test_transport.py:20-32``: a 3x3 cost of
0/100 at epsilon=0.01 must give the identity within atol=0.01).

Reference lines followed (all relative to /root/reference/):
  * growth loop                    wot/ot/optimal_transport.py:10-33
  * fdiv / fdivstar                wot/ot/optimal_transport.py:37-42
  * primal / dual objectives       wot/ot/optimal_transport.py:45-62
  * default solver (duality gap)   wot/ot/optimal_transport.py:67-164
  * fixed-iteration solver         wot/ot/optimal_transport.py:167-236
  * default cost matrix            wot/ot/ot_model.py:242-253 (scipy cdist 'sqeuclidean' via
                                   sklearn pairwise_distances, then division by np.median)
  * growth bookkeeping             wot/ot/ot_model.py:312-325

Two evaluations of the final-stage duality gap are provided:
  gap='dense'     the reference's own arithmetic (I x J temporaries in primal/dual)
  gap='marginal'  the algebraically identical form that needs only row sums, column sums and the
                  dual potentials (SURVEY.md section 8 a-note).  It is what the CUDA path evaluates
                  on the device; having both here lets the tests prove the identity.
"""
from __future__ import annotations

import logging
from dataclasses import dataclass, field

import numpy as np

log = logging.getLogger("wot")

N_EPS_STAGES = 6  # epsilon_scalings + 1, optimal_transport.py:101,116


# --------------------------------------------------------------------------------------------
# objectives
# --------------------------------------------------------------------------------------------
def kl_div(lam, x, ref, w):
    """lam * sum w (x log(x/ref) - x + ref)        -- optimal_transport.py:37-38 (fdiv)."""
    return lam * np.sum(w * (x * np.log(x / ref) - x + ref))


def kl_conj(lam, pot, ref, w):
    """lam * sum ref w (exp(pot/lam) - 1)           -- optimal_transport.py:41-42 (fdivstar)."""
    return lam * np.sum((ref * w) * (np.exp(pot / lam) - 1.0))


def primal_dense(C, K0, R, dx, dy, p, q, eps, lam1, lam2):
    """Primal objective with the reference's dense arithmetic -- optimal_transport.py:45-53."""
    n_i, n_j = len(p), len(q)
    with np.errstate(divide="ignore", invalid="ignore"):
        ent = R * np.nan_to_num(np.log(R)) - R + K0
        return (kl_div(lam1, R.dot(dy), p, dx) + kl_div(lam2, R.T.dot(dx), q, dy)
                + (eps * np.sum(ent) + np.sum(R * C)) / (n_i * n_j))


def dual_dense(K0, R, dx, dy, p, q, a_full, b_full, eps, lam1, lam2):
    """Dual objective with the reference's dense arithmetic -- optimal_transport.py:56-62."""
    n_i, n_j = len(p), len(q)
    return (-kl_conj(lam1, -eps * np.log(a_full), p, dx) - kl_conj(lam2, -eps * np.log(b_full), q, dy)
            - eps * np.sum(R - K0) / (n_i * n_j))


def gap_from_marginals(r, c, f, g, sum_k0, p, q, eps, lam1, lam2):
    """(primal, dual) from row sums r, column sums c of R = a K b and potentials f, g only.

    Identity (SURVEY.md 8 a-note): eps R log R + R C = R (f_i + g_j), hence
      primal = KL1(r/J) + KL2(c/I) + (f.r + g.c - eps sum(R) + eps sum(K0)) / (I J)
      dual   = -lam1 sum p/I (exp(-f/lam1) - 1) - lam2 sum q/J (exp(-g/lam2) - 1)
               - eps (sum(R) - sum(K0)) / (I J)
    """
    n_i, n_j = len(r), len(c)
    dx, dy = 1.0 / n_i, 1.0 / n_j
    s_r = np.sum(r)
    with np.errstate(divide="ignore", invalid="ignore"):
        pri = (kl_div(lam1, r * dy, p, dx) + kl_div(lam2, c * dx, q, dy)
               + (np.dot(f, r) + np.dot(g, c) - eps * s_r + eps * sum_k0) / (n_i * n_j))
    dua = (-kl_conj(lam1, -f, p, dx) - kl_conj(lam2, -g, q, dy)
           - eps * (s_r - sum_k0) / (n_i * n_j))
    return pri, dua


# --------------------------------------------------------------------------------------------
# solver bookkeeping returned next to the coupling (the reference keeps these as locals)
# --------------------------------------------------------------------------------------------
@dataclass
class SolveInfo:
    iters: int = 0
    batches: list = field(default_factory=lambda: [0] * N_EPS_STAGES)
    tau_absorptions: int = 0
    gaps: list = field(default_factory=list)      # final-stage duality gaps, one per batch
    primal: float = float("nan")
    dual: float = float("nan")
    gap: float = float("inf")
    hit_max_iter: bool = False
    f: np.ndarray | None = None                   # u + eps log a
    g: np.ndarray | None = None                   # v + eps log b
    eps_final: float = float("nan")


def _gibbs(u, v, C, eps):
    return np.exp((u[:, None] - C + v[None, :]) / eps)


def _half_steps(K, a, b, u, v, p, q, dx, dy, eps, lam1, lam2):
    """One Sinkhorn iteration -- optimal_transport.py:133-134 (and :206-207, :231-232)."""
    al1, al2 = lam1 / (lam1 + eps), lam2 / (lam2 + eps)
    a = (p / K.dot(b * dy)) ** al1 * np.exp(-u / (lam1 + eps))
    b = (q / K.T.dot(a * dx)) ** al2 * np.exp(-v / (lam2 + eps))
    return a, b


def _exceeds(a, b, tau):
    return max(np.max(np.abs(a)), np.max(np.abs(b))) > tau


# --------------------------------------------------------------------------------------------
# default solver
# --------------------------------------------------------------------------------------------
def optimal_transport_duality_gap(C, G, lambda1, lambda2, epsilon, batch_size, tolerance, tau,
                                  epsilon0, max_iter, gap="dense", info=None, **ignored):
    """Restatement of optimal_transport.py:67-164.  Returns the I x J coupling (float64).

    ``info`` (a SolveInfo) is filled with iteration/batch counts and the dual potentials.
    """
    info = info if info is not None else SolveInfo()
    C = np.asarray(C, dtype=np.float64)
    n_i, n_j = C.shape
    G = np.asarray(G, dtype=np.float64)
    dx, dy = np.full(n_i, 1.0 / n_i), np.full(n_j, 1.0 / n_j)
    p, q = G, np.full(n_j, np.average(G))                      # :107-108
    u, v = np.zeros(n_i), np.zeros(n_j)
    a, b = np.ones(n_i), np.ones(n_j)
    shrink = np.exp(-np.log(epsilon) / (N_EPS_STAGES - 1))     # :102
    eps = epsilon0 * shrink                                    # :113
    R = None
    gap_val = np.inf
    for stage in range(N_EPS_STAGES):
        last = stage == N_EPS_STAGES - 1
        gap_val = np.inf
        u = u + eps * np.log(a)                                # absorb at the OLD eps, :118-119
        v = v + eps * np.log(b)
        eps = eps / shrink                                     # :120
        K0 = np.exp(-C / eps) if gap == "dense" else None      # :121
        sum_k0 = float(np.sum(np.exp(-C / eps))) if (gap != "dense" and last) else 0.0
        K = _gibbs(u, v, C, eps)                               # :124
        a, b = np.ones(n_i), np.ones(n_j)
        prev_a, prev_b = a, b
        limit = tolerance if last else 1e-6                    # :127
        while gap_val > limit:                                 # NaN ends the loop, as in the reference
            for _ in range(batch_size if last else 5):         # :130
                info.iters += 1
                prev_a, prev_b = a, b
                a, b = _half_steps(K, a, b, u, v, p, q, dx, dy, eps, lambda1, lambda2)
                if _exceeds(a, b, tau):                        # :137-141
                    u = u + eps * np.log(a)
                    v = v + eps * np.log(b)
                    K = _gibbs(u, v, C, eps)
                    a, b = np.ones(n_i), np.ones(n_j)
                    info.tau_absorptions += 1
                if info.iters >= max_iter:                     # :143-145 (NOT divided by J)
                    log.warning("Reached max_iter with duality gap still above threshold. Returning")
                    info.hit_max_iter = True
                    info.f, info.g, info.eps_final = u + eps * np.log(a), v + eps * np.log(b), eps
                    return (K.T * a).T * b
            info.batches[stage] += 1
            a_full, b_full = a * np.exp(u / eps), b * np.exp(v / eps)   # :148-149
            if last:
                R = (K.T * a).T * b                            # :153
                if gap == "dense":
                    pri = primal_dense(C, K0, R, dx, dy, p, q, eps, lambda1, lambda2)
                    dua = dual_dense(K0, R, dx, dy, p, q, a_full, b_full, eps, lambda1, lambda2)
                else:
                    pri, dua = gap_from_marginals(R.sum(axis=1), R.sum(axis=0), u + eps * np.log(a),
                                                  v + eps * np.log(b), sum_k0, p, q, eps, lambda1, lambda2)
                gap_val = (pri - dua) / abs(pri)               # :156
                info.primal, info.dual = float(pri), float(dua)
                info.gaps.append(float(gap_val))
            else:                                              # :158-160
                gap_val = max(
                    np.linalg.norm(a_full - prev_a * np.exp(u / eps)) / (1 + np.linalg.norm(a_full)),
                    np.linalg.norm(b_full - prev_b * np.exp(v / eps)) / (1 + np.linalg.norm(b_full)))
    info.gap = float(gap_val)
    info.f, info.g, info.eps_final = u + eps * np.log(a), v + eps * np.log(b), eps
    if np.isnan(gap_val):                                      # :162-163
        raise RuntimeError("Overflow encountered in duality gap computation, please report this incident")
    return R / n_j                                             # :164


# --------------------------------------------------------------------------------------------
# fixed-iteration solver
# --------------------------------------------------------------------------------------------
def transport_stablev2(C, lambda1, lambda2, epsilon, scaling_iter, G, tau, epsilon0, extra_iter,
                       inner_iter_max, info=None, **ignored):
    """Restatement of optimal_transport.py:167-236."""
    info = info if info is not None else SolveInfo()
    C = np.asarray(C, dtype=np.float64)
    G = np.asarray(G, dtype=np.float64)
    n_i, n_j = C.shape
    warm = tau is not None                                      # :181
    eps = epsilon0 if warm else epsilon                         # :187
    dx, dy = np.full(n_i, 1.0 / n_i), np.full(n_j, 1.0 / n_j)
    p, q = G, np.full(n_j, np.average(G))
    u, v = np.zeros(n_i), np.zeros(n_j)
    a, b = np.ones(n_i), np.ones(n_j)
    K = np.exp(-C / eps)                                        # :197
    since, level = 0, 0
    for _ in range(int(scaling_iter)):                          # :204
        a, b = _half_steps(K, a, b, u, v, p, q, dx, dy, eps, lambda1, lambda2)
        info.iters += 1
        since += 1
        if _exceeds(a, b, tau):                                 # :211-216
            u = u + eps * np.log(a)
            v = v + eps * np.log(b)
            K = _gibbs(u, v, C, eps)
            a, b = np.ones(n_i), np.ones(n_j)
            info.tau_absorptions += 1
        if warm and since == inner_iter_max:                    # :218-228
            level += 1
            since = 0
            u = u + eps * np.log(a)
            v = v + eps * np.log(b)
            eps = (epsilon0 - epsilon) * np.exp(-level) + epsilon   # get_reg, :184-185
            K = _gibbs(u, v, C, eps)
            a, b = np.ones(n_i), np.ones(n_j)
    for _ in range(int(extra_iter)):                            # :230-232
        a, b = _half_steps(K, a, b, u, v, p, q, dx, dy, eps, lambda1, lambda2)
        info.iters += 1
    info.f, info.g, info.eps_final = u + eps * np.log(a), v + eps * np.log(b), eps
    return (K.T * a).T * b / n_j                                # :234-236


# --------------------------------------------------------------------------------------------
# growth loop
# --------------------------------------------------------------------------------------------
def compute_transport_matrix(solver, **params):
    """Restatement of optimal_transport.py:10-33.  Returns (tmap, [G_0 .. G_{n-1}])."""
    learned = []
    rows = params["G"]
    tmap = None
    for it in range(params["growth_iters"]):
        if it > 0:
            rows = tmap.sum(axis=1)                              # :27
        params["G"] = rows
        learned.append(rows)
        tmap = solver(**params)                                  # :30
    return tmap, learned


# --------------------------------------------------------------------------------------------
# default cost
# --------------------------------------------------------------------------------------------
def sqeuclidean(x, y, block=1024):
    """Direct-difference squared Euclidean distances, sum_k (x_ik - y_jk)^2, float64.

    This is what scipy's cdist('sqeuclidean') evaluates (the reference reaches it through
    sklearn.metrics.pairwise.pairwise_distances, ot_model.py:249-251); it is NOT the
    |x|^2 + |y|^2 - 2 x.y expansion.
    """
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    out = np.empty((x.shape[0], y.shape[0]))
    for s in range(0, x.shape[0], block):
        xs = x[s:s + block]
        acc = np.zeros((xs.shape[0], y.shape[0]))
        for k in range(x.shape[1]):
            d = xs[:, k, None] - y[None, :, k]
            acc += d * d
        out[s:s + block] = acc
    return out


def compute_default_cost_matrix(a, b, eigenvals=None):
    """Restatement of ot_model.py:242-253: optional scaling by diag(singular values), pairwise
    squared Euclidean distance, division by the median of all I*J entries."""
    a = np.asarray(a)
    b = np.asarray(b)
    if eigenvals is not None:
        a = a.dot(eigenvals)
        b = b.dot(eigenvals)
    cost = sqeuclidean(a, b)
    return cost / np.median(cost)


def growth_columns(learned_growth, tmap, delta_days):
    """obs columns g0..gN -- ot_model.py:319-325."""
    seq = list(learned_growth) + [tmap.sum(axis=1)]
    return {"g%d" % i: np.power(g, 1.0 / delta_days) for i, g in enumerate(seq)}


DEFAULTS = dict(local_pca=30, growth_iters=1, epsilon=0.05, lambda1=1, lambda2=50, epsilon0=1, tau=10000,
                scaling_iter=3000, inner_iter_max=50, tolerance=1e-8, max_iter=1e7, batch_size=5,
                extra_iter=1000)  # ot_model.py:85-87
