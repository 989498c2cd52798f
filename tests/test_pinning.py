"""Pinning a14 (manikde! bandwidth) and a15 (manifoldProduct) against MATHEMATICS instead of against the
builder's own restatement ("what's missing" #3).  The algorithms live in ApproxManifoldProducts /
KernelDensityEstimate (call sites ApproxConv.jl:36-42, GraphProductOperations.jl:53-60), which are not vendored and
cannot run here, so sample-level parity with Julia stays unpinned; what CAN be checked independently is

  a15  the product of F kernel density estimates with N kernels each is EXACTLY an N^F-component Gaussian mixture
       whose weights, means and variances have closed forms.  The multiscale Gibbs sampler must draw from it: the
       pooled output of many seeded products is compared with the exact mixture CDF (Kolmogorov-Smirnov distance).
       The separated-modes case quantifies the known one-sweep leakage of the sampler (Ihler et al. 2003; the
       reference's own test calls it "8/10 quality", test/testMultiHypo3Door.jl:7).
  a14  the returned bandwidth must minimise the leave-one-out negative log-likelihood: a dense scan of the objective
       (numpy, written independently of the oracle's C) over the search bracket.

The same checks run on the CPU oracle (`-m "not gpu"`) and on the CUDA kernels (`-m gpu`)."""
import numpy as np
import pytest
from scipy.special import ndtr

import oracle as O
import parity_cases as PC


# ------------------------------------------------------------------------------------------ exact product mixture
def exact_product_mixture(pts, bws):
    """1-D KDEs p_j = 1/N sum_i N(x; x_ji, h_j^2): prod_j p_j = sum over label tuples of w * N(x; mu, s^2) with
    s^-2 = sum_j h_j^-2, mu = s^2 sum_j x_j / h_j^2, log w = -1/2 sum_j (x_j - mu)^2 / h_j^2 (+ const)."""
    lam = [1.0 / b ** 2 for b in bws]
    L = sum(lam)
    grids = np.meshgrid(*pts, indexing="ij")
    mu = sum(l * g for l, g in zip(lam, grids)) / L
    logw = -0.5 * sum(l * (g - mu) ** 2 for l, g in zip(lam, grids))
    w = np.exp(logw - logw.max()).ravel()
    w /= w.sum()
    keep = w > 1e-14                      # drop numerically irrelevant components (keeps the CDF evaluation small)
    return w[keep] / w[keep].sum(), mu.ravel()[keep], 1.0 / np.sqrt(L)


def ks_distance(samples, w, mu, s):
    xs = np.sort(np.asarray(samples))
    cdf = np.empty_like(xs)
    for k in range(0, len(xs), 512):
        cdf[k:k + 512] = (ndtr((xs[k:k + 512, None] - mu[None, :]) / s) * w[None, :]).sum(axis=1)
    n = len(xs)
    return max(np.abs(cdf - np.arange(1, n + 1) / n).max(), np.abs(cdf - np.arange(0, n) / n).max())


def mixture_cases():
    R = np.random.default_rng(1)
    N = 100
    yield "overlap2", [R.normal(0, 1, N), R.normal(0.5, 1, N)], 0.03, None
    yield "bimodal_x_prior", [np.concatenate([R.normal(-30, 1, N // 2), R.normal(40, 1, N // 2)]),
                              R.normal(38, 3, N)], 0.03, None
    yield "wide_x_narrow", [R.normal(0, 3, N), R.normal(1.0, 0.2, N)], 0.03, None
    n3 = 40                                # N^3 = 64 000 exact components
    yield "overlap3", [R.normal(0, 1, n3), R.normal(1, 2, n3), R.normal(-1, 0.5, n3)], 0.04, None
    # separated modes: doors at 0/10/20/40 times modes at 10/50: the exact product has ONE mode (at 10).  A label
    # Gibbs sampler with one sweep per level cannot always leave the nearest wrong pair (40, 50): quantified, not hidden
    four = np.concatenate([R.normal(m, 0.5, N // 4) for m in (0.0, 10.0, 20.0, 40.0)])
    two = np.concatenate([R.normal(10, 0.5, N // 2), R.normal(50, 0.5, N // 2)])
    yield "separated", [four, two], None, (5.0, 15.0, 0.70)


def _check_product_sampler(product_fn, nseeds):
    for name, pts, ks_max, mode in mixture_cases():
        bws = [float(O.kde_bandwidth(p.reshape(-1, 1))[0]) for p in pts]
        w, mu, s = exact_product_mixture(pts, bws)
        dens = np.stack([p.reshape(-1, 1) for p in pts])
        pooled = product_fn(dens, np.array(bws).reshape(-1, 1), nseeds)
        assert pooled.shape == (nseeds * len(pts[0]),) and np.isfinite(pooled).all()
        if ks_max is not None:
            ks = ks_distance(pooled, w, mu, s)
            assert ks < ks_max, f"{name}: KS distance {ks:.4f} to the exact {len(w)}-component product mixture"
            ex_mean = float((w * mu).sum())
            ex_std = float(np.sqrt((w * (mu ** 2 + s ** 2)).sum() - ex_mean ** 2))
            assert abs(pooled.mean() - ex_mean) < 0.08 * ex_std + 4 * ex_std / np.sqrt(len(pooled)), name
            assert abs(pooled.std() / ex_std - 1.0) < 0.08, (name, pooled.std(), ex_std)
        else:
            lo, hi, frac = mode
            inside = ((pooled > lo) & (pooled < hi)).mean()
            exact_inside = w[(mu > lo) & (mu < hi)].sum()
            assert exact_inside > 0.999
            assert inside >= frac, f"{name}: only {inside:.3f} of the samples in the exact product's single mode"


def test_oracle_product_samples_the_exact_mixture():
    def product_fn(dens, bws, nseeds):
        out = [O.product(dens, bws, 1, seed=1000 + k, call_id=k)[0][:, 0] for k in range(nseeds)]
        return np.concatenate(out)
    _check_product_sampler(product_fn, 40)


@pytest.mark.gpu
def test_gpu_product_samples_the_exact_mixture(built):
    P, xs, fs = PC.chain_problem(n=2, N=8)
    eng = P.engine()

    def product_fn(dens, bws, nseeds):
        # one Philox key, one call id per product: independent streams
        return np.concatenate([eng.product(dens, bws, 1, call_id=7000 + k)[0][:, 0] for k in range(nseeds)])
    try:
        _check_product_sampler(product_fn, 64)
    finally:
        eng.close()


# ------------------------------------------------------------------------------------------ bandwidth = LOO minimiser
def loo_nll_numpy(x, h, circular=False):
    """-1/N sum_i log( 1/(N-1) sum_{j != i} N(x_i - x_j; 0, h^2) ), written from the definition"""
    d = x[:, None] - x[None, :]
    if circular:
        d = (d + np.pi) % (2 * np.pi) - np.pi
    k = np.exp(-0.5 * (d / h) ** 2) / (np.sqrt(2 * np.pi) * h)
    np.fill_diagonal(k, 0.0)
    with np.errstate(divide="ignore"):
        return float(-np.mean(np.log(k.sum(axis=1) / (len(x) - 1))))


def bandwidth_shapes():
    R = np.random.default_rng(5)
    for n, d, cm in [(100, 1, 0), (100, 2, 0), (150, 1, 1), (200, 3, 0b010), (256, 1, 0), (17, 4, 0), (64, 1, 0)]:
        pts = R.normal(0, 1, (n, d)) * R.uniform(0.1, 10, d)
        for c in range(d):
            if (cm >> c) & 1:
                pts[:, c] = PC.wrap(R.normal(2.5, 0.6, n))
        yield pts, cm
    yield np.concatenate([R.normal(-4, 0.3, (50, 1)), R.normal(6, 1.5, (50, 1))]), 0        # bimodal
    yield PC.wrap(R.normal(3.1, 0.25, (120, 1))), 1                                         # across the +-pi seam


def _check_bandwidth_is_minimiser(bw_fn):
    for pts, cm in bandwidth_shapes():
        bw = bw_fn(pts, cm)
        for c in range(pts.shape[1]):
            x, circ, h = pts[:, c], bool((cm >> c) & 1), float(bw[c])
            f0 = loo_nll_numpy(x, h, circ)
            if circ:
                lo, hi, tol = 1e-3, 2 * np.pi, 2.5e-3             # Optim golden section, rel_tol 1e-3 on [1e-3, 2 pi]
            else:
                xs = np.sort(x)
                lo, hi, tol = max(np.diff(xs).min(), 1e-6), xs[-1] - xs[0], 2.5e-2   # KDE golden(.., tol = 1e-2)
            grid = np.exp(np.linspace(np.log(lo), np.log(hi), 1500))
            f = np.array([loo_nll_numpy(x, g, circ) for g in grid])
            k = int(np.argmin(f))
            # 1. nothing on the dense scan beats the returned point by more than the flatness a 1 % (0.1 %) bracket
            #    leaves: f is smooth, so |f(h*) - f(h)| <= 1/2 f''(h*) (tol h*)^2, estimated from the scan itself
            curv = (f[min(k + 1, len(f) - 1)] - 2 * f[k] + f[max(k - 1, 0)]) / (np.log(grid[1] / grid[0]) ** 2)
            slack = 0.5 * max(curv, 0.0) * (3 * tol) ** 2 + 1e-12
            assert f0 <= f[k] + slack, (pts.shape, c, h, grid[k], f0 - f[k], slack)
            # 2. the returned point is a local minimum at the search tolerance
            assert f0 <= loo_nll_numpy(x, h * (1 + 2 * tol), circ) + 1e-12
            assert f0 <= loo_nll_numpy(x, h / (1 + 2 * tol), circ) + 1e-12
            # 3. and sits within the tolerance of the scan's global minimiser
            assert abs(np.log(h / grid[k])) <= 3 * tol + np.log(grid[1] / grid[0]), (pts.shape, c, h, grid[k])


def test_oracle_bandwidth_minimises_the_loo_objective():
    _check_bandwidth_is_minimiser(lambda pts, cm: O.kde_bandwidth(pts, cm))


@pytest.mark.gpu
def test_gpu_bandwidth_minimises_the_loo_objective(built):
    P, xs, fs = PC.chain_problem(n=2, N=8)
    eng = P.engine()
    try:
        _check_bandwidth_is_minimiser(lambda pts, cm: eng.kde_bandwidth(pts, cm))
    finally:
        eng.close()


def test_oracle_loo_objective_equals_the_definition():
    R = np.random.default_rng(3)
    for n, circ in ((50, False), (101, False), (64, True)):
        x = PC.wrap(R.normal(3.0, 0.4, n)) if circ else R.normal(0, 2, n)
        for h in (0.25, 0.6, 1.7):
            ref = loo_nll_numpy(x, h, circ)
            assert np.isfinite(ref) and abs(O.loo_nll(x, h, int(circ)) - ref) < 1e-10 * max(1.0, abs(ref))
