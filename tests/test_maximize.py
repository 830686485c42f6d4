"""Rows (f.1)-(f.3): constraint transforms, KL term and the batched Newton trust region -- CPU tests against
plain per-source restatements with autograd derivatives (tests/newton_oracle.py), plus the reference's own
optimiser-level checks (test/test_optimization.jl, test/test_constraints.jl round trips) driven through the
product code with the oracle as the ELBO evaluator (the CUDA plan takes that place on a GPU)."""
import math

import numpy as np
import pytest
import torch

import cases
import newton_oracle as no
import oracle_lib
import celeste_jl_b200 as cj
from celeste_jl_b200 import constraint_transforms as ct
from celeste_jl_b200 import elbo_maximize as em
from celeste_jl_b200 import synthetic
from celeste_jl_b200.kl import KLTerm
from celeste_jl_b200.model import ids


def _sample_vp(n, seed=0):
    rng = np.random.default_rng(seed)
    vp = np.stack([cj.generic_init_source([10.0 + rng.normal(), 12.0 + rng.normal()]) for _ in range(n)])
    vp[:, 2:4] = rng.uniform(0.05, 0.95, (n, 2))
    vp[:, 4] = rng.uniform(-2, 2, n)
    vp[:, 5] = rng.uniform(0.3, 8, n)
    vp[:, 6:8] = rng.uniform(0, 6, (n, 2))
    vp[:, 8:10] = rng.uniform(1e-3, 0.09, (n, 2))
    vp[:, 10:18] = rng.normal(0, 1, (n, 8))
    vp[:, 18:26] = rng.uniform(1e-3, 0.9, (n, 8))
    a = rng.uniform(0.05, 0.95, n)
    vp[:, 26], vp[:, 27] = a, 1 - a
    for i in range(2):
        k = rng.dirichlet(np.ones(8) * 2, n) * 0.98 + 0.0025
        vp[:, 28 + 8 * i:36 + 8 * i] = k / k.sum(axis=1, keepdims=True)
    return vp


def test_transform_round_trip_and_bounds():
    """test/test_constraints.jl round trips: to_bound(to_free(x)) == x; bound values inside their boxes."""
    vp = torch.tensor(_sample_vp(16))
    lo, hi = ct.box_bounds(vp, 1e-4)
    vp = ct.enforce(vp, lo, hi)
    free = ct.to_free(vp, lo, hi)
    assert free.shape == (16, 41) and torch.isfinite(free).all()
    back = ct.to_bound(free, lo, hi)
    assert torch.allclose(back, vp, rtol=1e-9, atol=1e-12)
    rnd = ct.to_bound(torch.randn(16, 41, dtype=torch.float64) * 3, lo, hi)
    assert ((rnd[:, :26] > lo) & (rnd[:, :26] < hi)).all()
    for first, n, lower in ct.SIMPLEXES:
        s = rnd[:, first:first + n]
        assert torch.allclose(s.sum(dim=1), torch.ones(16, dtype=torch.float64)) and (s > lower).all()
    # matches the plain per-source restatement
    for b in range(3):
        ref = no.to_bound_single(free[b], lo[b].numpy(), hi[b].numpy())
        assert torch.allclose(ref, back[b], rtol=1e-12, atol=1e-14)
        assert np.allclose(no.to_free_single(vp[b].numpy(), lo[b].numpy(), hi[b].numpy()), free[b].numpy(), rtol=1e-10)


def test_enforce_clips_like_reference():
    vp = torch.tensor(_sample_vp(4))
    lo, hi = ct.box_bounds(vp, 1e-4)
    bad = vp.clone()
    bad[:, 2] = 1.5           # gal_frac_dev above its box
    bad[:, 8] = 0.0           # flux_scale below
    bad[:, 26], bad[:, 27] = 1.2, -0.1
    out = ct.enforce(bad, lo, hi)
    assert (out[:, 2] < 0.99).all() and (out[:, 2] > 0.98).all()
    assert (out[:, 8] > 1e-4).all()
    assert torch.allclose(out[:, 26:28].sum(dim=1), torch.ones(4, dtype=torch.float64), atol=1e-8)
    assert (out[:, 26:28] > 0.005).all()
    assert torch.isfinite(ct.to_free(out, lo, hi)).all()


def test_propagate_derivatives_matches_autograd():
    """propagate_derivatives! (ConstraintTransforms.jl:373-396): closed forms == nested AD of to_bound!."""
    rng = np.random.default_rng(1)
    vp = torch.tensor(_sample_vp(5, seed=2))
    lo, hi = ct.box_bounds(vp, 1e-4)
    free = ct.to_free(ct.enforce(vp, lo, hi), lo, hi) + torch.tensor(rng.normal(0, 0.3, (5, 41)))
    g = torch.tensor(rng.normal(size=(5, 44)))
    A = rng.normal(size=(5, 44, 44))
    H = torch.tensor(A + A.transpose(0, 2, 1))
    gf, Hf = ct.propagate_derivatives(free, lo, hi, g, H)
    for b in range(5):
        rg, rH = no.propagate_ad(free[b].numpy(), lo[b].numpy(), hi[b].numpy(), g[b].numpy(), H[b].numpy())
        assert np.allclose(gf[b].numpy(), rg, rtol=1e-10, atol=1e-12 * np.abs(rg).max())
        assert np.allclose(Hf[b].numpy(), rH, rtol=1e-9, atol=1e-11 * np.abs(rH).max())
        assert np.array_equal(Hf[b].numpy(), Hf[b].numpy().T)


def test_kl_closed_form_matches_autograd():
    """elbo_kl.jl: value, gradient (ReverseDiff in the reference) and Hessian (ForwardDiff-over-ReverseDiff)."""
    vp = _sample_vp(6, seed=3)
    v, g, H = KLTerm("cpu")(torch.tensor(vp), order=2)
    for b in range(6):
        rv, rg, rH = no.kl_ad(vp[b])
        assert abs(float(v[b]) - rv) <= 1e-12 * abs(rv)
        assert np.allclose(g[b].numpy(), rg, rtol=1e-10, atol=1e-12 * np.abs(rg).max())
        assert np.allclose(H[b].numpy(), rH, rtol=1e-9, atol=1e-11 * np.abs(rH).max())
    v1, g1, _ = KLTerm("cpu")(torch.tensor(vp), order=1)
    assert torch.equal(v1, v) and torch.equal(g1, g)


def test_kl_known_values():
    """test/test_kl.jl spirit: KL of a distribution with itself is 0; gaussian_kl closed form."""
    prior = KLTerm("cpu")
    vs = _sample_vp(1, seed=4)[0]
    vs[26:28] = [0.95, 0.05]                       # == prior.is_star
    for i in range(2):
        vs[28 + 8 * i:36 + 8 * i] = np.exp(prior.log_k[i].numpy())
        vs[6 + i], vs[8 + i] = float(prior.flux_mean[i]), float(prior.flux_var[i])
    v, _, _ = prior(torch.tensor(vs[None]), order=0)
    # what remains is the colour KL and the radius log-probability
    rem = 0.0
    for i in range(2):
        for d in range(8):
            mu2, P = prior.mu2[i, d].numpy(), prior.prec[i, d].numpy()
            diff = mu2 - vs[10 + 4 * i:14 + 4 * i]
            var1 = vs[18 + 4 * i:22 + 4 * i]
            t = (np.diag(P) * var1).sum() - 4 + diff @ P @ diff + float(prior.logdet[i, d]) - np.log(var1).sum()
            rem -= vs[26 + i] * vs[28 + 8 * i + d] * 0.5 * t
    rem += -0.5 * (math.log(2 * math.pi) + math.log(prior.rad_var) + (vs[5] - prior.rad_mean) ** 2 / prior.rad_var)
    assert float(v[0]) == pytest.approx(rem, rel=1e-12)


def test_tr_subproblem_optimality():
    """Exact subproblem: |s| <= delta; (H + lam I) s = -g with lam >= 0, lam (delta - |s|) = 0, H + lam I >= 0;
    agrees with the per-source bisection solver, including indefinite and hard cases."""
    rng = np.random.default_rng(5)
    n, B = 12, 40
    A = rng.normal(size=(B, n, n))
    H = A + A.transpose(0, 2, 1)
    H[:10] = H[:10] @ H[:10].transpose(0, 2, 1) + 0.1 * np.eye(n)          # positive definite
    g = rng.normal(size=(B, n))
    delta = rng.uniform(0.05, 5.0, B)
    # a hard case: g orthogonal to the lowest eigenvector
    ev, Q = np.linalg.eigh(H[-1])
    g[-1] = Q[:, 1:] @ rng.normal(size=n - 1) * 1e-3
    delta[-1] = 3.0
    s, m, interior = em.solve_tr_subproblem(torch.tensor(g), torch.tensor(H), torch.tensor(delta))
    s, m = s.numpy(), m.numpy()
    for b in range(B):
        assert np.linalg.norm(s[b]) <= delta[b] * (1 + 1e-8)
        rs, rm, rint = no.tr_subproblem_single(g[b], H[b], delta[b])
        assert m[b] == pytest.approx(rm, rel=1e-6, abs=1e-10)
        assert bool(interior[b]) == rint
        assert m[b] <= 1e-12


def test_tr_subproblem_kernel_under_emulation():
    """csrc/newton_kernels.cuh (Jacobi + secular equation, one block per source) == the torch restatement,
    on 41 x 41 problems like the free-space Hessians (positive definite, indefinite and a hard case)."""
    import emul_lib
    rng = np.random.default_rng(6)
    n, B = 41, 10
    A = rng.normal(size=(B, n, n))
    H = A + A.transpose(0, 2, 1)
    H[:4] = np.einsum("bij,bkj->bik", A[:4], A[:4]) + 0.5 * np.eye(n)
    g = rng.normal(size=(B, n))
    delta = rng.uniform(0.1, 10.0, B)
    ev, Q = np.linalg.eigh(H[-1])
    g[-1] = Q[:, 1:] @ rng.normal(size=n - 1) * 1e-3
    delta[-1] = 5.0
    s, m, interior = emul_lib.tr_subproblem(g, H, delta)
    rs, rm, rint = em.solve_tr_subproblem(torch.tensor(g), torch.tensor(H), torch.tensor(delta))
    assert np.array_equal(interior, rint.numpy())
    assert np.allclose(m, rm.numpy(), rtol=1e-9, atol=1e-12)
    for b in range(B):
        assert np.linalg.norm(s[b]) <= delta[b] * (1 + 1e-9)
        if b != B - 1:        # the hard-case step is unique only up to the sign/choice of the eigenvector
            assert np.allclose(s[b], rs[b].numpy(), rtol=1e-7, atol=1e-9 * np.abs(rs[b].numpy()).max())
        model = g[b] @ s[b] + 0.5 * s[b] @ H[b] @ s[b]
        assert model == pytest.approx(m[b], rel=1e-9, abs=1e-11)


class OracleRunner:
    """ELBO evaluator for BatchMaximizer on a machine without a GPU: the oracle (checker)."""

    def __init__(self, images, patches, plan_like):
        self.of = oracle_lib.OracleField(images, patches)
        self.p = plan_like

    def __call__(self, bm):
        out = self.of.elbo_csr(self.p.task_ptr, self.p.src, self.p.active_ptr, self.p.act,
                               bm.vp_all.numpy().ravel(), mode=2, n_threads=8)
        bm.v.copy_(torch.from_numpy(out["v"]))
        bm.d.copy_(torch.from_numpy(out["d"]))
        bm.h.copy_(torch.from_numpy(out["h"]))
        bm.flags.copy_(torch.from_numpy(out["flags"]))


class PlanLike:
    def __init__(self, rows, act):
        from celeste_jl_b200.flatten import csr_tasks
        dummy = [(r, a, np.zeros((44, len(r)))) for r, a in zip(rows, act)]
        self.task_ptr, self.src, self.active_ptr, self.act, _ = csr_tasks(dummy)
        self.n_tasks = len(rows)


def _verify_sample_galaxy(vs, pos):
    """test/test_optimization.jl:10-32."""
    assert vs[ids.is_star[1]] >= 0.99
    assert abs(vs[0] - pos[0]) < 0.1 and abs(vs[1] - pos[1]) < 0.1
    assert abs(vs[ids.gal_axis_ratio] - 0.7) < 0.05
    assert abs(vs[ids.gal_frac_dev] - 0.1) < 0.08
    assert abs(vs[ids.gal_radius_px] - 4.0) < 0.2
    phi = vs[ids.gal_angle] - math.floor(vs[ids.gal_angle] / math.pi) * math.pi
    assert abs(phi - math.pi / 4) < 5 * math.pi / 180
    bright = math.exp(vs[ids.flux_loc[1]] + 0.5 * vs[ids.flux_scale[1]])
    assert abs(bright / synthetic.sample_galaxy_fluxes[2] - 1.0) < 0.05
    true_colors = np.log(synthetic.sample_galaxy_fluxes[1:5] / synthetic.sample_galaxy_fluxes[0:4])
    for b in range(4):
        assert abs(vs[ids.color_mean[b, 1]] - true_colors[b]) < 0.2


@pytest.mark.parametrize("n", [2, 3, 7, 20, 41])
def test_tr_subproblem_kernel_sizes_and_special_matrices(n):
    """The tridiagonal solver at other sizes and on the matrices that take its shortcuts: positive definite with the
    Newton step inside the region (no multisection, one solve) and outside it (secular iteration from lam = 0),
    diagonal H (every Householder reflector is the identity), indefinite, and a singular H."""
    import emul_lib
    rng = np.random.default_rng(60 + n)
    B = 8
    A = rng.normal(size=(B, n, n))
    H = A + A.transpose(0, 2, 1)
    H[0] = A[0] @ A[0].T + np.eye(n)                       # positive definite, interior (delta large)
    H[1] = A[1] @ A[1].T + np.eye(n)                       # positive definite, boundary (delta small)
    H[2] = np.diag(rng.uniform(0.5, 3.0, n))               # diagonal, positive definite
    H[3] = np.diag(np.linspace(-2.0, 3.0, n))              # diagonal, indefinite
    v = rng.normal(size=n)
    H[4] = np.outer(v, v)                                  # singular (rank 1), positive semi-definite
    g = rng.normal(size=(B, n))
    delta = rng.uniform(0.5, 3.0, B)
    delta[0], delta[1] = 1e3, 1e-2
    s, m, interior = emul_lib.tr_subproblem(g, H, delta)
    rs, rm, rint = em.solve_tr_subproblem(torch.tensor(g), torch.tensor(H), torch.tensor(delta))
    assert np.array_equal(interior, rint.numpy())
    assert interior[0] == 1 and interior[1] == 0
    assert np.allclose(m, rm.numpy(), rtol=1e-8, atol=1e-12)
    for b in range(B):
        assert np.linalg.norm(s[b]) <= delta[b] * (1 + 1e-9)
        assert np.allclose(s[b], rs[b].numpy(), rtol=1e-6, atol=1e-8 * max(np.abs(rs[b].numpy()).max(), 1e-30)), b
        model = g[b] @ s[b] + 0.5 * s[b] @ H[b] @ s[b]
        assert model == pytest.approx(m[b], rel=1e-9, abs=1e-11)


def test_galaxy_optimization_recovers_truth():
    """test/test_optimization.jl:53-58 (test_galaxy_optimization: include_kl = false, loc_width = 3)."""
    images, patches, vp, _ = synthetic.gen_sample_galaxy_dataset()
    pl = PlanLike([[1]], [[1]])
    bm = em.BatchMaximizer(pl, np.concatenate(vp), include_kl=False, device="cpu", loc_width=3.0,
                           runner=OracleRunner(images, patches, pl))
    res = bm.run()
    assert res.converged[0] and res.iterations[0] <= 50
    _verify_sample_galaxy(res.vp[0], [8.5, 9.6])


def test_full_elbo_optimization_and_matches_per_source_newton():
    """test/test_optimization.jl:61-67 (with KL, loc_width = 1) and: the batched lock-step driver reaches the
    same optimum as the plain per-source Newton trust region of the checker."""
    images, patches, vp, _ = synthetic.gen_sample_galaxy_dataset()
    pl = PlanLike([[1]], [[1]])
    bm = em.BatchMaximizer(pl, np.concatenate(vp), include_kl=True, device="cpu", loc_width=1.0,
                           runner=OracleRunner(images, patches, pl))
    res = bm.run()
    _verify_sample_galaxy(res.vp[0], [8.5, 9.6])

    def elbo_fn(b):
        v, d, h, _ = oracle_lib.oracle_elbo(images, patches, [b], [1])
        return v, d[:, 0], h
    ref_vs, ref_val, it, calls, conv = no.maximize_single(elbo_fn, vp[0], include_kl=True, loc_width=1.0)
    # both stop the same way (here: the 50-iteration cap of elbo_optim_options, ElboMaximize.jl:95) ...
    assert bool(res.converged[0]) == conv and int(res.iterations[0]) == it and int(res.f_calls[0]) == calls
    # ... at the same point
    assert res.value[0] == pytest.approx(ref_val, rel=1e-6)
    assert np.allclose(res.vp[0][:28], ref_vs[:28], rtol=2e-3, atol=2e-4)


def test_single_source_optimization_leaves_neighbours_alone():
    """test/test_optimization.jl:36-50: only the active source moves."""
    images, patches, vp, _ = synthetic.gen_three_body_dataset()
    pl = PlanLike([[2, 1, 3]], [[1]])
    flat = np.concatenate([vp[1], vp[0], vp[2]])
    bm = em.BatchMaximizer(pl, flat, include_kl=False, device="cpu", loc_width=1.0, max_iters=8,
                           runner=OracleRunner(images, patches, pl))
    res = bm.run()
    after = bm.vp_all.numpy()
    assert not np.allclose(after[0], vp[1])
    assert np.array_equal(after[1], vp[0]) and np.array_equal(after[2], vp[2])


def _newton_buffers(n, x, lo, hi, v, d, h, flags, vp_all, aslot, prior):
    """numpy-backed celeste_newton_buffers + the dict that keeps the arrays alive."""
    from celeste_jl_b200 import _lib
    st = dict(x=x, f=np.zeros(n), g=np.zeros((n, 41)), H=np.zeros((n, 41, 41)), delta=np.zeros(n),
              x_new=np.zeros((n, 41)), m_pred=np.zeros(n), interior=np.zeros(n, dtype=np.int32),
              active=np.ones(n, dtype=np.uint8), converged=np.zeros(n, dtype=np.uint8), iters=np.zeros(n, dtype=np.int32),
              f_calls=np.zeros(n, dtype=np.int32), lo=lo, hi=hi, v=v, d=d, h=h, flags=flags, vp_all=vp_all, aslot=aslot,
              prior=prior)
    for k, a in st.items():
        assert a is None or a.flags["C_CONTIGUOUS"], k
    buf = _lib.celeste_newton_buffers(**{k: (a.ctypes.data if a is not None else None) for k, a in st.items()})
    return buf, st


@pytest.mark.parametrize("include_kl", [True, False])
def test_newton_step_kernel_under_emulation(include_kl):
    """csrc/maximize_kernels.cuh, one step at a time, against the torch restatements it fuses: phase 0 =
    -KL + propagate_derivatives! + first subproblem + to_bound! of the candidate; phase 1 = the trust-region
    update (accepted and rejected steps, radius, convergence flags) + the next candidate."""
    import emul_lib
    rng = np.random.default_rng(11)
    n = 6
    vp = torch.tensor(_sample_vp(n, seed=12))
    lo, hi = ct.box_bounds(vp, 1e-4)
    x = ct.to_free(ct.enforce(vp, lo, hi), lo, hi) + torch.tensor(rng.normal(0, 0.2, (n, 41)))
    kl = KLTerm("cpu")

    def evaluation(seed, pd_shift):
        r = np.random.default_rng(seed)
        A = r.normal(size=(n, 44, 44))
        h = -(np.einsum("bij,bkj->bik", A, A) + pd_shift * np.eye(44))      # ELBO Hessians are mostly negative definite
        return r.normal(size=n) * 100, r.normal(size=(n, 44)), h

    def torch_eval(xx, v, d, h):
        b = ct.to_bound(xx, lo, hi)
        tv, tg, tH = torch.tensor(v), torch.tensor(d), torch.tensor(h)
        if include_kl:
            kv, kg, kH = kl(b, order=2)
            tv, tg, tH = tv + kv, tg + kg, tH + kH
        gf, Hf = ct.propagate_derivatives(xx, lo, hi, tg, tH)
        return -tv, -gf, -Hf, b

    v, d, h = evaluation(1, 5.0)
    flags = np.zeros(n, dtype=np.int32)
    flags[-1] = 1                                                            # a non-finite evaluation: never active
    vp_all = np.zeros((2 * n, 44))
    aslot = np.arange(n, dtype=np.int64) * 2
    xs = x.numpy().copy()
    prior = kl.packed().numpy() if include_kl else None
    buf, st = _newton_buffers(n, xs, lo.numpy().copy(), hi.numpy().copy(), v, d, h.copy(), flags, vp_all, aslot, prior)
    step = lambda ph: emul_lib.newton_stepper(ph, n, buf)

    step(2)
    assert np.allclose(vp_all[aslot], ct.to_bound(x, lo, hi).numpy(), rtol=1e-14, atol=1e-300)
    assert not vp_all[1::2].any()
    # ---- phase 0
    step(0)
    f0, g0, H0, _ = torch_eval(x, v, d, h)
    assert np.allclose(st["f"], f0.numpy(), rtol=1e-13)
    assert np.allclose(st["g"], g0.numpy(), rtol=1e-10, atol=1e-12 * np.abs(g0.numpy()).max())
    assert np.allclose(st["H"], H0.numpy(), rtol=1e-9, atol=1e-11 * np.abs(H0.numpy()).max())
    assert np.array_equal(st["H"], st["H"].transpose(0, 2, 1))
    assert np.array_equal(st["active"], [1] * (n - 1) + [0]) and not st["converged"].any()
    assert np.array_equal(st["delta"], np.ones(n)) and np.array_equal(st["f_calls"], [1] * n)
    delta = torch.ones(n, dtype=torch.float64)
    s, m, interior = em.solve_tr_subproblem(g0, H0, delta)
    ok = slice(0, n - 1)
    assert np.allclose(st["x_new"][ok], (x + s).numpy()[ok], rtol=1e-7, atol=1e-9)
    assert np.allclose(st["m_pred"][ok], m.numpy()[ok], rtol=1e-8)
    assert np.array_equal(st["interior"][ok].astype(bool), interior.numpy()[ok])
    assert np.allclose(vp_all[aslot][ok], ct.to_bound(torch.tensor(st["x_new"]), lo, hi).numpy()[ok], rtol=1e-13)
    # ---- phase 1: craft evaluations so that some steps are accepted and some rejected
    x_new = torch.tensor(st["x_new"].copy())
    v1, d1, h1 = evaluation(2, 5.0)
    m_pred = st["m_pred"].copy()
    # source 0: rho ~ 1 (accept, radius grows unless interior); source 1: worse value (reject, radius shrinks);
    # source 2: tiny improvement relative to the prediction (accept needs rho > 0.1: reject)
    f_target = np.array([f0[0] + m_pred[0], f0[1] + 1.0, f0[2] + 0.05 * m_pred[2], f0[3] + 0.5 * m_pred[3],
                         f0[4] + 0.2 * m_pred[4], 0.0])
    klv = kl(ct.to_bound(x_new, lo, hi), order=0)[0].numpy() if include_kl else np.zeros(n)
    v1[:] = -f_target - klv
    st["v"][:] = v1
    st["d"][:] = d1
    st["h"][:] = h1
    x_before, f_before, H_before = st["x"].copy(), st["f"].copy(), st["H"].copy()
    step(1)
    f1, g1, H1, _ = torch_eval(x_new, v1, d1, h1)
    rho = (f_before[:n - 1] - f1.numpy()[:n - 1]) / (-m_pred[:n - 1])
    for b in range(n - 1):
        accept = rho[b] > em.ETA
        if accept:
            assert np.allclose(st["x"][b], x_new[b].numpy()) and st["f"][b] == pytest.approx(float(f1[b]), rel=1e-13)
            assert np.allclose(st["H"][b], H1[b].numpy(), rtol=1e-9, atol=1e-11 * np.abs(H1[b].numpy()).max())
        else:
            assert np.array_equal(st["x"][b], x_before[b]) and st["f"][b] == f_before[b]
            assert np.array_equal(st["H"][b], H_before[b])
        grows = rho[b] > em.RHO_UPPER and not bool(interior[b])
        assert st["delta"][b] == (0.25 if rho[b] < em.RHO_LOWER else (2.0 if grows else 1.0))
        assert st["iters"][b] == 1 and st["f_calls"][b] == 2
    assert (rho[:5] > em.ETA).tolist() == [True, False, False, True, True]
    assert st["iters"][-1] == 0 and st["active"][-1] == 0
    # the next candidate comes from the ACCEPTED state with the updated radius
    s2, m2, int2 = em.solve_tr_subproblem(torch.tensor(st["g"]), torch.tensor(st["H"]), torch.tensor(st["delta"]))
    act = st["active"].astype(bool)
    assert act[:5].all()
    assert np.allclose(st["x_new"][act], (torch.tensor(st["x"]) + s2).numpy()[act], rtol=1e-7, atol=1e-9)
    assert np.allclose(st["m_pred"][act], m2.numpy()[act], rtol=1e-8)


def test_fused_maximize_under_emulation_matches_torch_driver():
    """The device-resident loop (newton_step_kernel, emulated) and the torch lock-step driver walk the same
    iterates: same stopping behaviour, same optimum (test/test_optimization.jl:61-67 data)."""
    import emul_lib
    images, patches, vp, _ = synthetic.gen_sample_galaxy_dataset()
    pl = PlanLike([[1]], [[1]])
    kw = dict(include_kl=True, device="cpu", loc_width=1.0, max_iters=12)
    ref = em.BatchMaximizer(pl, np.concatenate(vp), runner=OracleRunner(images, patches, pl), **kw).run()
    bm = em.BatchMaximizer(pl, np.concatenate(vp), runner=OracleRunner(images, patches, pl),
                           stepper=emul_lib.newton_stepper, **kw)
    assert bm.fused
    res = bm.run()
    assert res.total_steps == ref.total_steps and np.array_equal(res.iterations, ref.iterations)
    assert np.array_equal(res.f_calls, ref.f_calls) and np.array_equal(res.converged, ref.converged)
    assert res.value[0] == pytest.approx(ref.value[0], rel=1e-9)
    assert np.allclose(res.vp, ref.vp, rtol=1e-6, atol=1e-8)
