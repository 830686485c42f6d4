"""Batched Newton trust-region maximisation of the ELBO -- rows (f.1)/(f.2) of SURVEY.md section 8.

Mirror of src/deterministic_vi/ElboMaximize.jl: `maximize!(ea, vp)` (:228-242) drives
Optim.NewtonTrustRegion (initial_delta = 1, delta_hat = 1e9, :105-108) on the 41 free parameters with
`Optim.Options(x_tol = 1e-7, f_tol = 1e-6, g_tol = 1e-8, iterations = 50)` (:95-103); each new iterate costs
one `evaluate!` (:161-172) = to_bound! -> elbo (likelihood - KL) -> propagate_derivatives!.

The reference optimises one source per thread, one ELBO evaluation at a time (~10 us of GPU work each).
Here ALL sources of a batch step in lock-step: one iteration = one CUDA plan evaluation (value + gradient +
Hessian for every source) + ONE newton_step_kernel launch (csrc/maximize_kernels.cuh: -KL, the 44 -> 41
propagate_derivatives!, the trust-region update, the next 41 x 41 subproblem and to_bound! of the next
candidate, one block per source); converged sources are masked out of both.  Nothing leaves the device
between iterations except the "anyone still active?" flag.  This module's torch code (kl.py,
constraint_transforms.py, solve_tr_subproblem, BatchMaximizer._run) is the same algorithm for CPU tensors:
it is what the tests compare the kernels against and what runs when a test injects a CPU `runner`.

Optim.jl is an un-vendored dependency (REQUIRE:13, >= 0.7.4): its NewtonTrustRegion is restated from the
published algorithm (Nocedal & Wright, Alg. 4.1 for the radius update and the exact subproblem solution by
the secular equation in the eigenbasis, incl. the hard case); iterates need not match Optim's bit for bit,
the optimum and the stopping rules do.  PARITY UNPINNED for this row (no reference run is possible offline).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import constraint_transforms as ct
from .kl import KLTerm

ETA, RHO_LOWER, RHO_UPPER = 0.1, 0.25, 0.75     # Optim.NewtonTrustRegion defaults
X_TOL, F_TOL, G_TOL, MAX_ITERS = 1e-7, 1e-6, 1e-8, 50
INITIAL_DELTA, DELTA_HAT = 1.0, 1e9


def solve_tr_subproblem(g: torch.Tensor, H: torch.Tensor, delta: torch.Tensor, mask: Optional[torch.Tensor] = None):
    """min_s g's + 1/2 s'Hs  s.t. |s| <= delta, batched (B x n, B x n x n, B).
    Returns (s, m = predicted change, interior flag).  `mask` (uint8, CUDA only): sources with 0 are skipped
    and get s = 0, m = 0."""
    B, n = g.shape
    dev_in = g.device
    if dev_in.type == "cuda":
        # cuSOLVER's batched FP64 eigensolver and LAPACK on the host both cost ~0.5 s per 1000 sources
        # (100x the ELBO evaluation), so the device path is the library's own kernel
        # (csrc/newton_kernels.cuh: one block per source, Jacobi + secular equation).
        from . import _lib
        g, H, delta = g.contiguous(), H.contiguous(), delta.contiguous()
        s = torch.zeros_like(g)
        m = torch.zeros_like(delta)
        interior = torch.zeros(B, dtype=torch.int32, device=dev_in)
        _lib.check(_lib.load().celeste_tr_subproblem(B, n, g.data_ptr(), H.data_ptr(), delta.data_ptr(),
                                                     mask.data_ptr() if mask is not None else None,
                                                     s.data_ptr(), m.data_ptr(), interior.data_ptr(),
                                                     torch.cuda.current_stream().cuda_stream))
        return s, m, interior.bool()
    ev, Q = torch.linalg.eigh(H)                       # ascending
    qg = torch.einsum("bij,bi->bj", Q, g)              # Q' g
    d2 = delta * delta
    lam_min = ev[:, 0]

    def pnorm2(lam):                                   # |s(lam)|^2 = sum (qg_i / (ev_i + lam))^2
        den = ev + lam[:, None]
        return ((qg / den) ** 2).sum(dim=1)

    zero = torch.zeros_like(delta)
    pos_def = lam_min >= 1e-8
    interior = pos_def & (pnorm2(torch.where(pos_def, zero, 1.0 - lam_min)) <= d2)
    # boundary solution: find lam > max(0, -lam_min) with |s(lam)| = delta (secular equation, Newton on 1/|s|)
    lam_lb = torch.clamp(-lam_min, min=0.0)
    tiny = 1e-12 * (1.0 + ev.abs().max(dim=1).values)
    # hard case: g orthogonal to the eigenspace of lam_min and |s(-lam_min)| < delta
    scale = qg.abs().max(dim=1, keepdim=True).values + 1e-300
    at_min = (ev - lam_min[:, None]).abs() <= 1e-12 * (1.0 + ev.abs())
    ortho = ((qg.abs() <= 1e-12 * scale) | ~at_min).all(dim=1)
    den_h = torch.where(at_min, torch.ones_like(ev), ev - lam_min[:, None])
    s_h_coef = torch.where(at_min, torch.zeros_like(qg), -qg / den_h)
    ph2 = (s_h_coef ** 2).sum(dim=1)
    hard = (~interior) & ortho & (lam_min <= 1e-8) & (ph2 <= d2)
    lam = lam_lb + tiny                                # start just above the pole: Newton on 1/|s| is monotone
    for _ in range(60):
        den = ev + lam[:, None]
        p2 = ((qg / den) ** 2).sum(dim=1)
        pn = torch.sqrt(p2)
        # d|s|^2/dlam = -2 sum qg^2/(ev+lam)^3
        dp2 = -2.0 * ((qg ** 2) / den ** 3).sum(dim=1)
        # Newton on phi(lam) = 1/delta - 1/|s|:  lam += (|s| - delta)/delta * |s|^2 / (-0.5 dp2)
        step = (pn - delta) / delta * p2 / (-0.5 * dp2 + 1e-300)
        new = lam + step
        new = torch.where(new <= lam_lb, 0.5 * (lam + lam_lb) + tiny, new)
        done = (step.abs() <= 1e-12 * (1.0 + lam.abs()))
        lam = torch.where(done | interior | hard, lam, new)
    lam = torch.where(interior, zero, lam)
    lam = torch.where(hard, -lam_min, lam)
    den = ev + lam[:, None]
    coef = -qg / torch.where(den.abs() < 1e-300, torch.full_like(den, 1e-300), den)
    coef = torch.where(hard[:, None], s_h_coef, coef)
    tau = torch.sqrt(torch.clamp(d2 - ph2, min=0.0))
    coef[:, 0] = torch.where(hard, tau, coef[:, 0])   # move along the lowest eigenvector to the boundary
    s = torch.einsum("bij,bj->bi", Q, coef)
    m = (g * s).sum(dim=1) + 0.5 * torch.einsum("bi,bij,bj->b", s, H, s)
    return s, m, interior


@dataclass
class MaximizeResult:
    vp: np.ndarray            # n x 44 optimised bound parameters (active sources)
    value: np.ndarray         # n  maximised ELBO (likelihood - KL)
    iterations: np.ndarray    # n  Newton iterations taken
    f_calls: np.ndarray       # n  ELBO evaluations
    converged: np.ndarray     # n  bool
    total_steps: int          # lock-step iterations of the batch (== plan evaluations - 1)


class BatchMaximizer:
    """maximize! for every task of a plan at once (Sa = 1 per task; neighbours frozen during the solve,
    as inside one `maximize!` call of the reference)."""

    def __init__(self, plan, vp_flat: np.ndarray, include_kl: bool = True, device: Optional[str] = None,
                 loc_width: float = 1e-4, max_iters: int = MAX_ITERS, runner=None, stepper=None,
                 fused: Optional[bool] = None, box=None):
        """`runner(self)` evaluates the plan's tasks at self.vp_all and fills self.v/d/h/flags; the default
        launches the CUDA plan on the current stream.  (Tests inject a CPU checker here.)
        `stepper(phase, n, buffers)` runs newton_step_kernel; the default is the library on the current stream
        (tests inject the host-emulated kernel).  `fused` = use newton_step_kernel (default on CUDA).
        `box` = (lo, hi), n x 26 bounds built earlier by ct.box_bounds: a caller that optimises the same source
        several times (joint inference) passes the box of the FIRST visit so that the position constraint does not
        follow the source (one ElboConfig per target for all sweeps, ParallelRun.jl:99-101)."""
        self.plan = plan
        self.runner = runner
        self.stepper = stepper
        self.dev = torch.device(device or "cuda")
        self.n = plan.n_tasks
        self.max_iters = max_iters
        self.include_kl = include_kl
        self.kl = KLTerm(self.dev) if include_kl else None
        dt = torch.float64
        self.vp_all = torch.as_tensor(vp_flat, dtype=dt, device=self.dev).reshape(-1, 44).clone()   # n_slots x 44
        # slot of the active source of each task
        task_ptr = plan.task_ptr.astype(np.int64)
        act = plan.act.astype(np.int64)
        self.aslot = torch.as_tensor(task_ptr[:-1] + act - 1, device=self.dev)
        n = self.n
        self.v = torch.zeros(n, dtype=dt, device=self.dev)
        self.d = torch.zeros(n * 44, dtype=dt, device=self.dev)
        if fused is None:
            fused = stepper is not None or (self.dev.type == "cuda" and runner is None)
        self.fused = fused
        # the device-resident loop takes the Hessian in the packed layout (406 doubles per source instead of 1936)
        self.packed = bool(fused and runner is None and stepper is None and hasattr(plan, "set_hessian_layout"))
        self.h = torch.zeros(n * (406 if self.packed else 44 * 44), dtype=dt, device=self.dev)
        self.cnt = torch.zeros(2 * n, dtype=torch.int64, device=self.dev)
        self.flags = torch.zeros(n, dtype=torch.int32, device=self.dev)
        vp0 = self.vp_all[self.aslot]
        if box is None:
            self.lo, self.hi = ct.box_bounds(vp0, loc_width)      # position box fixed at the start (ElboMaximize.jl:68-71)
        else:
            self.lo = torch.as_tensor(box[0], dtype=dt, device=self.dev).reshape(self.n, ct.N_BOX).contiguous()
            self.hi = torch.as_tensor(box[1], dtype=dt, device=self.dev).reshape(self.n, ct.N_BOX).contiguous()
        vp0 = ct.enforce(vp0, self.lo, self.hi)                   # enforce! :230
        self.x = ct.to_free(vp0, self.lo, self.hi)                # to_free! :231
        self.f_calls = 0
        # device-side mask of sources still iterating: the kernels skip converged sources' tasks
        self.mask = torch.ones(n, dtype=torch.uint8, device=self.dev)
        self.use_mask = runner is None and hasattr(plan, "set_task_mask")
        self.profile = None        # set to {} to collect wall-clock per phase (synchronising; diagnostics only)

    def _tick(self, name, t0):
        if self.profile is not None:
            import time
            if self.dev.type == "cuda":
                torch.cuda.synchronize()
            self.profile[name] = self.profile.get(name, 0.0) + time.perf_counter() - t0
            return time.perf_counter()
        return t0

    def evaluate(self, x: torch.Tensor):
        """evaluate! (ElboMaximize.jl:161-172): negative ELBO, gradient and Hessian in free coordinates."""
        import time
        t0 = time.perf_counter() if self.profile is not None else 0.0
        bound = ct.to_bound(x, self.lo, self.hi)
        self.vp_all[self.aslot] = bound
        t0 = self._tick("to_bound", t0)
        if self.runner is not None:
            self.runner(self)
        else:
            self._run_plan(torch.cuda.current_stream())
        v = self.v.clone()
        g = self.d.reshape(self.n, 44).clone()
        H = self.h.reshape(self.n, 44, 44).clone()
        t0 = self._tick("elbo_plan", t0)
        if self.include_kl:
            kv, kg, kH = self.kl(bound, order=2)
            v += kv
            g += kg
            H += kH
        t0 = self._tick("kl", t0)
        gf, Hf = ct.propagate_derivatives(x, self.lo, self.hi, g, H)
        t0 = self._tick("propagate", t0)
        self.f_calls += 1
        bad = self.flags != 0
        return -v, -gf, -Hf, bad, bound

    def _run_plan(self, stream):
        self.plan.run_device(self.vp_all.data_ptr(), 2, self.v.data_ptr(), self.d.data_ptr(), self.h.data_ptr(),
                             self.cnt.data_ptr(), self.flags.data_ptr(), stream=stream.cuda_stream)

    def run(self) -> MaximizeResult:
        n, dev = self.n, self.dev
        x = self.x
        if self.use_mask:
            self.plan.set_task_mask(self.mask.data_ptr())
        if self.packed:
            self.plan.set_hessian_layout(True)
        try:
            return self._run_fused() if self.fused else self._run(n, dev, x)
        finally:
            if self.use_mask:
                self.plan.set_task_mask(0)
            if self.packed:
                self.plan.set_hessian_layout(False)

    def _step(self, phase):
        if self.stepper is not None:
            self.stepper(phase, self.n, self._buffers)
            return
        from . import _lib
        _lib.check(_lib.load().celeste_newton_step(phase, self.n, self._buffers, torch.cuda.current_stream().cuda_stream))

    def _evaluate_plan(self):
        import time
        t0 = time.perf_counter() if self.profile is not None else 0.0
        if self.runner is not None:
            self.runner(self)
        else:
            self._run_plan(torch.cuda.current_stream())
        self.f_calls += 1
        self._tick("elbo_plan", t0)

    def _run_fused(self) -> MaximizeResult:
        """The device-resident loop: per iteration one plan evaluation + one newton_step_kernel launch."""
        import time
        from . import _lib
        n, dev, f64 = self.n, self.dev, torch.float64
        z = lambda *shape, dtype=f64: torch.zeros(shape, dtype=dtype, device=dev)
        st = dict(x=self.x.contiguous(), f=z(n), g=z(n, ct.N_FREE), H=z(n, ct.N_FREE, ct.N_FREE), delta=z(n),
                  x_new=z(n, ct.N_FREE), m_pred=z(n), interior=z(n, dtype=torch.int32), active=self.mask,
                  converged=z(n, dtype=torch.uint8), iters=z(n, dtype=torch.int32), f_calls=z(n, dtype=torch.int32),
                  lo=self.lo.contiguous(), hi=self.hi.contiguous(), v=self.v, d=self.d, h=self.h, flags=self.flags,
                  vp_all=self.vp_all, aslot=self.aslot.contiguous(),
                  prior=self.kl.packed() if self.include_kl else None)
        self._st = st                                          # keeps the tensors alive while the kernels run
        self._buffers = _lib.celeste_newton_buffers(**{k: (t.data_ptr() if t is not None else None) for k, t in st.items()},
                                                    h_layout=1 if self.packed else 0)
        self.mask.fill_(1)
        self._step(2)                                          # vp_all[aslot] <- to_bound(x)
        self._evaluate_plan()
        t0 = time.perf_counter() if self.profile is not None else 0.0
        self._step(0)
        self._tick("newton_step", t0)
        steps = 0
        for _ in range(self.max_iters):
            if not bool(self.mask.any()):
                break
            steps += 1
            self._evaluate_plan()
            t0 = time.perf_counter() if self.profile is not None else 0.0
            self._step(1)
            self._tick("newton_step", t0)
        self._step(2)                                          # maximize! :239-240
        self.x = st["x"]
        bound = self.vp_all[self.aslot]
        res = MaximizeResult(bound.cpu().numpy(), (-st["f"]).cpu().numpy(), st["iters"].cpu().numpy().astype(np.int64),
                             st["f_calls"].cpu().numpy().astype(np.int64), st["converged"].cpu().numpy().astype(bool), steps)
        self.mask.fill_(1)
        return res

    def _run(self, n, dev, x):
        f, g, H, bad, bound = self.evaluate(x)
        delta = torch.full((n,), INITIAL_DELTA, dtype=torch.float64, device=dev)
        active = ~bad & (g.abs().max(dim=1).values >= G_TOL)       # initial g_tol check
        converged = ~bad & ~active
        iters = torch.zeros(n, dtype=torch.int64, device=dev)
        fcalls = torch.ones(n, dtype=torch.int64, device=dev)
        steps = 0
        for _ in range(self.max_iters):
            if not bool(active.any()):
                break
            steps += 1
            import time
            t0 = time.perf_counter() if self.profile is not None else 0.0
            self.mask.copy_(active)          # inactive sources keep their last outputs; `accept` ignores them
            s, m, interior = solve_tr_subproblem(g, H, delta, self.mask if dev.type == "cuda" else None)
            self._tick("tr_subproblem", t0)
            x_new = torch.where(active[:, None], x + s, x)
            f_new, g_new, H_new, bad_new, _ = self.evaluate(x_new)
            fcalls += active.to(torch.int64)
            f_diff = f - f_new
            eps = torch.finfo(torch.float64).eps
            rho = torch.where(m.abs() <= eps, torch.ones_like(m),
                              torch.where(m > 0, torch.full_like(m, RHO_LOWER - 1.0), f_diff / (-m)))
            rho = torch.where(bad_new | ~torch.isfinite(f_new), torch.full_like(rho, RHO_LOWER - 1.0), rho)
            shrink = rho < RHO_LOWER
            grow = (rho > RHO_UPPER) & ~interior
            delta = torch.where(active & shrink, delta * 0.25, delta)
            delta = torch.where(active & grow, torch.clamp(2 * delta, max=DELTA_HAT), delta)
            accept = active & (rho > ETA)
            # convergence is assessed only on accepted steps (Optim assess_convergence for NewtonTrustRegion)
            x_conv = (x_new - x).abs().max(dim=1).values < X_TOL
            f_conv = (f_new - f).abs() <= F_TOL * f_new.abs()
            g_conv = g_new.abs().max(dim=1).values < G_TOL
            newly = accept & (x_conv | f_conv | g_conv)
            a1, a2 = accept[:, None], accept[:, None, None]
            x = torch.where(a1, x_new, x)
            f = torch.where(accept, f_new, f)
            g = torch.where(a1, g_new, g)
            H = torch.where(a2, H_new, H)
            iters += active.to(torch.int64)
            converged = converged | newly
            # a trust region collapsed to nothing cannot make progress any more
            dead = active & (delta < 1e-14)
            active = active & ~newly & ~dead
        self.x = x
        self.mask.fill_(1)
        bound = ct.to_bound(x, self.lo, self.hi)                  # maximize! :239-240
        self.vp_all[self.aslot] = bound
        return MaximizeResult(bound.cpu().numpy(), (-f).cpu().numpy(), iters.cpu().numpy(), fcalls.cpu().numpy(),
                              converged.cpu().numpy(), steps)
