"""CPU oracle for the SGMCMCJax hot path — TEST INFRASTRUCTURE, NOT PRODUCT.

This package is a plain-numpy restatement of what the reference
(jeremiecoullon/SGMCMCJax v0.2.13, pure Python on jax==0.4.13) computes on the
path  minibatch-gradient estimation -> diffusion update -> K-step sampling
loop, plus ``ksd.imq_KSD``.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and
only as the checker or the timed CPU baseline.  The product package
``sgmcmcjax_b200`` never imports it and has no CPU fallback.

Parity status
-------------
The reference cannot be imported here (jax / jaxlib / optax are not installed
and cannot be: no network, no wheel), and its arithmetic lives in the
third-party dependency ``jax==0.4.13`` / ``jaxlib==0.4.13``
(``requirements.txt:3-4``) whose source is not under ``/root/reference``.
The reference's own tests pin no numerical values on this path (shape / NaN
smoke only; ``ksd.py`` has no test at all).

* PINNED (external known answers, ``tests/golden/jax_prng_kat.json``):
  threefry2x32 (Random123 vectors), ``random.split``, ``random.uniform``,
  ``random.normal`` (incl. XLA's fp32 ``erf_inv`` polynomial) for the
  documented ``PRNGKey(0)`` / ``PRNGKey(42)`` values.
* PARITY UNPINNED: ``random.randint`` / ``random.choice`` (restated from the
  published jax 0.4.13 algorithm; no external vector exists) and every
  floating-point trajectory (defined by the reference's formulas, restated
  here with file:line citations; no reference-run fixture can be generated in
  this image).  Self-derived vectors are frozen under ``tests/golden/`` as
  regression anchors and are labelled as such.
"""
