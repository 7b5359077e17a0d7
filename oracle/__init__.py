"""CPU oracle for the Neural-SDE integration hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in plain eager PyTorch on the CPU, the algorithm of the
reference hot path ``torchsde.sdeint(Diffusion_model, ...)``:

* ``oracle.spline``       - torchcde 0.2.5 ``CubicSpline.evaluate`` and the two
                            coefficient builders the reference feeds it
                            (Hermite/backward-differences, natural cubic spline).
* ``oracle.vector_field`` - ``Diffusion_model.f/g`` for every
                            ``input_option`` x ``noise_option``
                            (reference benchmark_classification/models_sde/neuralsde.py:123-307)
                            and the tutorial ``NeuralLSDEFunc``.
* ``oracle.solver``       - torchsde 0.2.5 fixed-step ``integrate`` + ``Euler.step`` +
                            diagonal Ito ``Milstein.step`` (vjp with ``create_graph`` under
                            autograd, as ``ForwardSDE.gdg_prod_diagonal`` takes it) +
                            ``SRK.diagonal_or_scalar_step`` (SRID2) + ``names=`` remapping +
                            explicit-increment Brownian source; ``sdeint_with_grad`` leaves
                            autograd on (the oracle of the engine's backward passes).
* ``oracle.latent``       - the LatentSDE augmented system ``f_aug``/``g_aug`` and its forward
                            (reference torch-ists/torch_ists/diff_module/NSDE/latent_sde.py:24-147).
* ``oracle.philox``       - numpy replica of the Philox4x32-10 counter RNG bits.
* ``oracle.wrapper``      - ``NeuralSDE.forward`` output-time selection / gather
                            (reference neuralsde.py:84-120) on top of the oracle solver.

Who may import this package: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` - as the *checker*
or the *timed CPU baseline*, never as the product.  The product package
(``stable-neural-sdes_b200``) must not import it and fails loudly when the CUDA
extension is missing.

PARITY PINNING STATUS
---------------------
* ``Diffusion_model.f/g``: PINNED.  ``tests/golden/make_golden.py`` imports the
  reference's *own unmodified* ``Diffusion_model`` (with two-symbol shims for the
  absent torchsde/torchcde/controldiffeq imports) and freezes its f/g outputs for
  all 140 option pairs under a seeded init into ``tests/golden/fg_golden.pt``;
  ``tests/test_oracle_golden.py`` checks this oracle against them.
* ``LatentSDE.f_aug/g_aug`` and ``LatentSDE.forward`` (given the oracle solver): PINNED the same way
  (``latent_golden.pt``, minted from the reference's own class).
* natural cubic spline coefficients + evaluate: PINNED against the reference's
  in-tree ``controldiffeq.interpolate`` (same script, ``spline_golden.pt``).
* torchsde 0.2.5 (``integrate``/``Euler``/``Milstein``/``linear_interp``) and
  torchcde 0.2.5 (``CubicSpline``, Hermite builder): **parity unpinned**.  Both
  packages are absent from /root/reference (un-vendored pip dependencies pinned
  in environment.yml:20-21), cannot be installed (no network), and the reference
  holds no golden vectors for them.  They are restated here from their published
  v0.2.5 algorithm, anchored on the reference's call sites
  (neuralsde.py:78-82,184,296) and checked by closed-form/structural tests only.
"""

ORACLE_IS_TEST_INFRASTRUCTURE = True
