"""CPU oracle for the ProMP hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement of the reference algorithm (jonasrothfuss/ProMP @ 93ae339):

* ``numpy_half``  - envs, NormalizedEnv affine/clip, iterative vec-env executor, the sampler loop,
  discounted returns, LinearFeatureBaseline, GAE, advantage normalisation (numpy, float64 like
  the reference; every function cites the reference file:line it follows).
* ``tf_half``     - the TF1 graph half (Gaussian MLP policy, DiagonalGaussian, inner adapt step,
  ProMP / TRPO-MAML meta objectives, TF1 Adam, CG + finite-difference HVP) restated on PyTorch-CPU
  autograd, because TensorFlow 1.x is not installable here.
* ``cheetah_surrogate`` - the *specification* of the MuJoCo-free HalfCheetahRandDirec surrogate
  (its dynamics are new, defined by this repo; obs/act dims, reward decomposition, task sampling,
  reset noise and env_infos keys follow the reference).

Pinning status
--------------
* numpy half: PINNED.  ``oracle/make_golden.py`` imports the unmodified reference classes from
  ``/root/reference`` (with the stub packages under ``oracle/stubs``) and writes their outputs to
  ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this restatement against them,
  and against the known-answer vectors of the reference's own tests/test_samplers.py.
* TF1 half: PINNED to the reference's own graph code.  ``oracle/stubs_tf/tensorflow`` is a torch-backed, lazily
  evaluated stand-in for the ~45 TF-1.x symbols the reference uses (placeholders, variables, dense layers, math ops,
  ``tf.gradients`` via ``torch.autograd.grad(create_graph=True)``, ``Session.run``, the published TF1 Adam rule), so
  ``make_golden.py`` imports the UNMODIFIED ``policies/*``, ``meta_algos/{base,pro_mp,trpo_maml,vpg_maml}.py``,
  ``optimizers/*`` and ``meta_trainer.py`` from ``/root/reference`` and records what they compute:
  ``tests/golden/tf_half_graph.npz`` (adapted parameters, meta objective, inner / outer KL, second-order meta-gradient,
  5-epoch Adam end point, TRPO-MAML gradients / finite-difference Hx / CG direction / accepted step, E-MAML and VPG-MAML;
  14 cases incl. the BASELINE.json shapes 40 x 2000 and 40 x 4000, each evaluated in float32 and float64) and
  ``tests/golden/trainer_run.npz`` (3 meta-iterations of the unmodified ``Trainer`` over the unmodified sampler /
  processor / baseline / policy / ProMP).  ``tests/test_tf_golden.py`` checks ``tf_half.py`` against them (float32:
  2e-5, float64: 1e-6 - the reference rounds ``inner_lr`` and ``log(2 pi)`` to float32) and the CUDA path too.  What the
  stand-in cannot reproduce is TensorFlow's floating-point kernels (Eigen vs torch CPU: last-bit differences) and its
  random streams (action noise is recorded and injected instead).  The numpy twins inside the TF-half modules
  (DiagonalGaussian.kl / log_likelihood / entropy, conjugate_gradients, _adapt_kl_coeff) are executed directly
  (``tests/golden/tf_half_known.npz``).
* HalfCheetah surrogate dynamics: a NEW analytic model (MuJoCo is absent and the north star asks for a MuJoCo-free
  surrogate), so there is no reference output to pin its dynamics to; obs/act dims, reward decomposition, task
  sampling, reset noise and env_infos keys follow the reference source.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this package.  The product (``promp_b200``) never does.
"""
