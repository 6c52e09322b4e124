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
* TF1 half: the graph itself (MLP forward, tf.gradients through the inner steps, TF1 Adam) is PARITY UNPINNED by
  reference outputs (TF1 cannot run here).  Pinned pieces: the distribution math and the conjugate-gradient solver,
  whose numpy twins inside the reference's TF-half modules (DiagonalGaussian.kl / log_likelihood / entropy,
  optimizers.conjugate_gradients) ARE executed by make_golden.py (tests/golden/tf_half_known.npz).  The rest is pinned
  only by (i) the reference tests that touch it and can be restated (likelihood ratio == 1 at the first inner step,
  tests/test_integration.py:150-175; get_actions == distribution_info, tests/test_policies.py:43-64) and (ii) fp64
  finite-difference checks of every gradient.
* HalfCheetah surrogate dynamics: PARITY UNPINNED (new model; MuJoCo absent).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this package.  The product (``promp_b200``) never does.
"""
