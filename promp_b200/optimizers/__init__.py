"""Outer-loop optimizers: TF1-style Adam on the device (promp_adam_tf1) and the host-side CG / line search of TRPO-MAML, both
fed by device passes of the meta objective."""
from promp_b200.optimizers.maml_first_order_optimizer import MAMLPPOOptimizer  # noqa: F401
from promp_b200.optimizers.conjugate_gradient_optimizer import ConjugateGradientOptimizer  # noqa: F401
