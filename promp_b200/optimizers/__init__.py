from promp_b200.optimizers.maml_first_order_optimizer import MAMLPPOOptimizer  # noqa: F401
from promp_b200.optimizers.conjugate_gradient_optimizer import ConjugateGradientOptimizer  # noqa: F401
