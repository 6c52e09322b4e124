from promp_b200.policies.meta_gaussian_mlp_policy import MetaGaussianMLPPolicy  # noqa: F401
from promp_b200.policies.distributions import DiagonalGaussian  # noqa: F401
