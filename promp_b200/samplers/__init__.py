from promp_b200.samplers.meta_sampler import MetaSampler  # noqa: F401
from promp_b200.samplers.meta_sample_processor import MetaSampleProcessor  # noqa: F401
from promp_b200.samplers.vectorized_env_executor import MetaDeviceEnvExecutor  # noqa: F401
