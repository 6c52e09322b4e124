"""Device-backed baselines: the fit / predict of LinearFeatureBaseline runs inside promp_process_samples (one Gram + Cholesky
per task on the GPU); ZeroBaseline selects the no-baseline path of the same kernel."""
from promp_b200.baselines.linear_baseline import LinearFeatureBaseline  # noqa: F401
from promp_b200.baselines.zero_baseline import ZeroBaseline  # noqa: F401
