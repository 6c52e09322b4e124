from promp_b200.baselines.linear_baseline import LinearFeatureBaseline  # noqa: F401
from promp_b200.baselines.zero_baseline import ZeroBaseline  # noqa: F401
