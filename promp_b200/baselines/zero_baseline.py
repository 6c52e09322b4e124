"""ZeroBaseline (ref: meta_policy_search/baselines/zero_baseline.py:5-55): selects baseline_kind 0 in
promp_process_samples."""
import numpy as np


class ZeroBaseline(object):
    device_kind = 0

    def get_param_values(self, **kwargs):
        return None

    def set_param_values(self, value, **kwargs):
        pass

    def set_params(self, value, **kwargs):
        pass

    def fit(self, paths, **kwargs):
        pass

    def predict(self, path):
        return np.zeros_like(path["rewards"])

    def log_diagnostics(self, paths, prefix=''):
        pass
