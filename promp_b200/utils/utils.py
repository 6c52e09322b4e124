"""Host helpers mirroring meta_policy_search/utils/utils.py where the hot path needs them."""
import random

import numpy as np


def set_seed(seed):
    """ref: utils/utils.py:161-177 (without the TF seed: policy noise is Philox, keyed by the sampler)."""
    seed %= 4294967294
    random.seed(seed)
    np.random.seed(seed)
    print('using seed %s' % (str(seed)))
