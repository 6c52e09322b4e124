from promp_b200.meta_algos.pro_mp import ProMP  # noqa: F401
from promp_b200.meta_algos.trpo_maml import TRPOMAML  # noqa: F401
from promp_b200.meta_algos.vpg_maml import VPGMAML  # noqa: F401
