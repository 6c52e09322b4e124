from promp_b200.utils import logger  # noqa: F401
from promp_b200.utils.utils import set_seed  # noqa: F401
