"""Re-export of the synthetic model definitions (they live in /workloads.py so that the
product benchmark does not import anything from oracle/)."""
from workloads import *  # noqa: F401,F403
from workloads import MODEL_ZOO  # noqa: F401
