import numpy as np  # noqa: F401  (ZoneEnvBase relies on the star import for nothing else)
from oracle.sg_engine import *  # noqa: F401,F403
from oracle.sg_engine import Engine, ResamplingError  # noqa: F401
