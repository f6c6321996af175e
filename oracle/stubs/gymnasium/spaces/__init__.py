from gym.spaces import Space, Box, Discrete, Dict, Tuple, MultiDiscrete
from . import box
