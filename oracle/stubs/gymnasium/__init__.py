"""Stand-in for `gymnasium==0.28.1`, which the reference only uses for `spaces` (see gym stub)."""
from . import spaces, utils
