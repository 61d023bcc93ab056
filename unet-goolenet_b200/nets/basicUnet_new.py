"""分类/nets/basicUnet_new.py is a copy of 分割/nets/basicUnet.py; same shells."""
from .basicUnet import *  # noqa: F401,F403
from .basicUnet import UNetTaskAligWeight  # noqa: F401
