"""Drop-in counterpart of the reference's `nets` package for the inference hot path
(分割/nets, 分类/nets): same constructors, attribute tree and state_dict keys; forward runs on the engine."""
from .basicUnet import UNetTaskAligWeight  # noqa: F401
