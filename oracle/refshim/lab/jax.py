from . import *  # noqa: F401,F403   (`import lab.jax as B` only registers the JAX backend in the real package)
from . import sum, epsilon, pi  # noqa: F401
