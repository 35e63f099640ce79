from .mpg_learner import MPGLearner
from .nadp import NADPLearner

__all__ = ['MPGLearner', 'NADPLearner']
