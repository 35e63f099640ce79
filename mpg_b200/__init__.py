"""mpg_b200 -- B200-native (sm_100a) implementation of the MPG model-based learner hot path.

Host side mirrors the reference's plugin interfaces:
    mpg_b200.learners.MPGLearner / NADPLearner   (learners/mpg_learner.py, learners/nadp.py)
    mpg_b200.envs_and_models.NAME2MODELCLS      (envs_and_models/__init__.py)
    mpg_b200.policy.PolicyWithQs                (policy.py)
All arithmetic runs in hand-written CUDA kernels behind a C ABI (include/mpg_b200.h,
mpg_b200/libmpg_b200.so).  There is no CPU fallback.
"""
__version__ = '0.1.0'
