"""gtsam_b200 — B200-native Gauss-Newton / Levenberg-Marquardt inner loop behind the GTSAM optimizer API.

Host-side Python mirror of the reference interface over the C-ABI library (``include/gtsam_b200.h``,
``libgtsam_b200.so``); the C++ drop-in lives in ``gtsam_b200/shim``.  Submodules:

* ``problem``   typed factor tables / packed Values (the C-ABI's problem description)
* ``capi``      ctypes bindings: ``Context``, ``DeviceProblem``
* ``optimizer`` ``LevenbergMarquardtOptimizer``, ``GaussNewtonOptimizer``, ``DoglegOptimizer``, ``Marginals``
* ``gnc``       ``GncOptimizer`` / ``GncParams``
* ``io``        BAL and g2o readers / writers
* ``datasets``  synthetic generators of the BASELINE configurations

Nothing is imported eagerly: ``import gtsam_b200`` must not need the shared library or a GPU.
"""
__all__ = ["problem", "capi", "optimizer", "gnc", "io", "datasets"]
