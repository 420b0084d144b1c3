"""deepfly3d_b200: B200-native (sm_100a) 2D->3D pose path behind the df3d.core.Core API surface.

Importing the package does not load CUDA; ``deepfly3d_b200._lib`` (pulled in by ``ops``,
``hourglass``, ``core``) dlopens ``libdf3d_b200.so`` and fails loudly if it has not been built.
"""
__version__ = "0.1.0"
