"""pydnmfk_b200 -- B200-native implementation of pyDNMFk's distributed NMF update loop.

Same class / argument surface as the reference for that path (``PyNMF``, ``PyNMFk``/``sample``,
``nmf_algorithms_1D/2D``, ``MPI_comm``, ``parse``/``var_init``/``determine_block_params``/
``data_operations``); the arithmetic runs in hand-written sm_100a CUDA kernels behind a C-ABI
shared library (``libdnmf.so``, ``include/dnmf.h``) called through ctypes.  No CPU fallback.
"""
__version__ = '0.1.0'
