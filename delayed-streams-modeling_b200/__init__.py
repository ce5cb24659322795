"""dsm_b200: B200-native (sm_100a) batched streaming-STT hot path behind a C ABI.

The product is ``csrc/`` (CUDA kernels + the ``extern "C"`` boundary declared in ``include/dsm_b200.h``)
built into ``libdsm_b200.so``.  The Python modules here are a thin ctypes mirror of the reference's
``moshi`` crate surface (``Mimi::encode_step``, ``LmModel::forward``, ``asr::State::step_*``) used by the
tests and the benchmark; they never compute on the CPU and raise if the library is missing.
"""
__version__ = "0.1.0"
