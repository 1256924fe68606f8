"""hyquas_b200: B200-native (sm_100a) state-vector simulator behind the HyQuas Circuit/Gate API.

The package is a thin ctypes shell; all work happens in libhyquas_b200.so (hyquas_b200/csrc).
`import hyquas_b200.circuits` (QASM generators) works without the library; everything else needs it.
"""
__all__ = ["circuits"]
