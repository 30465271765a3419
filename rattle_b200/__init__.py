"""rattle_b200 — B200-native (sm_100a) build of RATTLE's two hot paths behind a C ABI.

This package is the Python host-side mirror of the reference's two entry points
(`cluster_reads`, /root/reference/cluster.hpp:44 and `correct_reads`, /root/reference/correct.hpp:44) over
`librattle_b200.so` (include/rattle_b200.h).  Everything computes on the GPU; there is no CPU fallback:
loading fails loudly when the CUDA library has not been built, and `Context()` fails when no B200 is visible.
"""
import os as _os

# the POA engine drives up to 8 units x (1 + 4) CUDA streams; give them their own hardware queues (read by the CUDA
# runtime when the context is created, so it must be set before the first CUDA call of the process)
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from .api import (Context, ClusterSet, RattleError, cluster_reads, correct_reads, hps_decode, hps_encode, lib_path,
                  load_library)

__all__ = ["Context", "ClusterSet", "RattleError", "cluster_reads", "correct_reads", "hps_encode", "hps_decode",
           "lib_path", "load_library"]
